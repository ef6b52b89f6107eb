#!/bin/bash
# per-instruction shared-memory wavefronts / bank conflicts of one velocity-row launch (edge nodes) at T3D(92)
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'gather_lane_kernel.*0>\(' \
    --launch-skip 25 --launch-count 1 -o gpurun_out/r02i_lane3d_src -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-solve --no-parity > /dev/null 2>&1
ncu -i gpurun_out/r02i_lane3d_src.ncu-rep --page source --csv --print-source sass > gpurun_out/r02i_lane3d_source_sass.csv 2>/dev/null
rm -f gpurun_out/r02i_lane3d_src.ncu-rep
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r02i_lane3d_source_sass.csv")))
hdr = rows[0]
print(hdr)
PY
