#!/bin/bash
# r02o: ncu --set full + source page of the short-row (mid-edge) velocity launch, persistent-group kernel
T=${1:-r02o}
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'gather_urow_kernel' \
    --launch-skip 6 --launch-count 2 -o gpurun_out/${T}_full_urow_t3d92 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-solve --no-parity > /dev/null 2>&1
ncu -i gpurun_out/${T}_full_urow_t3d92.ncu-rep --page raw --csv > gpurun_out/${T}_full_urow_t3d92.csv 2>/dev/null
ncu -i gpurun_out/${T}_full_urow_t3d92.ncu-rep --page source --csv --print-source sass > gpurun_out/${T}_urow_source_sass.csv 2>/dev/null
rm -f gpurun_out/${T}_full_urow_t3d92.ncu-rep
python bench.py --steps 5 --warmup 3 --no-cpu --no-solve --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
