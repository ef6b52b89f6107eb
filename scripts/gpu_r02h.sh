#!/bin/bash
# `ncu --set full` of the main velocity-row launch (gather_lane_kernel<3,10,4,4,10,4,1,1,0>) in an assembly pass at T3D(92)
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'gather_lane_kernel.*0>\(' \
    --launch-skip 24 --launch-count 8 -o gpurun_out/r02h_full_lane3d_urows_t3d92 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-solve --no-parity > /dev/null 2>&1
ncu -i gpurun_out/r02h_full_lane3d_urows_t3d92.ncu-rep --page raw --csv > gpurun_out/r02h_full_lane3d_urows_t3d92.csv 2>/dev/null
rm -f gpurun_out/r02h_full_lane3d_urows_t3d92.ncu-rep
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r02h_full_lane3d_urows_t3d92.csv")))
hdr = rows[0]; idx = {n: i for i, n in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
for r in rows[2:]:
    print([(w.split("__")[1][:28], r[idx[w]]) for w in want if w in idx])
PY
