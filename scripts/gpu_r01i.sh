# r01i: evidence for the default path (T3D): launch list of the bench command, full capture of the dominant kernels
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01i_launches_bench_t3d92.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-solve > gpurun_out/bench_under_ncu_r01i.log 2>&1
M=gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__registers_per_thread,smsp__inst_executed.sum
timeout 900 ncu --metrics $M --clock-control none -k regex:"gather_|element_state" -s 40 -c 10 --csv --log-file gpurun_out/r01i_lane_launch_metrics_t3d92.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-solve > gpurun_out/bench_under_ncu_r01i2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gather_lane_kernel|element_state" -s 30 -c 10 -f -o gpurun_out/prof_lane3d \
    python bench.py --size 48 --steps 1 --warmup 3 --no-cpu --no-solve > gpurun_out/ncu_lane3d.log 2>&1
ls -la gpurun_out | tail -8
