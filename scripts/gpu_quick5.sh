set -x
mkdir -p gpurun_out
timeout 600 python bench.py --workload t3d --steps 5 --warmup 3 --no-cpu 2>gpurun_out/bench_q5.err | tee gpurun_out/bench_t3d_q5.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>>gpurun_out/bench_q5.err | tee gpurun_out/bench_t2d_q5.json
M=gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__registers_per_thread,smsp__inst_executed.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:"gather_|element_state" -s 30 -c 12 --csv --log-file gpurun_out/lane_launches_t3d.csv \
    python bench.py --workload t3d --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu3.log 2>&1
