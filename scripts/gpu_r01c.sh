# Round-1 third GPU pass: lane-per-column gather kernels vs thread-per-node kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_parity.log
for k in node lane; do
  B200_GATHER_KERNEL=$k timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>gpurun_out/bench_$k.err | tee gpurun_out/bench_t2d_$k.json
  B200_GATHER_KERNEL=$k timeout 600 python bench.py --workload t3d --steps 5 --warmup 3 --no-cpu 2>>gpurun_out/bench_$k.err | tee gpurun_out/bench_t3d_$k.json
done
timeout 600 ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__registers_per_thread,smsp__inst_executed.sum --clock-control none -k regex:"gather_|element_state" -s 24 -c 8 --csv --log-file gpurun_out/lane_launches_t2d.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__registers_per_thread,smsp__inst_executed.sum --clock-control none -k regex:"gather_|element_state" -s 24 -c 8 --csv --log-file gpurun_out/lane_launches_t3d.csv \
    python bench.py --workload t3d --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu3.log 2>&1
ls -la gpurun_out
