# lane-group gather kernels with element-state pre-pass: parity for every lanes-per-node setting, then a sweep
set -x
mkdir -p gpurun_out
for L in default 1 2 3; do
  if [ $L = default ]; then unset B200_GATHER_LANES; else export B200_GATHER_LANES=$L; fi
  timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_parity_L$L.log
done
for L in 1 2 3 6; do
  B200_GATHER_LANES=$L timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>gpurun_out/bench_L$L.err | tee gpurun_out/bench_t2d_L$L.json
done
for L in 1 2 5 10; do
  B200_GATHER_LANES=$L timeout 600 python bench.py --workload t3d --steps 5 --warmup 3 --no-cpu 2>>gpurun_out/bench_L$L.err | tee gpurun_out/bench_t3d_L$L.json
done
unset B200_GATHER_LANES
M=gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__registers_per_thread,smsp__inst_executed.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:"gather_|element_state" -s 24 -c 12 --csv --log-file gpurun_out/lane_launches_t2d.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu2.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:"gather_|element_state" -s 48 -c 20 --csv --log-file gpurun_out/lane_launches_t3d.csv \
    python bench.py --workload t3d --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu3.log 2>&1
