set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__shared_mem_per_block_dynamic,launch__grid_size --clock-control none -k regex:gather_ -s 8 -c 10 --csv --log-file gpurun_out/gather_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
