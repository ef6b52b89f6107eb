# Round-1 second GPU pass: parity tests, bench lines, ncu launch list, full captures of the gather and SpMV kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --solve 2>gpurun_out/bench.err | tee gpurun_out/bench_t2d.json
python bench.py --workload t3d --steps 5 --warmup 3 --no-cpu 2>>gpurun_out/bench.err | tee gpurun_out/bench_t3d.json
python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/bench.err | tee gpurun_out/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__registers_per_thread --clock-control none -k regex:gather_ -s 30 -c 12 --csv --log-file gpurun_out/gather_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gather_ -s 30 -c 10 -o gpurun_out/prof_gather2d \
    python bench.py --size 512 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmv -s 3 -c 1 -o gpurun_out/prof_spmv \
    python bench.py --size 512 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full_spmv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gather_ -s 12 -c 4 -o gpurun_out/prof_gather3d \
    python bench.py --workload t3d --size 32 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full3d.log 2>&1
ls -la gpurun_out
