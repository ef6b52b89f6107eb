# One GPU round: parity tests, bench lines, ncu launch list and one full capture of the assembly kernels.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench_t2d.json
python bench.py --steps 10 --warmup 3 --no-cpu --assembly scatter 2>>gpurun_out/bench.err | tee gpurun_out/bench_t2d_scatter.json
python bench.py --workload t3d --steps 5 --warmup 3 --no-cpu 2>>gpurun_out/bench.err | tee gpurun_out/bench_t3d.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gather_ -s 6 -c 2 -o gpurun_out/prof_gather2d \
    python bench.py --size 512 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
