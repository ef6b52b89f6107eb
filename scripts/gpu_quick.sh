set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu 2>gpurun_out/bench.err | tee gpurun_out/bench_t2d.json
python bench.py --workload t3d --steps 5 --warmup 3 --no-cpu 2>>gpurun_out/bench.err | tee gpurun_out/bench_t3d.json
ncu --set full --clock-control none --import-source on -k regex:gather_u -s 3 -c 2 -o gpurun_out/prof_gather2d \
    python bench.py --size 512 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
