set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu 2>gpurun_out/bench.err | tee gpurun_out/bench_t2d.json
bash scripts/gpu_quick2.sh
