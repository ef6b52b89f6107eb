set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r01l.log
B200_VERBOSE=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-solve 2>&1 | grep -v "chunk merged" | cut -c1-260 | tail -8
