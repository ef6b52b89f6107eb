# patch kernel (block-slot owners over Morton patches): parity, then bench lines for both workloads
set -x
mkdir -p gpurun_out
export B200_VERBOSE=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_parity_patch.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>gpurun_out/bench_patch.err | tee gpurun_out/bench_t2d_patch.json
timeout 600 python bench.py --workload t3d --steps 5 --warmup 3 --no-cpu 2>>gpurun_out/bench_patch.err | tee gpurun_out/bench_t3d_patch.json
tail -5 gpurun_out/bench_patch.err
