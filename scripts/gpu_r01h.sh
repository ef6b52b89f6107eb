set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r01h.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_r01h.log
( time timeout 1200 python bench.py 2>gpurun_out/bench_default.err | tee gpurun_out/bench_default.json | cut -c1-200 ) 2>&1 | tail -6
tail -3 gpurun_out/bench_default.err
timeout 600 python bench.py --workload t2d --steps 10 --warmup 3 2>>gpurun_out/bench_default.err | tee gpurun_out/bench_t2d_r01h.json | cut -c1-200
