# r01r: final validation of the round: tests, smoke, the three workloads, the reference arms
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r01r_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r01r_smoke.log
( time timeout 1200 python bench.py 2>gpurun_out/r01r_bench.err | tee gpurun_out/r01r_bench_t3d92.json | cut -c1-200 ) 2>&1 | tail -5
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/r01r_bench.err | tee gpurun_out/r01r_bench_reference_t3d.json | cut -c1-200
timeout 900 python bench.py --workload t2d 2>>gpurun_out/r01r_bench.err | tee gpurun_out/r01r_bench_t2d1024.json | cut -c1-200
timeout 900 python bench.py --workload t2d --impl reference --steps 3 --warmup 1 2>>gpurun_out/r01r_bench.err | tee gpurun_out/r01r_bench_reference_t2d.json | cut -c1-200
timeout 900 python bench.py --workload chns 2>>gpurun_out/r01r_bench.err | tee gpurun_out/r01r_bench_chns_t2d512.json | cut -c1-200
tail -5 gpurun_out/r01r_bench.err
