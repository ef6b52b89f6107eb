set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chns.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --workload chns --steps 5 --warmup 3 --no-cpu 2>gpurun_out/bench_chns2.err | tee gpurun_out/bench_chns_t2d512_r128.json | cut -c1-160
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r01k_launches_bench_t3d92.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-solve > gpurun_out/bench_under_ncu_r01k.log 2>&1
wc -l gpurun_out/r01k_launches_bench_t3d92.csv
