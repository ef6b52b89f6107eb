# r01j: config 5 (CHNS) bench line + ncu capture of chns_kernel
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --workload chns --steps 5 --warmup 3 2>gpurun_out/bench_chns.err | tee gpurun_out/bench_chns_t2d512.json | cut -c1-400
tail -3 gpurun_out/bench_chns.err
timeout 600 python bench.py --workload chns --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench_chns.err | tee gpurun_out/bench_chns_reference.json | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chns_kernel -s 3 -c 1 -f -o gpurun_out/prof_chns \
    python bench.py --workload chns --size 256 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_chns.log 2>&1
tail -2 gpurun_out/ncu_chns.log
