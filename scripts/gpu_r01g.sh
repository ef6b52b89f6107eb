set -x
mkdir -p gpurun_out
export B200_VERBOSE=1
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
( time timeout 1200 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_t3d92.err | tee gpurun_out/bench_t3d92.json | cut -c1-3000 ) 2>&1 | tail -8
tail -5 gpurun_out/bench_t3d92.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/bench_t3d92.err | tee gpurun_out/bench_reference_t3d.json ) 2>&1 | tail -6
