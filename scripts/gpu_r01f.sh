set -x
mkdir -p gpurun_out
export B200_VERBOSE=1
timeout 600 python scripts/patch_debug.py 2>&1 | tail -20
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>gpurun_out/bench_patch.err | tee gpurun_out/bench_t2d_patch.json | cut -c1-200
for PE in 32 48 96 128; do
B200_PATCH_ELEMS=$PE timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>>gpurun_out/bench_patch.err | tee gpurun_out/bench_t2d_patch_pe$PE.json | cut -c1-200
done
for PE in 16 32 48; do
B200_PATCH_ELEMS=$PE timeout 600 python bench.py --workload t3d --steps 5 --warmup 3 --no-cpu 2>>gpurun_out/bench_patch.err | tee gpurun_out/bench_t3d_patch_pe$PE.json | cut -c1-200
done
B200_PATCH_THREADS=512 B200_PATCH_ELEMS=32 timeout 600 python bench.py --workload t3d --steps 5 --warmup 3 --no-cpu 2>>gpurun_out/bench_patch.err | tee gpurun_out/bench_t3d_patch_pe32_t512.json | cut -c1-200
grep feng_b200 gpurun_out/bench_patch.err | tail
