set -x
timeout 600 python scripts/slice_debug.py 2>&1 | tail -14
B200_GATHER_KERNEL=patch B200_PATCH_KERNEL=slice timeout 600 python bench.py --workload t2d --steps 10 --warmup 3 --no-cpu --no-solve 2>/dev/null | cut -c1-200
B200_GATHER_KERNEL=patch B200_PATCH_KERNEL=slice B200_PATCH_ELEMS=48 timeout 600 python bench.py --workload t2d --steps 10 --warmup 3 --no-cpu --no-solve 2>/dev/null | cut -c1-200
B200_GATHER_KERNEL=patch B200_PATCH_KERNEL=slice B200_PATCH_ELEMS=32 timeout 600 python bench.py --workload t2d --steps 10 --warmup 3 --no-cpu --no-solve 2>/dev/null | cut -c1-200
