# memcheck of the assembly paths (row-owner node/lane kernels, patch kernel, CHNS kernels, pattern build, GMRES)
set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_patch.py tests/test_chns.py -m gpu -x -q -k "not adapter" 2>&1 | tail -15 | tee gpurun_out/sanitizer_memcheck.log
