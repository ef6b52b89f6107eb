set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chns.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_chns.log
