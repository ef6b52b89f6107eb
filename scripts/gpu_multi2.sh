set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_multi2.log
B200_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/bench_n2.err | tee gpurun_out/bench_t3d92_n2.json | cut -c1-300
grep 'row-lane' gpurun_out/bench_n2.err | head -6
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_t3d92_n2.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus")}, d.get("parity_check"), d.get("newton_step", {}).get("converged"), d.get("newton_step", {}).get("iterations"))
PY
