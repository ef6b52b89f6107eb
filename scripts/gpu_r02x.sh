# 2 GPUs, T3D(92) per rank (weak scaling, the default bench): global hierarchy from level 1
set -x
mkdir -p gpurun_out
B200_AMG_GLOBAL_LEVEL=1 B200_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 \
  bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu 2>gpurun_out/r02x_n2.err | tail -1 > gpurun_out/r02x_bench_t3d92_n2_gl1.json
grep -i "amg" gpurun_out/r02x_n2.err | sort | uniq -c | sort -rn | head -8
python - <<PY
import json
d = json.loads(open("gpurun_out/r02x_bench_t3d92_n2_gl1.json").read().strip().splitlines()[-1])
ns = d.get("newton_step", {})
print({k: d.get(k) for k in ("value", "n_gpus")}, d.get("parity_check", {}).get("ok"), {k: ns.get(k) for k in ("converged", "gmres_iterations", "solve_ms", "solve_ms_per_iteration", "first_step_solve_ms", "ms")})
PY
nvidia-smi --query-gpu=memory.used --format=csv | head -3
