# 8 GPUs, T3D(92) per rank (weak scaling, the default bench) with the new multigrid defaults (global hierarchy from level 1)
mkdir -p gpurun_out
B200_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29681 \
  bench.py --gpus 8 --steps 5 --warmup 3 2>gpurun_out/r02z_n8.err | tail -1 > gpurun_out/r02z_bench_t3d92_n8.json
grep -i "amg" gpurun_out/r02z_n8.err | sort | uniq -c | sort -rn | head -6
tail -3 gpurun_out/r02z_n8.err | cut -c1-300
python - <<PY
import json
d = json.loads(open("gpurun_out/r02z_bench_t3d92_n8.json").read().strip().splitlines()[-1])
ns = d.get("newton_step", {})
print({k: d.get(k) for k in ("value", "n_gpus")}, d["e2e"]["value"], d.get("parity_check", {}).get("ok"), {k: ns.get(k) for k in ("converged", "gmres_iterations", "solve_ms", "solve_ms_per_iteration", "first_step_solve_ms", "ms")})
PY
nvidia-smi --query-gpu=memory.used --format=csv | head -3
