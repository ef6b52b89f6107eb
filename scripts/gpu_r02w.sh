# 4 GPUs, T3D(48) per rank, weak scaling: Newton solve with the global hierarchy from level 2 (default) and from level 1
set -x
mkdir -p gpurun_out
for gl in 2 1; do
  B200_AMG_GLOBAL_LEVEL=$gl B200_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2966$gl \
    bench.py --gpus 4 --size 48 --steps 3 --warmup 3 --no-cpu 2>gpurun_out/r02w_n4_gl$gl.err | tail -1 > gpurun_out/r02w_bench_t3d48_n4_gl$gl.json
  grep -i "amg" gpurun_out/r02w_n4_gl$gl.err | sort | uniq -c | sort -rn | head -8
  python - <<PY
import json
d = json.loads(open("gpurun_out/r02w_bench_t3d48_n4_gl$gl.json").read().strip().splitlines()[-1])
ns = d.get("newton_step", {})
print("gl=$gl", {k: d.get(k) for k in ("value", "n_gpus")}, d.get("parity_check", {}).get("ok"), {k: ns.get(k) for k in ("converged", "gmres_iterations", "solve_ms", "solve_ms_per_iteration", "first_step_solve_ms", "ms")})
PY
done
