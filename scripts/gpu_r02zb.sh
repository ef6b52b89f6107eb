# r02zb: Schur scaling with the stress form counted twice: solver-level GPU tests + the default bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_precond.py tests/test_gpu_configs.py tests/test_gpu_adapter.py tests/test_gpu_stepping.py -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r02zb_pytest_solver.log
timeout 600 python bench.py 2>gpurun_out/r02zb_bench.err | tail -1 > gpurun_out/r02zb_bench_t3d92.json
python - <<PY
import json
d = json.loads(open("gpurun_out/r02zb_bench_t3d92.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step")}, d["roofline"]["frac"], d["e2e"]["value"], d.get("parity_check", {}).get("ok"))
ns = d["newton_step"]; print({k: ns.get(k) for k in ("ms", "gmres_iterations", "converged", "solve_ms", "solve_ms_per_iteration", "first_step_ms")})
PY
