# r02z: full GPU suite + default bench line with the new multigrid defaults (MIS(1) aggregates, W-cycle below the finest level, over-correction 1.75)
T=${1:-r02z}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1
tail -5 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py 2>gpurun_out/${T}_bench.err | tail -1 > gpurun_out/${T}_bench_t3d92.json
python - <<PY
import json
d = json.loads(open("gpurun_out/${T}_bench_t3d92.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus")}, d["roofline"]["frac"], d["e2e"]["value"], d.get("parity_check"))
print(d.get("newton_step"))
print(d.get("cpu_baseline"))
PY
