# r02z: full GPU suite + default bench line with the new multigrid defaults (MIS(1) aggregates, W-cycle below the finest level,
# over-correction 1.5), launch list of a Newton step, and the aspect-ratio check of the Krylov iteration count (one GPU, undecomposed)
T=${1:-r02z}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu.log 2>&1
tail -4 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py 2>gpurun_out/${T}_bench.err | tail -1 > gpurun_out/${T}_bench_t3d92.json
python - <<PY
import json
d = json.loads(open("gpurun_out/${T}_bench_t3d92.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus")}, d["roofline"]["frac"], d["e2e"]["value"], d.get("parity_check"))
print(d.get("newton_step"))
PY
for cfg in "--n 40" "--n 20 --box 8" "--n 25 --box 4" "--n 32 --box 2"; do
  echo "== pc_scale $cfg"; python scripts/pc_scale.py $cfg 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['n'], d['n_inc'], d['solve1']['its'], round(d['solve1']['solve_ms'], 1), d['solve1']['converged'])"
done 2>&1 | tee gpurun_out/${T}_aspect_ratio.log
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/${T}_launches_newton_t3d92.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-parity --solve-maxit 12 > /dev/null 2>&1
wc -l gpurun_out/${T}_launches_newton_t3d92.csv
