#!/bin/bash
# launch list of one Newton step (assembly + 40 GMRES iterations with the Schur/multigrid preconditioner) at T3D(92)
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02c_launches_newton_t3d92.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu --solve-maxit 40 > gpurun_out/r02c_bench_under_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r02c_launches_newton_t3d92.csv")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names = rows[hdr]; ki = names.index("Kernel Name"); vi = names.index("Metric Value")
# keep the last 2500 launches (the Newton step is at the end)
data = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[hdr + 2:] if len(r) > vi]
tail = data[-2200:]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in tail:
    k = k.split("(")[0][:70]
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{t/1e6:9.2f} ms {n:6d} x {t/n/1e3:9.1f} us  {100*t/tot:5.1f}%  {k}")
PY
