#!/bin/bash
# Counters of ONE assembly pass at the bench size (T3D(92)): DRAM bytes and executed FP64 instructions of every assembly launch,
# summed into profiles/traffic.json (read by bench.py for roofline.traffic and roofline.fp64), plus `ncu --set full` captures of the
# dominant kernels.  Run on the GPU box:  gpurun -- bash scripts/gpu_r02e.sh
set -u
M="dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,gpu__time_duration.sum"
# warm-up 3 passes + 1 timed pass + the per-launch passes: keep the LAST pass of the assembly kernels
ncu --metrics $M --clock-control none -k regex:'gather_lane_kernel|element_state_kernel' --csv \
    --log-file gpurun_out/r02e_asm_counters_t3d92.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-solve --no-parity \
    > gpurun_out/r02e_bench_under_ncu.log 2>&1
python - <<'PY'
import csv, json, subprocess, collections
rows = list(csv.reader(open("gpurun_out/r02e_asm_counters_t3d92.csv")))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names = rows[h]
idi, ki, mi, vi = names.index("ID"), names.index("Kernel Name"), names.index("Metric Name"), names.index("Metric Value")
launch = collections.OrderedDict()
for r in rows[h + 2:]:
    if len(r) <= vi:
        continue
    d = launch.setdefault(r[idi], {"kernel": r[ki]})
    d[r[mi]] = float(r[vi].replace(",", ""))
L = list(launch.values())
# one pass = 1 element_state launch followed by the gather_lane launches up to the next element_state launch
starts = [i for i, d in enumerate(L) if "element_state" in d["kernel"]]
last = L[starts[-1]:]
tot = collections.Counter()
for d in last:
    for k, v in d.items():
        if k != "kernel":
            tot[k] += v
nE = 6 * 92 ** 3
flop = 2 * tot["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"] + tot["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"] + tot["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
out = {"t3d": (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / nE, "t3d_flop_per_element": flop / nE,
       "t3d_launches_per_pass": len(last), "t3d_kernel_ns_under_ncu": tot["gpu__time_duration.sum"],
       "t3d_source": "ncu counters of the last assembly pass of `bench.py --steps 1 --warmup 3` at T3D(92): scripts/gpu_r02e.sh, "
                     "profiles/r02e_asm_counters_t3d92.csv"}
try:
    old = json.load(open("profiles/traffic.json"))
except Exception:
    old = {}
old.update(out)
json.dump(old, open("gpurun_out/traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
PY
# full captures: the main velocity-row launch of the last pass, the outer SpMV and the single-precision multigrid SpMV
ncu --set full --clock-control none --import-source on -k regex:'gather_lane_kernel' --launch-skip 60 --launch-count 3 \
    -o gpurun_out/r02e_full_lane3d_t3d92 -f python bench.py --steps 1 --warmup 3 --no-cpu --no-solve --no-parity > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'amg_spmv_f32_kernel|spmv_rpg_kernel|pc_rhs_tail' --launch-skip 40 --launch-count 6 \
    -o gpurun_out/r02e_full_solver_t3d92 -f python bench.py --steps 1 --warmup 3 --no-cpu --no-parity --solve-maxit 20 > /dev/null 2>&1
for f in r02e_full_lane3d_t3d92 r02e_full_solver_t3d92; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.csv 2>/dev/null
done
ls -la gpurun_out/ | tail -12
