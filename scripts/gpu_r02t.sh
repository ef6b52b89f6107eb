#!/bin/bash
# r02t: evidence for the row-lane velocity kernels at the bench size (T3D(92)):
#  1. full GPU test suite
#  2. counters of EVERY assembly launch of one pass -> profiles/traffic.json (bench.py: roofline.traffic, roofline.fp64)
#  3. ncu --set full of the pre-pass and the velocity-row launches (+ source page of the mid-edge launch)
#  4. launch list (gpu__time_duration) of the default bench command
#  5. the default bench line (assembly + Newton step + CPU arm), un-profiled
T=${1:-r02t}
set -u
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${T}_pytest_gpu.log
M="dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,gpu__time_duration.sum"
ncu --metrics $M --clock-control none --kernel-name-base demangled -k regex:'gather_lane_kernel|gather_urow_kernel|element_state' --csv \
    --log-file gpurun_out/${T}_asm_counters_t3d92.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-solve --no-parity \
    > gpurun_out/${T}_bench_under_ncu.log 2>&1
python - <<PY
import csv, json, collections
rows = list(csv.reader(open("gpurun_out/${T}_asm_counters_t3d92.csv")))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names = rows[h]
idi, ki, mi, vi = names.index("ID"), names.index("Kernel Name"), names.index("Metric Name"), names.index("Metric Value")
launch = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    d = launch.setdefault(r[idi], {"kernel": r[ki]})
    try:
        d[r[mi]] = float(r[vi].replace(",", ""))
    except ValueError:
        pass
L = list(launch.values())
starts = [i for i, d in enumerate(L) if "element_state" in d["kernel"]]
last = L[starts[-1]:]
tot = collections.Counter()
for d in last:
    print(d["kernel"][:60], "%.3f ms" % (d["gpu__time_duration.sum"] / 1e6), "dram %.2f + %.2f GB" % (d["dram__bytes_read.sum"] / 1e9, d["dram__bytes_write.sum"] / 1e9))
    for k, v in d.items():
        if k != "kernel":
            tot[k] += v
nE = 6 * 92 ** 3
flop = 2 * tot["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"] + tot["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"] + tot["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
out = {"t3d": (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / nE, "t3d_flop_per_element": flop / nE,
       "t3d_launches_per_pass": len(last), "t3d_kernel_ns_under_ncu": tot["gpu__time_duration.sum"],
       "t3d_source": "ncu counters of the last assembly pass of \`bench.py --steps 1 --warmup 3\` at T3D(92): scripts/gpu_r02t.sh, "
                     "profiles/r02t_asm_counters_t3d92.csv"}
try:
    old = json.load(open("profiles/traffic.json"))
except Exception:
    old = {}
old.update(out)
json.dump(old, open("gpurun_out/traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
PY
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'gather_urow_kernel|element_state_urow' --launch-skip 9 --launch-count 3 \
    -o gpurun_out/${T}_full_asm_t3d92 -f python bench.py --steps 1 --warmup 3 --no-cpu --no-solve --no-parity > /dev/null 2>&1
ncu -i gpurun_out/${T}_full_asm_t3d92.ncu-rep --page raw --csv > gpurun_out/${T}_full_asm_t3d92.csv 2>/dev/null
ncu -i gpurun_out/${T}_full_asm_t3d92.ncu-rep --page source --csv --print-source sass > gpurun_out/${T}_asm_source_sass.csv 2>/dev/null
rm -f gpurun_out/${T}_full_asm_t3d92.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 600 --csv --log-file gpurun_out/${T}_launches_bench_t3d92.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --solve-maxit 10 > gpurun_out/${T}_bench_under_ncu2.log 2>&1
timeout 900 python bench.py > gpurun_out/${T}_bench_t3d92.json 2> gpurun_out/${T}_bench.err
cut -c1-400 gpurun_out/${T}_bench_t3d92.json; tail -2 gpurun_out/${T}_bench.err
ls -la gpurun_out/${T}_*
