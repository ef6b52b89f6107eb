#!/bin/bash
# r02j: urowical-orientation velocity rows (gather_urow.cuh): parity at test and bench size, bench line, launch list of one pass
T=${1:-r02j}
B200_VERBOSE=1 timeout 300 python - > gpurun_out/${T}_plan.log 2>&1 <<'PY'
import numpy as np
from feng_b200 import mesh as M, problems as PB
from feng_b200.linear_system import LinearSystemB200
from oracle import fe_oracle as O
m = M.cube_mesh(4)
pb = PB.taylor_hood(m, "ns_div", 6, 3, 1 / 40., 1.0)
sol = PB.perturb_unknowns(pb)
ls = LinearSystemB200(pb)
ls.sys.set_solution(sol); ls.sys.set_to_zero(3); ls.sys.assemble(3)
v, r = ls.sys.get_matrix_values(), ls.sys.get_rhs()
forms = [O.Form(f.kind, f.coeff, f.param, f.source) for f in pb.forms]
opb = O.Problem(pb.dim, m.xyz, m.cells, pb.adrU, pb.adrP, pb.ncomp, pb.w, pb.LU, pb.dLU, pb.LP, pb.n_inc, forms)
ov, orr = O.assemble(opb, pb.ia, pb.ja, sol)
print("T3D(4) urow: matrix rel err %.3e rhs rel err %.3e" % (np.abs(v - ov).max() / np.abs(ov).max(), np.abs(r - orr).max() / np.abs(orr).max()))
PY
grep -v 'MiB' gpurun_out/${T}_plan.log | tail -5; grep "row-lane plan" gpurun_out/${T}_plan.log
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_stepping.py -q -x -k "3d or t3d or stepping" > gpurun_out/${T}_pytest.log 2>&1
tail -4 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-solve > gpurun_out/${T}_bench_t3d92.json 2> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_bench_t3d92.json; tail -3 gpurun_out/${T}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,launch__occupancy_limit_shared_mem,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum \
  --clock-control none --kernel-name-base demangled -k regex:'gather_|element_state' --launch-skip 40 --launch-count 20 --csv --log-file gpurun_out/${T}_launch_metrics_t3d92.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-solve --no-parity > /dev/null 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/${T}_launch_metrics_t3d92.csv")))
st = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[st]; d = collections.OrderedDict()
for r in rows[st + 1:]:
    if len(r) < len(hdr): continue
    rec = dict(zip(hdr, r)); d.setdefault((rec["ID"], rec["Kernel Name"][:60], rec["Grid Size"]), {})[rec["Metric Name"]] = rec["Metric Value"]
for k, v in d.items():
    print(k, {m.split("__")[-1][:34]: x for m, x in v.items()})
PY
