import sys, json, numpy as np
sys.path.insert(0, "/root/repo")
from feng_b200 import mesh as M, problems as PB
from feng_b200.linear_system import LinearSystemB200
n = int(sys.argv[1])
for L in [int(x) for x in sys.argv[2:]]:
    m = M.box_mesh(n, n, n * L, float(L)); m.point_pressure = 0
    pb = PB.taylor_hood(m, "ns_div", 6, 3, 1/40., 1.0, with_source=False, build_pattern=False)
    sol = PB.perturb_unknowns(pb)
    ls = LinearSystemB200(pb, device_pattern=True); S = ls.sys
    S.set_solution(sol); S.set_to_zero(3); S.assemble(3); S.constrain()
    info = S.solve(1e-8, 1e-14, 1e6, 3000, 30, 6, raise_on_fail=False)
    print("box 1x1x%d at n=%d: n_inc %d its %d conv %d" % (L, n, pb.n_inc, info.iterations, info.converged), flush=True)
    S.close()
