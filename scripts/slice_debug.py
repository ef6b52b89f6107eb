"""Row-slice kernel (B200_GATHER_KERNEL=patch B200_PATCH_KERNEL=slice) against the row-owner kernels on small 2-D problems."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from feng_b200 import mesh as M, problems as PB
from feng_b200.linear_system import LinearSystemB200

def run(env, n, kind):
    for k in ("B200_GATHER_KERNEL", "B200_PATCH_KERNEL"):
        os.environ.pop(k, None)
    os.environ.update(env)
    pb = PB.taylor_hood(M.square_mesh(n), kind, 8, 1, 1 / 40., 1.3, build_pattern=False, with_source=False)
    sol = PB.perturb_unknowns(pb)
    S = LinearSystemB200(pb, device=0, device_pattern=True).sys
    S.set_solution(sol); S.set_to_zero(3); S.assemble(3)
    v, r = S.get_matrix_values().copy(), S.get_rhs().copy()
    S.set_to_zero(3); S.assemble(2); v2 = S.get_matrix_values().copy()
    return v, r, v2

for n, kind in ((8, "ns_div"), (37, "ns_div"), (21, "ns_lap"), (16, "stokes_div")):
    a = run({"B200_GATHER_KERNEL": "patch", "B200_PATCH_KERNEL": "slice"}, n, kind)
    b = run({"B200_GATHER_KERNEL": "node"}, n, kind)
    for name, x, y in zip(("val", "rhs", "val-only"), a, b):
        print(n, kind, name, "max rel diff %.3e" % (np.abs(x - y).max() / np.abs(y).max()), "nan" if np.isnan(x).any() else "")
