"""Forces the patch kernel (B200_GATHER_KERNEL=patch) on small problems and compares with the row-owner kernels."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from feng_b200 import mesh as M, problems as PB, capi
from feng_b200.linear_system import LinearSystemB200

def run(kern, dim, n, transient=False):
    os.environ["B200_GATHER_KERNEL"] = kern
    m = M.square_mesh(n) if dim == 2 else M.cube_mesh(n)
    pb = PB.taylor_hood(m, "ns_div", 8 if dim == 2 else 6, 1 if dim == 2 else 3, 1 / 40., 1., build_pattern=False, with_source=False)
    sol = PB.perturb_unknowns(pb)
    ls = LinearSystemB200(pb, device=0, device_pattern=True)
    S = ls.sys
    S.set_solution(sol)
    S.set_to_zero(3)
    S.assemble(3)
    v, r = S.get_matrix_values().copy(), S.get_rhs().copy()
    S.set_to_zero(3); S.assemble(1); r1 = S.get_rhs().copy()
    S.set_to_zero(3); S.assemble(2); v2 = S.get_matrix_values().copy()
    return v, r, r1, v2

for dim, n in ((2, 8), (2, 37), (3, 3), (3, 7)):
    try:
        a = run("patch", dim, n)
    except Exception as ex:
        print("patch failed:", dim, n, ex)
        continue
    b = run("node" if dim == 2 else "lane", dim, n)
    for name, x, y in zip(("val", "rhs", "rhs-only", "val-only"), a, b):
        print(dim, n, name, "max rel diff %.3e" % (np.abs(x - y).max() / np.abs(y).max()), "nan" if np.isnan(x).any() else "")
