"""TEST/SET-UP INFRASTRUCTURE -- regenerates feng_b200/data/quadrature.json from the compiled reference.

The symmetric triangle / tetrahedron rules are plain published data (nodes and weights); the reference hard-codes
them in src/feQuadratureTri.cpp and src/feQuadratureTet.cpp and selects them by polynomial degree
(src/feQuadrature.cpp).  The engine takes quadrature tables as INPUT; for the synthetic benchmarks, which run where
the reference is absent, the tables come from this JSON.  Run here (needs /root/reference + oracle/_ref):

    python oracle/gen_quadrature.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from feng_b200 import mesh as M  # noqa: E402
from oracle import ref  # noqa: E402

out = {"tri": {}, "tet": {}}
m2 = M.square_mesh(2)
M.write_msh(m2, "/tmp/_q2.msh")
m3 = M.cube_mesh(1)
M.write_msh(m3, "/tmp/_q3.msh")
for deg in range(1, 15):
    P = ref.RefProblem("/tmp/_q2.msh", "diffusion", order=2, quad_degree=deg)
    w, r, s, t = P.quadrature()
    out["tri"][str(deg)] = {"w": w.tolist(), "r": r.tolist(), "s": s.tolist()}
    P.close()
for deg in range(1, 11):
    P = ref.RefProblem("/tmp/_q3.msh", "diffusion", order=2, quad_degree=deg)
    w, r, s, t = P.quadrature()
    out["tet"][str(deg)] = {"w": w.tolist(), "r": r.tolist(), "s": s.tolist(), "t": t.tolist()}
    P.close()
path = os.path.join(os.path.dirname(__file__), "..", "feng_b200", "data", "quadrature.json")
with open(path, "w") as f:
    json.dump(out, f)
print({k: {d: len(v["w"]) for d, v in out[k].items()} for k in out})
