import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def to_oracle_problem(pb):
    """HostProblem (product-side tables) -> oracle.fe_oracle.Problem (checker-side container)."""
    from oracle import fe_oracle as O
    forms = [O.Form(f.kind, f.coeff, f.param, f.source) for f in pb.forms]
    return O.Problem(pb.dim, pb.mesh.xyz, pb.mesh.cells, pb.adrU, pb.adrP, pb.ncomp, pb.w, pb.LU, pb.dLU, pb.LP,
                     pb.n_inc, forms)


def assert_close_rows(a, b, ia, rtol=1e-12, what=""):
    """|a - b| <= rtol * max|b| over the entry's matrix row (entries that are sums cancelling far below their terms
    carry an absolute error proportional to the terms, SURVEY.md section 7 'hard parts')."""
    rowmax = np.maximum.reduceat(np.abs(b), ia[:-1])
    scale = np.repeat(rowmax, np.diff(ia))
    err = np.abs(a - b)
    bad = err > rtol * np.maximum(scale, 1e-300)
    assert not bad.any(), f"{what}: {bad.sum()} entries differ, worst rel-to-row {np.max(err / np.maximum(scale, 1e-300)):.3e}"


def assert_close_vec(a, b, rtol=1e-12, what=""):
    scale = np.abs(b).max()
    err = np.abs(a - b).max()
    assert err <= rtol * max(scale, 1e-300), f"{what}: max err {err:.3e} vs scale {scale:.3e}"


@pytest.fixture(scope="session")
def have_ref():
    from oracle import ref
    return ref.available()


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(prefix=""):
    """fixtures of tests/golden/make_golden.py (the CHNS fixtures of make_golden_chns.py have their own tests, test_chns.py)"""
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith(prefix) and "chns" not in f)


def load_golden(name):
    """Fixture written by tests/golden/make_golden.py from the compiled, unmodified reference."""
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))


def golden_to_oracle_problem(g):
    """oracle.fe_oracle.Problem on the reference's OWN tables (element->DOF maps, basis and quadrature tables as the
    reference tabulated them), with the form list of the harness recipe (oracle/ref_harness.cpp ref_create)."""
    from oracle import fe_oracle as O
    dim = int(g["dim"])
    vector = g["L0"].ndim == 3
    if vector:   # vector Lagrange tables: function a*dim+c is phi_a e_c (src/feSpace_2D.cpp:41-55)
        LU = np.ascontiguousarray(g["L0"][:, 0::dim, 0])
        d = [g["dLdr0"][:, 0::dim, 0], g["dLds0"][:, 0::dim, 0]] + ([g["dLdt0"][:, 0::dim, 0]] if dim == 3 else [])
    else:
        LU = g["L0"]
        d = [g["dLdr0"], g["dLds0"]] + ([g["dLdt0"]] if dim == 3 else [])
    dLU = np.ascontiguousarray(np.stack(d, 2))
    LP = g["L1"] if "L1" in g else None
    adrP = g["adr1"] if "adr1" in g else None
    mu, rho, field = float(g["mu"]), float(g["rho"]), int(g["field"])
    # quadrature points in physical space: x = sum_v L1_v(xi_k) x_v (P1 geometry)
    pts = np.stack([g["qr"], g["qs"]] + ([g["qt"]] if dim == 3 else []), 1)
    lam = np.concatenate([1.0 - pts.sum(1, keepdims=True), pts], 1)
    xq = np.einsum("kv,evm->ekm", lam, g["xyz"][g["cells"]])
    from feng_b200 import problems as PB          # numpy twins of the harness callbacks (host-side set-up code)
    forms = []
    for f, (M, N, has_mat, sys_id, transient) in enumerate(g["form_info"]):
        k = int(sys_id)
        kind = str(g["kind"])
        if k == O.VECTOR_CONVECTIVE_ACCELERATION:
            forms.append(O.Form(k, -rho))
        elif k == O.MIXED_DIVERGENCE:
            forms.append(O.Form(k, 1.0))
        elif k == O.VECTOR_SOURCE:
            forms.append(O.Form(k, 1.0, 0.0, PB.u_source(field, xq[..., :dim], mu, rho, kind.startswith("ns"))))
        elif k == O.DIV_NEWTONIAN_STRESS:
            forms.append(O.Form(k, 1.0, mu))
        elif k == O.VECTOR_DIFFUSION:
            forms.append(O.Form(k, -1.0, mu))
        elif k == O.MIXED_GRADIENT:
            forms.append(O.Form(k, -1.0))
        elif k == O.TRANSIENT_VECTOR_MASS:
            forms.append(O.Form(k, -rho))
        elif k == O.DIFFUSION:
            forms.append(O.Form(k, 1.0, mu))
        elif k == O.SOURCE:
            forms.append(O.Form(k, 1.0, 0.0, PB.s_source(field, xq[..., :dim], mu)))
        elif k == O.TRANSIENT_MASS:
            forms.append(O.Form(k, rho))
        else:
            raise ValueError(k)
    return O.Problem(dim, g["xyz"], g["cells"], g["adr0"], adrP, dim if vector else 1, g["w"], LU, dLU, LP,
                     int(g["n_inc"]), forms)
