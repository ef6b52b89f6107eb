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
