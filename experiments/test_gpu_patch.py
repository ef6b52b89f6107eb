"""GPU parity of the opt-in patch kernel (feng_b200/csrc/patch.cu, B200_GATHER_KERNEL=patch): block-slot owners over
Morton patches of elements, against the CPU oracle and against the fixtures assembled by the unmodified reference.
Same tolerance as tests/test_gpu_parity.py (1e-12 relative to the largest entry of the row / of the rhs)."""
import numpy as np
import pytest

from conftest import assert_close_rows, assert_close_vec, to_oracle_problem

pytestmark = pytest.mark.gpu


@pytest.fixture
def patch_kernel(monkeypatch):
    monkeypatch.setenv("B200_GATHER_KERNEL", "patch")


def _assemble(pb, sol, sol_dot=None, c0=0.0, patch_elems=None, monkeypatch=None):
    from feng_b200.linear_system import LinearSystemB200
    if patch_elems is not None:
        monkeypatch.setenv("B200_PATCH_ELEMS", str(patch_elems))
    ls = LinearSystemB200(pb)
    assert ls.sys.gather_plan_kind() == 2, "the patch plan was not built"
    ls.sys.set_solution(sol, sol_dot, c0, 0.0)
    ls.sys.set_to_zero(3)
    ls.sys.assemble(3)
    return ls, ls.sys.get_matrix_values(), ls.sys.get_rhs()


@pytest.mark.parametrize("kind", ["ns_div", "ns_lap", "stokes_div", "stokes_lap"])
@pytest.mark.parametrize("patch_elems", [7, 64])
def test_patch_2d_vs_oracle(kind, patch_elems, patch_kernel, monkeypatch):
    from feng_b200 import mesh as M, problems as PB
    from oracle import fe_oracle as O
    m = M.square_mesh(12)
    pb = PB.taylor_hood(m, kind, 8, 0, 0.025, 1.3)
    sol = PB.perturb_unknowns(pb)
    ov, orr = O.assemble(to_oracle_problem(pb), pb.ia, pb.ja, sol)
    _, v, r = _assemble(pb, sol, patch_elems=patch_elems, monkeypatch=monkeypatch)
    assert_close_rows(v, ov, pb.ia, 1e-12, "matrix")
    assert_close_vec(r, orr, 1e-12, "rhs")


def test_patch_2d_transient_split_passes_and_determinism(patch_kernel):
    from feng_b200 import mesh as M, problems as PB
    from oracle import fe_oracle as O
    m = M.square_mesh(9)
    pb = PB.taylor_hood(m, "ns_div", 8, 0, 0.1, 1.3, transient=True, p_essential=True)
    sol = PB.perturb_unknowns(pb)
    sd = np.random.default_rng(3).standard_normal(pb.n_dof)
    ov, orr = O.assemble(to_oracle_problem(pb), pb.ia, pb.ja, sol, sd, 3.5)
    ls, v, r = _assemble(pb, sol, sd, 3.5)
    assert_close_rows(v, ov, pb.ia, 1e-12, "matrix")
    assert_close_vec(r, orr, 1e-12, "rhs")
    ls.sys.set_to_zero(3)
    ls.sys.assemble(2, False)
    ls.sys.assemble(1, False)
    assert np.array_equal(ls.sys.get_matrix_values(), v) and np.array_equal(ls.sys.get_rhs(), r)   # bitwise
    # transient-only matrix = mass form alone (assembleOnlyTransientMatrices)
    opb = to_oracle_problem(pb)
    opb.forms = [f for f in opb.forms if f.kind == O.TRANSIENT_VECTOR_MASS]
    mv, _ = O.assemble(opb, pb.ia, pb.ja, sol, sd, 3.5, residual=False)
    ls.sys.set_to_zero(3)
    ls.sys.assemble(2, True)
    assert np.abs(ls.sys.get_matrix_values() - mv).max() <= 1e-12 * np.abs(mv).max()


@pytest.mark.parametrize("kind", ["ns_div", "ns_lap"])
def test_patch_3d_vs_oracle(kind, patch_kernel):
    from feng_b200 import mesh as M, problems as PB
    from oracle import fe_oracle as O
    m = M.cube_mesh(3)
    pb = PB.taylor_hood(m, kind, 6, 3, 0.05, 1.1)
    sol = PB.perturb_unknowns(pb)
    ov, orr = O.assemble(to_oracle_problem(pb), pb.ia, pb.ja, sol)
    _, v, r = _assemble(pb, sol)
    assert_close_rows(v, ov, pb.ia, 1e-12, "matrix")
    assert_close_vec(r, orr, 1e-12, "rhs")


@pytest.mark.parametrize("name", ["ref_square1_ns_div", "ref_square1_ns_lap", "ref_square1_ns_div_transient"])
def test_patch_against_reference_fixtures(name, patch_kernel):
    """Unstructured data/square1.msh with the reference's own numbering, tables and pattern: values the unmodified
    reference assembled (tests/golden/*.npz)."""
    from conftest import golden_to_oracle_problem, load_golden
    from feng_b200 import capi
    from feng_b200.problems import form_layout
    g = load_golden(name)
    opb = golden_to_oracle_problem(g)
    S = capi.System(0)
    S.set_mesh(opb.dim, opb.xyz, opb.cells)
    S.set_quadrature(opb.w)
    su = S.add_space(opb.LU.shape[1], opb.ncomp, opb.adrU, opb.LU, opb.dLU)
    sp = S.add_space(opb.LP.shape[1], 1, opb.adrP, opb.LP, np.zeros(opb.LP.shape + (opb.dim,)))
    S.set_pattern(int(g["n_inc"]), int(g["n_dof"]), g["ia"], g["ja"])
    for f in opb.forms:
        rows, cols = form_layout(f.kind)
        if rows == ("P",):
            S.add_form(f.kind, sp, su, f.coeff, f.param, f.source)
        else:
            S.add_form(f.kind, su, sp if "P" in cols else -1, f.coeff, f.param, f.source)
    S.finalize()
    assert S.gather_plan_kind() == 2, "the patch plan was not built"
    sd = g["sol_dot"] if "sol_dot" in g else None
    S.set_solution(g["sol"], sd, float(g["c0"]), 0.0)
    S.set_to_zero(3)
    S.assemble(3)
    assert_close_rows(S.get_matrix_values(), g["vals"], g["ia"], 1e-12, "matrix vs reference")
    assert_close_vec(S.get_rhs(), g["rhs"], 1e-12, "rhs vs reference")


@pytest.mark.parametrize("kind", ["ns_div", "ns_lap", "stokes_div"])
@pytest.mark.parametrize("patch_elems", [9, 64])
def test_row_slice_kernel_2d_vs_oracle(kind, patch_elems, patch_kernel, monkeypatch):
    """B200_PATCH_KERNEL=slice (csrc/slice.cuh): element-centric producers with constant-bank reference tensors, slot owners
    that keep their sums in registers across the row slices; same plan, same numbers."""
    from feng_b200 import mesh as M, problems as PB
    from oracle import fe_oracle as O
    monkeypatch.setenv("B200_PATCH_KERNEL", "slice")
    m = M.square_mesh(13)
    pb = PB.taylor_hood(m, kind, 8, 0, 0.025, 1.3, with_source=False)
    sol = PB.perturb_unknowns(pb)
    ov, orr = O.assemble(to_oracle_problem(pb), pb.ia, pb.ja, sol)
    ls, v, r = _assemble(pb, sol, patch_elems=patch_elems, monkeypatch=monkeypatch)
    assert_close_rows(v, ov, pb.ia, 1e-12, "matrix")
    assert_close_vec(r, orr, 1e-12, "rhs")
    ls.sys.set_to_zero(3)
    ls.sys.assemble(2, False)                       # matrix-only pass: same numbers, rhs untouched (lazy zero -> zeros)
    assert np.array_equal(ls.sys.get_matrix_values(), v)
