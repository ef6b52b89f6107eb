"""Live cross-checks of the numpy oracle against the compiled reference (oracle/_ref; skipped where it is absent, e.g.
on the GPU box -- the committed fixtures cover that case) plus reference-free consistency checks of the dim=3 vector
forms, which the reference does not implement (src/feVectorSysElm.cpp:1246,1536 instantiate <2> only)."""
import numpy as np
import pytest

from conftest import assert_close_rows, assert_close_vec, to_oracle_problem


@pytest.mark.parametrize("kind,transient", [("ns_div", False), ("ns_lap", True), ("stokes_div", True),
                                            ("stokes_lap", False)])
def test_taylor_hood_2d_all_elements(kind, transient, tmp_path, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from feng_b200 import mesh as M, problems as PB
    from oracle import fe_oracle as O, ref
    m = M.rect_mesh(7, 5, 1.3, 0.9, -0.2, 0.1)
    path = str(tmp_path / "m.msh")
    M.write_msh(m, path)
    P = ref.RefProblem(path, kind, 2, 8, field=0, mu=0.07, rho=1.7, transient=transient, p_essential=True)
    pb = PB.taylor_hood(m, kind, 8, 0, 0.07, 1.7, transient=transient, p_essential=True)
    sol = PB.perturb_unknowns(pb, 5e-2, seed=11)
    sd = np.random.default_rng(1).standard_normal(pb.n_dof) if transient else None
    c0 = 1.7 if transient else 0.0
    P.set_solution(sol, sd, c0, 0.0)
    v, r, _ = P.assemble()
    opb = to_oracle_problem(pb)
    ov, orr = O.assemble(opb, pb.ia, pb.ja, sol, sd, c0)
    assert_close_rows(ov, v, pb.ia, 1e-13, "matrix")
    assert_close_vec(orr, r, 1e-13, "rhs")
    from feng_b200.coloring import color_elements
    assert np.array_equal(P.colors(), color_elements(m.cells, m.n_vertices))     # feCncGeo::colorElements(1)
    P.close()


@pytest.mark.parametrize("dim,order,deg", [(2, 2, 12), (2, 1, 4), (3, 2, 4), (3, 2, 6), (3, 1, 2)])
def test_scalar_diffusion_all_elements(dim, order, deg, tmp_path, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from feng_b200 import mesh as M, problems as PB
    from oracle import fe_oracle as O, ref
    m = M.square_mesh(6) if dim == 2 else M.cube_mesh(3)
    path = str(tmp_path / "m.msh")
    M.write_msh(m, path)
    P = ref.RefProblem(path, "diffusion", order, deg, field=0, mu=0.7, rho=1.3, transient=True)
    pb = PB.scalar_diffusion(m, order, deg, 0, 0.7, transient=True, rho=1.3)
    assert np.array_equal(P.adr(0), pb.adrU)
    ia, ja = P.pattern()
    assert np.array_equal(ia, pb.ia) and np.array_equal(ja, pb.ja)
    sol = PB.perturb_unknowns(pb, 5e-2, seed=5)
    sd = np.random.default_rng(2).standard_normal(pb.n_dof)
    P.set_solution(sol, sd, 2.5, 0.0)
    v, r, _ = P.assemble()
    ov, orr = O.assemble(to_oracle_problem(pb), pb.ia, pb.ja, sol, sd, 2.5)
    assert_close_rows(ov, v, pb.ia, 1e-13, "matrix")
    assert_close_vec(orr, r, 1e-13, "rhs")
    P.close()


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("kind", ["ns_div", "ns_lap"])
def test_jacobian_is_the_derivative_of_the_residual(dim, kind):
    """compareAnalyticalAndFDMatrices logic (src/feBilinearForm.cpp:430-475) on the oracle itself: the dim=3 vector
    forms (restatement only) must satisfy Ae = -dBe/du like their dim=2 twins, which are pinned on the reference."""
    from feng_b200 import mesh as M, problems as PB
    from oracle import fe_oracle as O
    m = M.square_mesh(2) if dim == 2 else M.cube_mesh(1)
    pb = PB.taylor_hood(m, kind, 8 if dim == 2 else 6, 0 if dim == 2 else 3, 0.3, 1.2, p_essential=False)
    opb = to_oracle_problem(pb)
    n = pb.n_dof
    ia = np.arange(0, n * n + 1, n, dtype=np.int64)           # dense pattern over ALL DOFs (n_inc := n_dof)
    ja = np.tile(np.arange(n, dtype=np.int32), n)
    opb.n_inc = n
    sol = pb.sol + np.random.default_rng(0).uniform(-0.1, 0.1, n)
    A, b0 = O.assemble(opb, ia, ja, sol)
    A = A.reshape(n, n)
    h = 1e-6
    for j in np.random.default_rng(1).choice(n, 12, replace=False):
        sp, sm = sol.copy(), sol.copy()
        sp[j] += h
        sm[j] -= h
        _, bp = O.assemble(opb, ia, ja, sp, matrix=False)
        _, bm = O.assemble(opb, ia, ja, sm, matrix=False)
        fd = -(bp - bm) / (2 * h)                              # Be = -(residual)  =>  Ae = -dBe/du
        assert np.abs(fd - A[:, j]).max() <= 1e-7 * max(1.0, np.abs(A[:, j]).max())


def test_3d_vector_forms_reduce_to_reference_scalar_blocks(tmp_path, have_ref):
    """The U-U block of VECTOR_DIFFUSION in 3-D is diag(K, K, K) with K the scalar P2 diffusion matrix, which the
    reference does implement on tetrahedra (feSysElm_Diffusion<3>, src/feSysElm.cpp:586-589)."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from feng_b200 import mesh as M, problems as PB, tables as T
    from oracle import fe_oracle as O, ref
    m = M.cube_mesh(2)
    path = str(tmp_path / "m.msh")
    M.write_msh(m, path)
    P = ref.RefProblem(path, "diffusion", 2, 6, field=0, mu=0.37)
    w, q = T.quadrature(3, 6)
    LU, dLU = T.basis(3, 2, q)
    geo = O.geometry(m.xyz, m.cells, 3)
    nE = m.n_cells
    uloc = np.random.default_rng(4).standard_normal((nE, 10, 3))
    Ae, Be = O.element_forms(O.Form(O.VECTOR_DIFFUSION, 1.0, 0.37), 3, geo, w, LU, dLU, None, uloc)
    for e in (0, 7, nE - 1):
        K, _, _, _ = P.element(0, e)
        for c in range(3):
            assert np.abs(Ae[e][c::3, c::3] - K).max() <= 1e-13 * np.abs(K).max()
            for c2 in range(3):
                if c2 != c:
                    assert np.all(Ae[e][c::3, c2::3] == 0.0)
    P.close()
