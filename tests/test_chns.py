"""Cahn-Hilliard Navier-Stokes weak form with finite-difference Jacobian.

CPU: the numpy restatement (oracle/chns_oracle.py) against the compiled reference (live, where oracle/_ref exists) and
against the committed fixture generated from it; the host-side problem builder (numbering, element->DOF tables, CSR
pattern of four fields) bit-exactly against the same.  GPU: the CUDA kernel through the C ABI against the oracle.

Tolerances: residual 1e-12 relative to max|rhs| (north_star).  The Jacobian is a FORWARD DIFFERENCE with
delta = sqrt(eps) max(|u|, 1) (src/feBilinearForm.cpp:170, :404-422): rounding differences of 1e-16 |R| between two
correct evaluations of the residual are amplified by 1/delta = 6.7e7, so two implementations of the same formula agree
to ~1e-8 of the row scale, not 1e-12.  The bound is MEASURED on the reference itself: tests/golden/make_golden_chns.py
recomputes the reference's Jacobian from a state moved by one unit in the last place in every unknown and stores the
largest change relative to the row maximum in the fixture (`fd_noise`: 2.9e-8 CHNS_Abels P1, 3.2e-8 P2, 7.1e-8
CHNS_MassAveraged, 5.2e-8 CHNS_Khanwale).  Matrix tolerance: 4 x that noise against the fixtures, 3e-7 (4 x the largest
measured value) for the synthetic cases.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, assert_close_rows, assert_close_vec

FD_TOL = 3e-7          # 4 x the largest finite-difference noise measured on the reference (see the module docstring)


def fixture_fd_tol(g):
    return 4.0 * float(g["fd_noise"]) if "fd_noise" in g else FD_TOL
KHANWALE = (40., 25., 0.07, 3.0, 1.7, 1.3, 0.7)      # Re, Pe, Cn, We, Fr, rhoA, rhoB
DT = 0.02


def _variant(formulation):
    """extra ChnsModel / ChnsParams fields of a formulation"""
    return dict(formulation=formulation, alpha=-0.3 if formulation == "mass_averaged" else 0.0,
                khanwale=KHANWALE if formulation == "khanwale" else (1.,) * 7)


MODELS = {
    "plain": dict(rhoA=1.3, rhoB=0.7, viscA=0.05, viscB=0.02, mobility=0.01, sigma=0.5, epsilon=0.07),
    "full": dict(rhoA=1.3, rhoB=0.7, viscA=0.05, viscB=0.02, mobility=0.01, sigma=0.5, epsilon=0.07, force=(0.1, -0.98),
                 src_u=(0.2, 0.3), src_p=0.1, src_phi=-0.2, src_mu=0.4, limiter=True, degenerate_mobility=True),
}


def _oracle_problem(pb):
    from oracle import chns_oracle as CO
    m = pb.chns
    prm = CO.ChnsParams(**{k: getattr(m, k) for k in CO.ChnsParams.__dataclass_fields__})
    return CO.ChnsProblem(pb.dim, pb.mesh.xyz, pb.mesh.cells, [pb.adrU, pb.adrP, pb.adrF, pb.adrM], pb.w,
                          [pb.LU, pb.LP, pb.LF, pb.LF], [pb.dLU, pb.dLP, pb.dLF, pb.dLF], pb.n_inc, prm)


def _params_from_fixture(cls, a):
    """ChnsParams / ChnsModel from the fixture's parameter block (oracle/chns_oracle.ChnsParams.as_array)"""
    kw = dict(force=tuple(a[7:9]), src_u=tuple(a[9:11]), src_p=a[11], src_phi=a[12], src_mu=a[13], limiter=bool(a[14]),
              degenerate_mobility=bool(a[15]), phi_order=int(a[16]))
    if len(a) > 17 and a[17] == 1.0:
        kw.update(formulation="mass_averaged", alpha=float(a[18]))
    if len(a) > 17 and a[17] == 2.0:
        kw.update(formulation="khanwale", khanwale=tuple(float(x) for x in a[19:26]))
    return cls(*a[:7], **kw)


def _state(pb, seed=5):
    from feng_b200 import problems as PB
    sol = PB.perturb_unknowns(pb, 1e-2, seed=seed)
    sd = np.random.default_rng(seed + 1).standard_normal(pb.n_dof)
    return sol, sd, 3.5


@pytest.mark.parametrize("model,phi_order,formulation", [("plain", 1, "abels"), ("full", 1, "abels"), ("full", 2, "abels"),
                                                         ("plain", 1, "mass_averaged"), ("full", 2, "mass_averaged"),
                                                         ("plain", 1, "khanwale"), ("full", 2, "khanwale")])
def test_oracle_and_host_tables_vs_compiled_reference(model, phi_order, formulation, tmp_path, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from feng_b200 import mesh as M, problems as PB
    from oracle import chns_oracle as CO, ref
    m = M.rect_mesh(5, 4, 1.3, 0.9, -0.2, 0.1)
    path = str(tmp_path / "m.msh")
    M.write_msh(m, path)
    mdl = PB.ChnsModel(phi_order=phi_order, **_variant(formulation), **MODELS[model])
    pb = PB.chns(m, mdl, 8, 1, 0.05, 1.3)
    opb = _oracle_problem(pb)
    P = ref.RefProblem(path, "chns", 2, 8, 1, 0.05, 1.3, chns=opb.prm.as_array())
    # numbering, element->DOF tables, pattern: bit-exact
    assert (P.n_dof, P.n_inc) == (pb.n_dof, pb.n_inc)
    for s, a in enumerate([pb.adrU, pb.adrP, pb.adrF, pb.adrM]):
        assert np.array_equal(P.adr(s), a)
    ia, ja = P.pattern()
    assert np.array_equal(ia, pb.ia) and np.array_equal(ja, pb.ja)
    rsol, _ = P.solution()
    assert np.abs(rsol - pb.sol).max() <= 1e-14
    sol, sd, c0 = _state(pb)
    P.set_solution(sol, sd, c0, 0.0)
    sol_n = None
    if formulation != "abels":     # state at the previous time step (the reference's global solAtTimeN) != current state
        sol_n = sol + np.random.default_rng(9).uniform(-0.05, 0.05, sol.shape)
        P.set_solution_n(sol_n, DT)
    fi = P.form_info(0)
    assert fi.sys_id == {"abels": CO.CHNS_ABELS, "mass_averaged": CO.CHNS_MASS_AVERAGED, "khanwale": CO.CHNS_KHANWALE}[formulation]
    assert fi.M == opb.adr[0].shape[1] + 3 + 2 * pb.LF.shape[1]
    v, r, _ = P.assemble()
    ov, orr = CO.assemble(opb, pb.ia, pb.ja, sol, sd, c0, sol_n=sol_n, dt=DT)
    assert_close_vec(orr, r, 1e-13, "rhs")
    assert_close_rows(ov, v, pb.ia, FD_TOL, "FD matrix")
    P.close()


@pytest.mark.parametrize("name", ["ref_square1_chns_abels_p1", "ref_square1_chns_abels_p2",
                                  "ref_square1_chns_mass_averaged_p1", "ref_square1_chns_khanwale_p1"])
def test_oracle_vs_golden_fixture(name):
    """Fixture = inputs as the reference tabulated them + outputs of its own CPU path (tests/golden/make_golden.py)."""
    from oracle import chns_oracle as CO
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    prm = _params_from_fixture(CO.ChnsParams, g["chns_params"])
    sol_n = g["sol_n"] if "sol_n" in g else None
    dt = float(g["dt"]) if "dt" in g else 0.0
    LU = np.ascontiguousarray(g["L0"][:, 0::2, 0])
    dLU = np.ascontiguousarray(np.stack([g["dLdr0"][:, 0::2, 0], g["dLds0"][:, 0::2, 0]], 2))
    Ls, dLs = [LU], [dLU]
    for s in (1, 2, 3):
        Ls.append(g[f"L{s}"])
        dLs.append(np.ascontiguousarray(np.stack([g[f"dLdr{s}"], g[f"dLds{s}"]], 2)))
    pb = CO.ChnsProblem(2, g["xyz"], g["cells"], [g[f"adr{s}"] for s in range(4)], g["w"], Ls, dLs, int(g["n_inc"]), prm)
    Ae, Be, adr = CO.element_systems(pb, g["sol"], g["sol_dot"], float(g["c0"]), sol_n=sol_n, dt=dt)
    for i, e in enumerate(g["elements"]):
        assert np.abs(Be[e] - g["Be"][i]).max() <= 1e-13 * np.abs(g["Be"][i]).max()
        assert np.abs(Ae[e] - g["Ae"][i]).max() <= fixture_fd_tol(g) * np.abs(g["Ae"][i]).max()
    ov, orr = CO.assemble(pb, g["ia"], g["ja"], g["sol"], g["sol_dot"], float(g["c0"]), sol_n=sol_n, dt=dt)
    assert_close_vec(orr, g["rhs"], 1e-13, "rhs")
    assert_close_rows(ov, g["values"], g["ia"], fixture_fd_tol(g), "FD matrix")


def test_fd_jacobian_is_the_derivative_of_the_residual():
    """the FD matrix of the restatement against central differences of its own assembled residual"""
    from feng_b200 import mesh as M, problems as PB
    from oracle import chns_oracle as CO
    m = M.square_mesh(2)
    pb = PB.chns(m, PB.ChnsModel(**MODELS["full"]), 8, 1, 0.05, 1.3)
    opb = _oracle_problem(pb)
    sol, _, _ = _state(pb)
    v, _ = CO.assemble(opb, pb.ia, pb.ja, sol)
    import scipy.sparse as sp
    A = sp.csr_matrix((v, pb.ja, pb.ia), shape=(pb.n_inc, pb.n_inc)).toarray()
    h = 1e-6
    for j in np.random.default_rng(1).choice(pb.n_inc, 10, replace=False):
        sp_, sm_ = sol.copy(), sol.copy()
        sp_[j] += h
        sm_[j] -= h
        _, bp = CO.assemble(opb, pb.ia, pb.ja, sp_, matrix=False)
        _, bm = CO.assemble(opb, pb.ia, pb.ja, sm_, matrix=False)
        fd = -(bp - bm) / (2 * h)
        assert np.abs(fd - A[:, j]).max() <= 1e-5 * max(1.0, np.abs(A[:, j]).max())


# ---- GPU -------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("model,phi_order,device_pattern,formulation",
                         [("plain", 1, False, "abels"), ("full", 1, True, "abels"), ("full", 2, False, "abels"),
                          ("plain", 1, True, "mass_averaged"), ("full", 1, False, "mass_averaged"),
                          ("full", 2, False, "mass_averaged"), ("plain", 1, False, "khanwale"),
                          ("full", 1, True, "khanwale"), ("full", 2, False, "khanwale")])
def test_cuda_chns_vs_oracle(model, phi_order, device_pattern, formulation):
    from feng_b200 import mesh as M, problems as PB
    from feng_b200.linear_system import LinearSystemB200
    from oracle import chns_oracle as CO
    m = M.square_mesh(10)
    mdl = PB.ChnsModel(phi_order=phi_order, **_variant(formulation), **MODELS[model])
    pb = PB.chns(m, mdl, 8, 1, 0.05, 1.3)
    sol, sd, c0 = _state(pb)
    sol_n = None if formulation == "abels" else sol + np.random.default_rng(9).uniform(-0.05, 0.05, sol.shape)
    ov, orr = CO.assemble(_oracle_problem(pb), pb.ia, pb.ja, sol, sd, c0, sol_n=sol_n, dt=DT)
    ls = LinearSystemB200(pb, device_pattern=device_pattern)
    if sol_n is not None:
        ls.sys.set_solution_n(sol_n, DT)
    if device_pattern:
        ia, ja = ls.sys.get_pattern()
        assert np.array_equal(ia, pb.ia) and np.array_equal(ja, pb.ja)
    ls.sys.set_solution(sol, sd, c0, 0.0)
    ls.sys.set_to_zero(3)
    ls.sys.assemble(3, False)
    assert_close_vec(ls.sys.get_rhs(), orr, 1e-12, "rhs")
    assert_close_rows(ls.sys.get_matrix_values(), ov, pb.ia, FD_TOL, "FD matrix")
    # residual-only pass gives the same rhs; the matrix is left alone
    ls.sys.set_to_zero(1)
    ls.sys.assemble(1, False)
    assert_close_vec(ls.sys.get_rhs(), orr, 1e-12, "rhs (residual-only)")
    assert_close_rows(ls.sys.get_matrix_values(), ov, pb.ia, FD_TOL, "matrix untouched")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ref_square1_chns_abels_p1", "ref_square1_chns_mass_averaged_p1",
                                  "ref_square1_chns_khanwale_p1"])
def test_cuda_chns_vs_golden_fixture(name):
    """the CUDA path on the reference's OWN tables (fixture) against the reference's own outputs"""
    from feng_b200 import capi
    from feng_b200 import problems as PB
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    mdl = _params_from_fixture(PB.ChnsModel, g["chns_params"])
    S = capi.System(0)
    S.set_mesh(2, g["xyz"], g["cells"])
    S.set_quadrature(g["w"])
    LU = np.ascontiguousarray(g["L0"][:, 0::2, 0])
    dLU = np.ascontiguousarray(np.stack([g["dLdr0"][:, 0::2, 0], g["dLds0"][:, 0::2, 0]], 2))
    ids = [S.add_space(6, 2, g["adr0"], LU, dLU)]
    for s in (1, 2, 3):
        ids.append(S.add_space(g[f"L{s}"].shape[1], 1, g[f"adr{s}"], g[f"L{s}"],
                               np.ascontiguousarray(np.stack([g[f"dLdr{s}"], g[f"dLds{s}"]], 2))))
    S.set_pattern(int(g["n_inc"]), int(g["n_dof"]), g["ia"], g["ja"])
    S.add_form_chns(*ids, mdl)
    S.finalize()
    S.set_solution(g["sol"], g["sol_dot"], float(g["c0"]), 0.0)
    if "sol_n" in g:
        S.set_solution_n(g["sol_n"], float(g["dt"]))
    S.set_to_zero(3)
    S.assemble(3, False)
    assert_close_vec(S.get_rhs(), g["rhs"], 1e-12, "rhs")
    assert_close_rows(S.get_matrix_values(), g["values"], g["ia"], fixture_fd_tol(g), "FD matrix")


@pytest.mark.gpu
@pytest.mark.parametrize("phi_order,device_pattern,formulation", [(1, False, "abels"), (2, True, "abels"),
                                                                  (1, True, "mass_averaged"), (2, False, "mass_averaged"),
                                                                  (1, False, "khanwale"), (2, True, "khanwale")])
def test_chns_through_the_cpp_adapter(phi_order, device_pattern, formulation):
    """The reference's own host objects (mesh reader, feSpace, feMetaNumber, CHNS_Abels<2> / CHNS_MassAveraged<2> with
    their property CALLBACKS, the global solAtTimeN)
    drive the CUDA backend through adapter/feLinearSystemB200.h: the adapter probes the callbacks, recognises the laws
    of CHNS_Solver (src/CHNS_Solver.cpp:124-235) and registers the monolithic form; the result must match the
    reference's own CPU assembly of the same state (both computed here, in the same process)."""
    from oracle import chns_oracle as CO, ref
    if not ref.available_b200():
        pytest.skip("oracle/_ref/libfeng_ref_b200.so not built (make -C oracle)")
    prm = CO.ChnsParams(phi_order=phi_order, **_variant(formulation), **MODELS["full"])
    P = ref.RefProblem(os.path.join(ref.DATA_DIR, "square2.msh"), "chns", 2, 8, 1, 0.05, 1.3, b200=True,
                       chns=prm.as_array())
    sol, _ = P.solution()
    rng = np.random.default_rng(3)
    sol[:P.n_inc] += rng.uniform(-1e-2, 1e-2, P.n_inc)
    sd = rng.standard_normal(P.n_dof)
    P.set_solution(sol, sd, 2.5, 0.0)
    if formulation != "abels":
        P.set_solution_n(sol + rng.uniform(-5e-2, 5e-2, P.n_dof), DT)
    v, r, _ = P.assemble()                        # reference CPU path (colour loop + FD Jacobian + scatter)
    gv, gr = P.assemble_b200(device_pattern=device_pattern)
    ia, _ = P.pattern()
    assert_close_vec(gr, r, 1e-12, "rhs")
    assert_close_rows(gv, v, ia, FD_TOL, "FD matrix")
    P.close()
