"""CPU oracle (oracle/fe_oracle.py) pinned against golden fixtures generated from the compiled, unmodified reference
(tests/golden/make_golden.py): per-element matrices and residuals of every weak form, Jacobian determinants, the
assembled CSR values and rhs, the essential-component constraint, and the Newton solution whose error norms are the
numbers printed in the reference's own goldens (tests/withLinearSolver/navier_stokes_MMS.output:4,11,19,26,
stokes_MMS.output:4,11,19,26, convergenceLaplace.output:10)."""
import numpy as np
import pytest

from conftest import assert_close_rows, assert_close_vec, golden_names, golden_to_oracle_problem, load_golden

ALL = golden_names()


@pytest.mark.parametrize("name", ALL)
def test_geometry_matches_reference_jacobians(name):
    from oracle import fe_oracle as O
    g = load_golden(name)
    geo = O.geometry(g["xyz"], g["cells"], int(g["dim"]))
    # feCncGeo::_J[nq*e + k] is constant over k on straight simplices (src/feCncGeo.cpp:278-418)
    assert np.abs(geo.detJ[:, None] - g["detJ"]).max() <= 1e-15 * np.abs(g["detJ"]).max()


@pytest.mark.parametrize("name", ALL)
def test_element_matrices_and_residuals_per_form(name):
    """Ae / Be of each form on the sampled elements, as feBilinearForm::computeMatrix/computeResidual gave them."""
    from oracle import fe_oracle as O
    g = load_golden(name)
    pb = golden_to_oracle_problem(g)
    dim = pb.dim
    geo_all = O.geometry(pb.xyz, pb.cells, dim)
    el = g["elements"]
    geo = O.Geometry(geo_all.detJ[el], geo_all.G[el])
    sol = g["sol"]
    sd = g["sol_dot"] if "sol_dot" in g else None
    nS = pb.LU.shape[1]
    uloc = sol[pb.adrU[el]].reshape(len(el), nS, pb.ncomp)
    udot = None if sd is None else sd[pb.adrU[el]].reshape(len(el), nS, pb.ncomp)
    ploc = None if pb.adrP is None else sol[pb.adrP[el]]
    for f, form in enumerate(pb.forms):
        if form.source is not None:
            form = O.Form(form.kind, form.coeff, form.param, form.source[el])
        Ae, Be = O.element_forms(form, dim, geo, pb.w, pb.LU, pb.dLU, pb.LP, uloc, ploc, udot, float(g["c0"]))
        gA, gB = g[f"Ae{f}"], g[f"Be{f}"]
        if Ae is not None:
            assert np.abs(Ae - gA).max() <= 1e-13 * max(np.abs(gA).max(), 1e-300), (name, f)
        assert np.abs(Be - gB).max() <= 1e-13 * max(np.abs(gB).max(), 1e-300), (name, f)
        # element -> DOF maps of the form (adrI x adrJ layout of createElementarySystem)
        rows, cols = O.form_layout(form.kind)
        adr = {"U": pb.adrU, "P": pb.adrP}
        assert np.array_equal(np.concatenate([adr[r][el] for r in rows], 1), g[f"adrI{f}"])
        assert np.array_equal(np.concatenate([adr[c][el] for c in cols], 1), g[f"adrJ{f}"])


@pytest.mark.parametrize("name", ALL)
def test_assembled_system(name):
    from oracle import fe_oracle as O
    g = load_golden(name)
    pb = golden_to_oracle_problem(g)
    sd = g["sol_dot"] if "sol_dot" in g else None
    v, r = O.assemble(pb, g["ia"], g["ja"], g["sol"], sd, float(g["c0"]))
    assert_close_rows(v, g["vals"], g["ia"], 1e-13, name + " matrix")
    assert_close_vec(r, g["rhs"], 1e-13, name + " rhs")


@pytest.mark.parametrize("name", [n for n in ALL if "ns_" in n or "stokes" in n])
def test_constraint_semantics(name):
    """constrainEssentialComponents as restated from src/feLinearSystemMklPardiso.cpp:1092-1114."""
    from oracle import fe_oracle as O
    g = load_golden(name)
    v, r = O.constrain(g["ia"], g["ja"], g["vals"], g["rhs"], g["constraint_rows"])
    assert np.array_equal(v, g["vals_constrained"]) and np.array_equal(r, g["rhs_constrained"])


@pytest.mark.parametrize("name", [n for n in ALL if "newton_sol" in load_golden(n)])
def test_newton_solution_and_printed_error_norms(name):
    """Host Newton loop (feng_b200.linear_system.solve_newton_raphson, the mirror of solveNewtonRaphson) driving a
    CPU stand-in backend built from the oracle + a sparse direct solve reproduces the reference's Newton solution."""
    from cpu_backend import OracleLinearSystem
    from feng_b200.linear_system import NLSolverOptions, solve_newton_raphson
    g = load_golden(name)
    pb = golden_to_oracle_problem(g)
    ls = OracleLinearSystem(pb, g["ia"], g["ja"], g["constraint_rows"])
    sol = g["sol_init"].copy()
    status, hist = solve_newton_raphson(ls, sol, NLSolverOptions(1e-10, 1e-10, 1e4, 10, 3, 1e-1))
    assert status == 0
    assert 1 <= len(hist) <= int(g["newton_info"][7])
    assert np.abs(sol - g["newton_sol"]).max() <= 1e-10 * np.abs(g["newton_sol"]).max()
