"""Preconditioners of the device GMRES: multigrid on scalar systems, Schur-complement +
multigrid on Taylor-Hood saddle-point systems.  Every solve is compared with a sparse direct solve of the SAME assembled
matrix (scipy) within the solver tolerance, with the reference's KSP defaults: GMRES(30), rtol 1e-8 on the preconditioned
residual, at most 1e4 iterations (src/feLinearSystem.h:66-69)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _direct(pb, ls):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    A = sp.csr_matrix((ls.sys.get_matrix_values(), pb.ja, pb.ia), shape=(pb.n_inc, pb.n_inc))
    r = ls.sys.get_rhs()
    return A, r, spla.spsolve(A.tocsc(), r)


def _assemble(pb, sol):
    from feng_b200.linear_system import LinearSystemB200
    ls = LinearSystemB200(pb)
    ls.sys.set_solution(sol)
    ls.sys.set_to_zero(3)
    ls.sys.assemble(3)
    return ls


@pytest.mark.parametrize("dim,n,order,max_its", [(2, 32, 2, 40), (2, 32, 1, 40), (3, 6, 2, 40), (3, 8, 1, 40)])
def test_amg_scalar_diffusion(dim, n, order, max_its):
    from feng_b200 import capi, mesh as M, problems as PB
    m = M.square_mesh(n) if dim == 2 else M.cube_mesh(n)
    pb = PB.scalar_diffusion(m, order, 6 if dim == 3 else 8)
    ls = _assemble(pb, pb.sol.copy())
    A, r, du_ref = _direct(pb, ls)
    info = ls.sys.solve(rel_tol=1e-8, restart=30, pc=capi.PC_AMG)
    assert info.converged == 1 and info.iterations <= max_its, (info.iterations, info.rel_residual)
    du = ls.sys.get_du()
    assert np.abs(du - du_ref).max() <= 1e-6 * np.abs(du_ref).max()
    # AUTO picks the same preconditioner for a scalar system
    info2 = ls.sys.solve(rel_tol=1e-8, restart=30, pc=capi.PC_AUTO)
    assert info2.iterations == info.iterations


@pytest.mark.parametrize("dim,n,kind,max_its", [(2, 32, "ns_div", 120), (2, 32, "ns_lap", 120), (2, 24, "stokes_div", 80),
                                                 (3, 6, "ns_div", 150), (3, 6, "ns_lap", 150)])
def test_schur_amg_taylor_hood(dim, n, kind, max_its):
    """T2D(32) / T3D(6) Navier-Stokes Jacobians at Re = 40 (the bench state), pinned pressure: GMRES(30) + Schur/AMG against
    scipy's direct solve; point-Jacobi needs thousands of iterations on the same systems."""
    from feng_b200 import capi, mesh as M, problems as PB
    m = M.square_mesh(n) if dim == 2 else M.cube_mesh(n)
    pb = PB.taylor_hood(m, kind, 8 if dim == 2 else 6, 1 if dim == 2 else 3, 1. / 40., 1.0, with_source=False)
    ls = _assemble(pb, PB.perturb_unknowns(pb))
    A, r, du_ref = _direct(pb, ls)
    info = ls.sys.solve(rel_tol=1e-8, restart=30, pc=capi.PC_SCHUR_AMG)
    assert info.converged == 1 and info.iterations <= max_its, (info.iterations, info.rel_residual)
    du = ls.sys.get_du()
    # rtol 1e-8 on the preconditioned residual: the true residual and the error follow within the conditioning
    assert np.abs(du - du_ref).max() <= 2e-5 * np.abs(du_ref).max()
    assert info.norm_axb <= 1e-5 * np.abs(r).max()
    # a second solve re-uses the hierarchy (same matrix): identical iteration count
    info2 = ls.sys.solve(rel_tol=1e-8, restart=30, pc=capi.PC_AUTO)
    assert info2.converged == 1 and abs(info2.iterations - info.iterations) <= 2


def test_schur_amg_pressure_dirichlet_boundary():
    """pressure essential on the whole boundary (no pinned mode): the rank-one term is off"""
    from feng_b200 import capi, mesh as M, problems as PB
    pb = PB.taylor_hood(M.square_mesh(16), "ns_div", 8, 0, 1.0, 1.0, p_essential=True)
    ls = _assemble(pb, PB.perturb_unknowns(pb, 1e-3))
    A, r, du_ref = _direct(pb, ls)
    info = ls.sys.solve(rel_tol=1e-10, restart=30, pc=capi.PC_SCHUR_AMG)
    assert info.converged == 1 and info.iterations <= 80, info.iterations
    assert np.abs(ls.sys.get_du() - du_ref).max() <= 1e-7 * np.abs(du_ref).max()


def test_restart_is_validated():
    from feng_b200 import capi, mesh as M, problems as PB
    pb = PB.scalar_diffusion(M.square_mesh(4), 1, 4)
    ls = _assemble(pb, pb.sol.copy())
    with pytest.raises(RuntimeError):
        ls.sys.solve(restart=100000, pc=capi.PC_JACOBI)


def test_block_jacobi_node_blocks_and_validation():
    """B200_PC_BLOCK_JACOBI: velocity node blocks (the dim components of one node) on a Stokes system with the pressure essential on
    the boundary; argument validation of b200_set_blocks; a block of pressure rows alone (zero P-P block) is reported singular."""
    from feng_b200 import capi, mesh as M, problems as PB
    pb = PB.taylor_hood(M.square_mesh(6), "stokes_div", 8, 0, 1.0, 1.0, p_essential=True)
    ls = _assemble(pb, pb.sol.copy())
    A, r, du_ref = _direct(pb, ls)
    nodes = pb.adrU.reshape(-1, pb.dim)
    nodes = np.unique(nodes[(nodes < pb.n_inc).all(1)], axis=0)              # both components unknown
    ptr = np.arange(nodes.shape[0] + 1, dtype=np.int64) * pb.dim
    ls.sys.set_blocks(ptr, nodes.reshape(-1).astype(np.int64))
    info = ls.sys.solve(rel_tol=1e-10, max_iter=20000, restart=100, pc=capi.PC_BLOCK_JACOBI)
    assert info.converged == 1
    assert np.abs(ls.sys.get_du() - du_ref).max() <= 1e-6 * np.abs(du_ref).max()
    info_j = ls.sys.solve(rel_tol=1e-10, max_iter=20000, restart=100, pc=capi.PC_JACOBI)
    assert info.iterations <= info_j.iterations
    with pytest.raises(RuntimeError):                                        # row out of range
        ls.sys.set_blocks(np.array([0, 2], np.int64), np.array([0, pb.n_inc], np.int64))
    with pytest.raises(RuntimeError):                                        # the same row twice
        ls.sys.set_blocks(np.array([0, 2], np.int64), np.array([3, 3], np.int64))
    prow = np.unique(pb.adrP[pb.adrP < pb.n_inc])[:2].astype(np.int64)
    ls.sys.set_blocks(np.array([0, 2], np.int64), prow)                      # pressure rows only: singular block
    with pytest.raises(RuntimeError):
        ls.sys.solve(rel_tol=1e-10, max_iter=10, restart=10, pc=capi.PC_BLOCK_JACOBI)
