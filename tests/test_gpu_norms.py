"""feNorm on the device: Lp error norms and H1 seminorms of the device-resident state against the same
quadrature sums in numpy (src/feNorm.cpp:323-398, :1399-1443, :1643-1732), for scalar and vector spaces in 2-D and 3-D, and the
convergence of the nodal interpolant of a smooth field (rates 3 / 2 for P2 in L2 / H1)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _numpy_norms(pb, adr, L, dL, ncomp, sol, exact, grad_exact, p):
    m = pb.mesh
    X = m.xyz[m.cells][:, :, :pb.dim]                      # (nE, nv, dim)
    F = np.transpose(X[:, 1:] - X[:, :1], (0, 2, 1))       # F[e][m][al]
    J = np.linalg.det(F)
    G = np.linalg.inv(F)                                   # G[e][al][m]
    nE, nq, nS = adr.shape[0], L.shape[0], L.shape[1]
    U = sol[adr].reshape(nE, nS, ncomp)
    uh = np.einsum("kb,ebc->ekc", L, U)
    lp = (np.abs(exact - uh) ** p).sum(2)
    lp = ((lp * pb.w[None, :]).sum(1) * J).sum() ** (1. / p)
    gr = np.einsum("kba,ebc->ekca", dL, U)                 # reference gradient
    gp = np.einsum("eam,ekca->ekcm", G, gr)
    h1 = (((gp - grad_exact) ** 2).sum((2, 3)) * pb.w[None, :]).sum(1)
    return lp, np.sqrt((h1 * J).sum())


@pytest.mark.parametrize("dim,n", [(2, 7), (3, 3)])
@pytest.mark.parametrize("p", [1, 2, 4])
def test_lp_and_h1_norms_match_numpy(dim, n, p):
    from feng_b200 import mesh as M, problems as PB
    from feng_b200.linear_system import LinearSystemB200
    m = M.square_mesh(n) if dim == 2 else M.unstructured_tet_mesh(n, seed=5)
    pb = PB.taylor_hood(m, "ns_div", 8 if dim == 2 else 6, 0 if dim == 2 else 3, 0.05, 1.1)
    sol = PB.perturb_unknowns(pb)
    ls = LinearSystemB200(pb)
    ls.sys.set_solution(sol, None, 0.0, 0.0)
    rng = np.random.default_rng(17)
    nq = pb.w.size
    for adr, Lt, dLt, nc, sid in ((pb.adrU, pb.LU, pb.dLU, dim, ls.su), (pb.adrP, pb.LP, pb.dLP, 1, ls.sp)):
        Lk = np.asarray(Lt).reshape(nq, -1)
        dLk = np.asarray(dLt).reshape(nq, Lk.shape[1], dim)
        ex = rng.standard_normal((m.n_cells, nq, nc))
        gex = rng.standard_normal((m.n_cells, nq, nc, dim))
        lp, h1 = _numpy_norms(pb, np.asarray(adr), Lk, dLk, nc, sol, ex, gex, p)
        assert abs(ls.sys.error_norm(sid, 0, p, ex) - lp) <= 1e-12 * lp
        assert abs(ls.sys.error_norm(sid, 1, 2, gex) - h1) <= 1e-12 * h1
        lp0, h10 = _numpy_norms(pb, np.asarray(adr), Lk, dLk, nc, sol, 0 * ex, 0 * gex, p)
        assert abs(ls.sys.error_norm(sid, 0, p, None) - lp0) <= 1e-12 * lp0
        assert abs(ls.sys.error_norm(sid, 1, 2, None) - h10) <= 1e-12 * h10
        # fixed summation order: bitwise repeatable
        assert ls.sys.error_norm(sid, 0, p, ex) == ls.sys.error_norm(sid, 0, p, ex)


def test_interpolation_error_converges_at_the_p2_rates():
    """L2 / H1 error of the P2 nodal interpolant of sin(pi x) cos(pi y): rates 3 and 2 (the figures feNorm produces for the MMS
    goldens of the reference, tests/withLinearSolver/*.output)."""
    from feng_b200 import mesh as M, problems as PB
    from feng_b200.linear_system import LinearSystemB200
    errs = []
    for n in (8, 16, 32):
        m = M.square_mesh(n)
        pb = PB.scalar_diffusion(m, 2, 8, 0, 1.0)
        xq = PB.quad_points_physical(m, pb.qpts)[..., :2]
        f = lambda x: np.sin(np.pi * x[..., 0]) * np.cos(np.pi * x[..., 1])
        ex = f(xq)[..., None]
        gex = np.stack([np.pi * np.cos(np.pi * xq[..., 0]) * np.cos(np.pi * xq[..., 1]),
                        -np.pi * np.sin(np.pi * xq[..., 0]) * np.sin(np.pi * xq[..., 1])], -1)[:, :, None, :]
        xd = PB.dof_coordinates(m, pb.num, "U", pb.n_dof)
        sol = f(xd[:, :2])
        ls = LinearSystemB200(pb)
        ls.sys.set_solution(sol, None, 0.0, 0.0)
        errs.append((ls.sys.error_norm(ls.su, 0, 2, ex), ls.sys.error_norm(ls.su, 1, 2, gex)))
    r_l2 = np.log2(errs[1][0] / errs[2][0])
    r_h1 = np.log2(errs[1][1] / errs[2][1])
    assert 2.9 < r_l2 < 3.1 and 1.9 < r_h1 < 2.1, (errs, r_l2, r_h1)
