"""Device-resident time stepping: BDF time derivative, history shift and essential-BC refresh on the device
against the host path that re-uploads the whole state (feSolutionContainer, src/feSolutionContainer.cpp:82-105, :338-348;
feSolution::initializeEssentialBC, src/feTimeIntegration.cpp:523-536)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,n", [(2, 6), (3, 2)])
def test_bdf2_state_on_the_device_matches_full_uploads(dim, n):
    from feng_b200 import mesh as M, problems as PB
    from feng_b200.linear_system import LinearSystemB200
    m = M.square_mesh(n) if dim == 2 else M.cube_mesh(n)
    pb = PB.taylor_hood(m, "ns_div", 8 if dim == 2 else 6, 1 if dim == 2 else 3, 0.05, 1.2, transient=True, with_source=False)
    rng = np.random.default_rng(4)
    u = [pb.sol + rng.uniform(-0.05, 0.05, pb.n_dof) for _ in range(3)]          # u_{n-1}, u_n, u_{n+1}: essential DOFs differ too
    dt = 0.01
    c = np.array([1.5 / dt, -2.0 / dt, 0.5 / dt])                                # BDF2
    ls = LinearSystemB200(pb)
    S = ls.sys
    # host path: the whole state and its time derivative cross the bus
    dot = c[0] * u[2] + c[1] * u[1] + c[2] * u[0]
    S.set_solution(u[2], dot, c[0], 3 * dt)
    S.set_to_zero(3)
    S.assemble(3)
    v_ref, r_ref = S.get_matrix_values(), S.get_rhs()
    # device path: history kept on the GPU, only the essential values of the new level are uploaded
    S.set_solution(u[0])
    S.state_push()
    S.set_solution(u[1])
    S.state_push()
    ess = np.arange(pb.n_inc, pb.n_dof, dtype=np.int64)
    start = u[1].copy()
    start[:pb.n_inc] = u[2][:pb.n_inc]          # unknowns of the new level (what correctSolution leaves on the device) ...
    S.set_solution(start)
    S.set_essential(ess, np.ascontiguousarray(u[2][ess]))                       # ... and its essential values, sparsely
    assert np.array_equal(S.get_solution(), u[2])
    S.state_bdf(c, 3 * dt, dt)
    S.set_to_zero(3)
    S.assemble(3)
    v, r = S.get_matrix_values(), S.get_rhs()
    assert np.abs(v - v_ref).max() <= 1e-13 * np.abs(v_ref).max()
    assert np.abs(r - r_ref).max() <= 1e-12 * np.abs(r_ref).max()
    # a second refresh with the same index array re-uses the device copy of the indices
    S.set_essential(ess, np.ascontiguousarray(u[1][ess]))
    assert np.array_equal(S.get_solution()[ess], u[1][ess])
    S.close()


def test_newton_loop_without_state_uploads():
    """Stationary Newton with the state resident on the device: one upload, then only corrections (b200_correct_solution)."""
    from feng_b200 import mesh as M, problems as PB
    from feng_b200.linear_system import LinearSystemB200, NLSolverOptions, solve_newton_raphson
    pb = PB.taylor_hood(M.square_mesh(8), "ns_div", 8, 0, 1.0, 1.0, p_essential=True)
    sols = []
    for resident in (False, True):
        ls = LinearSystemB200(pb)
        ls.device_resident = resident
        sol = pb.sol.copy()
        sol[:pb.n_inc] = 0.0
        status, hist = solve_newton_raphson(ls, sol, NLSolverOptions(1e-10, 1e-10, 1e4, 10, 4, 1e-1))
        assert status == 0
        sols.append((sol, ls.uploads))
    # same Newton iterates up to the reduction order of the Krylov dot products (atomics)
    assert np.abs(sols[0][0] - sols[1][0]).max() <= 1e-9 * np.abs(sols[0][0]).max()
    assert sols[1][1] == 1 and sols[0][1] > 1
