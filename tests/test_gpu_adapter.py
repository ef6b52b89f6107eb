"""End-to-end drop-in check on the GPU: the UNMODIFIED reference host code (mesh reader, feSpace, feMetaNumber,
createTimeIntegrator -> solveNewtonRaphson, feNorm) drives the CUDA engine through the C++ adapter
adapter/feLinearSystemB200.h -- the one-line change of tests/withLinearSolver/navier_stokes.cpp:101-108 -- and must
print the reference's own goldens (6 significant digits, tests/tests.h:4).

Needs oracle/_ref/libfeng_ref_b200.so and oracle/_ref/data/*.msh, both built in the container by `make -C oracle`
(they travel to the GPU box with the snapshot; /root/reference itself is not needed at run time)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# tests/withLinearSolver/navier_stokes_MMS.output:4-7,11-14 (divergence form) and :19-22,26-29 (Laplacian form)
NS_DIV = {1: (1.730601e-03, 1.064397e-02), 2: (2.102148e-04, 2.316230e-03), 3: (2.593958e-05, 5.440602e-04),
          4: (3.221725e-06, 1.315003e-04)}
NS_LAP = {1: (1.729202e-03, 1.064192e-02), 2: (2.115442e-04, 2.276928e-03), 3: (2.645172e-05, 5.376581e-04)}
# tests/withLinearSolver/stokes_MMS.output:4,11 / :19,26
STOKES_DIV = {1: (1.730482e-03, 1.054679e-02)}
STOKES_LAP = {1: (1.729104e-03, 1.055463e-02)}
# tests/withLinearSolver/convergenceLaplace.output:10-13 (P2 Poisson, quadrature degree 12)
LAPLACE_P2 = {1: 3.209814e-03, 2: 4.066278e-04, 3: 5.111231e-05, 4: 6.408018e-06}


def _ref():
    from oracle import ref
    if not ref.available_b200():
        pytest.skip("oracle/_ref/libfeng_ref_b200.so not built (make -C oracle)")
    return ref


def _fmt(x):
    return f"{x:.6e}"


@pytest.mark.parametrize("kind,table", [("ns_div", NS_DIV), ("ns_lap", NS_LAP), ("stokes_div", STOKES_DIV),
                                        ("stokes_lap", STOKES_LAP)])
def test_taylor_hood_mms_goldens_through_the_adapter(kind, table):
    ref = _ref()
    for i, (eu, ep) in table.items():
        P = ref.RefProblem(os.path.join(ref.DATA_DIR, f"square{i}.msh"), kind, 2, 8, field=0, mu=1.0, rho=1.0,
                           p_essential=True, b200=True)
        # tight linear tolerance: the printed goldens come from a direct solve (Pardiso / MUMPS); everything else is the
        # reference's KSP default: GMRES(30), at most 1e4 iterations (src/feLinearSystem.h:66-69), preconditioner AUTO
        sol, info = P.newton_b200(1e-10, 1e-10, 10, rel_tol=1e-12)
        assert info["krylov_iterations"] <= 150 * info["n_solves"], info
        assert info["converged"], (kind, i, info)
        assert _fmt(info["errU"]) == _fmt(eu) and _fmt(info["errP"]) == _fmt(ep), (kind, i, info, eu, ep)
        P.close()


def test_poisson_p2_goldens_through_the_adapter():
    ref = _ref()
    for i, e in LAPLACE_P2.items():
        P = ref.RefProblem(os.path.join(ref.DATA_DIR, f"square{i}.msh"), "diffusion", 2, 12, field=0, mu=1.0, b200=True)
        sol, info = P.newton_b200(1e-10, 1e-10, 10, rel_tol=1e-12)
        assert info["krylov_iterations"] <= 40 * info["n_solves"], info
        assert info["converged"]
        assert _fmt(info["errU"]) == _fmt(e), (i, info, e)
        P.close()


def test_adapter_matches_cpu_stub_backend_on_a_synthetic_mesh(tmp_path):
    """same host objects, two backends: CPU stub (Eigen SparseLU) vs CUDA engine (device pattern, coloured scatter)"""
    ref = _ref()
    from feng_b200 import mesh as M
    path = str(tmp_path / "m.msh")
    M.write_msh(M.square_mesh(10), path)
    P = ref.RefProblem(path, "ns_div", 2, 8, field=0, mu=0.5, rho=1.3, p_essential=True, b200=True)
    s_cpu, _ = P.newton(1e-10, 1e-10, 10)
    for scatter, devpat in ((0, True), (1, False)):
        s_gpu, info = P.newton_b200(1e-10, 1e-10, 10, rel_tol=1e-12, scatter=scatter, device_pattern=devpat)
        assert info["converged"]
        assert np.abs(s_gpu - s_cpu).max() <= 1e-8 * np.abs(s_cpu).max(), info
    P.close()
