"""The C++/OpenMP CPU arm of the 3-D workload (oracle/port_cpp.cpp) pinned on the numpy oracle in 2-D and 3-D and, in 2-D, on the
compiled unmodified reference (oracle/_ref): pattern bit-identical, values and rhs to 1e-13."""
import os

import numpy as np
import pytest


def _port():
    from oracle import port
    if not port.available():
        pytest.skip("oracle/_ref/libfeng_port.so not built (make -C oracle)")
    return port


@pytest.mark.parametrize("dim,n,kind", [(2, 5, "ns_div"), (2, 5, "ns_lap"), (3, 3, "ns_div"), (3, 3, "ns_lap"), (3, 2, "stokes_div")])
def test_port_matches_numpy_oracle(dim, n, kind):
    port = _port()
    from conftest import to_oracle_problem
    from feng_b200 import coloring, mesh as M, problems as PB
    from oracle import fe_oracle as O
    m = M.square_mesh(n) if dim == 2 else M.cube_mesh(n)
    pb = PB.taylor_hood(m, kind, 8 if dim == 2 else 6, 1 if dim == 2 else 3, 1. / 40., 1.3, with_source=False)
    sol = PB.perturb_unknowns(pb)
    P = port.PortProblem(pb, coloring.color_elements(m.cells, m.n_vertices))
    assert np.array_equal(P.ia, pb.ia) and np.array_equal(P.ja, pb.ja)          # EZCRS pattern, bit-identical
    port.set_threads(4)
    v, r, sec = P.assemble(sol)
    ov, orr = O.assemble(to_oracle_problem(pb), pb.ia, pb.ja, sol)
    assert np.abs(v - ov).max() <= 1e-13 * np.abs(ov).max()
    assert np.abs(r - orr).max() <= 1e-13 * np.abs(orr).max()
    # thread count does not change the result beyond the summation order inside a row (colours make rows exclusive)
    port.set_threads(1)
    v1, r1, _ = P.assemble(sol)
    assert np.array_equal(v1, v) and np.array_equal(r1, r)


def test_port_matches_compiled_reference_2d():
    port = _port()
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/libfeng_ref.so not built")
    from feng_b200 import coloring, mesh as M, problems as PB
    path = os.path.join(ref.DATA_DIR, "square2.msh")
    R = ref.RefProblem(path, "ns_div", 2, 8, field=1, mu=1. / 40., rho=1.0, p_essential=False)
    xyz, conn = R.mesh()
    adrU, adrP = R.adr(0), R.adr(1)
    sol, _ = R.solution()
    rng = np.random.default_rng(7)
    sol[:R.n_inc] += rng.uniform(-1e-2, 1e-2, R.n_inc)
    R.set_solution(sol)
    vr, rr, _ = R.assemble()
    ia, ja = R.pattern()

    class PB_:                                      # the reference's own tables in the HostProblem layout
        pass
    from feng_b200 import tables as T
    pb = PB_()
    pb.dim, pb.n_inc = 2, R.n_inc
    pb.mesh = PB_()
    pb.mesh.xyz, pb.mesh.cells = xyz, conn
    pb.adrU, pb.adrP = adrU, adrP
    pb.w, q = T.quadrature(2, 8)
    pb.LU, pb.dLU = T.basis(2, 2, q)
    pb.LP, _ = T.basis(2, 1, q)
    # the reference's source form is a zero field here (field 1): skip it, as the port does
    pb.forms = [PB.FormSpec(PB.VECTOR_CONVECTIVE_ACCELERATION, -1.0), PB.FormSpec(PB.MIXED_DIVERGENCE, 1.0),
                PB.FormSpec(PB.DIV_NEWTONIAN_STRESS, 1.0, 1. / 40.)]
    P = port.PortProblem(pb, R.colors())
    assert np.array_equal(P.ia, ia) and np.array_equal(P.ja, ja)
    v, r, _ = P.assemble(sol)
    assert np.abs(v - vr).max() <= 1e-13 * np.abs(vr).max()
    assert np.abs(r - rr).max() <= 1e-13 * np.abs(rr).max()
