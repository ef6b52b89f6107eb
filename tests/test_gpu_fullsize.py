"""Parity at the sizes bench.py runs: size-independent cross-checks against the CPU.

The assembled operator of the CUDA row-owner kernels (thread per node in 2-D, lane groups in 3-D) is compared, on EVERY row,
with the CPU restatement of the reference's element loops applied matrix-free (oracle/port_cpp.cpp: port_apply; pinned on the
numpy oracle and on the compiled reference by tests/test_port_cpp.py): A x for seeded random x -- a checksum of every row of
the 1.86 G-entry matrix -- and the rhs.  The independent CUDA quadrature-loop kernel with atomic scatter must agree as well,
and the write-once kernels must repeat bitwise.  Tolerance 1e-12 relative to the largest entry of the compared vector."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _probe(pb, sol, xs, kernel, monkeypatch, scatter=False):
    from feng_b200 import capi
    from feng_b200.linear_system import LinearSystemB200
    if kernel:
        monkeypatch.setenv("B200_GATHER_KERNEL", kernel)
    else:
        monkeypatch.delenv("B200_GATHER_KERNEL", raising=False)
    ls = LinearSystemB200(pb, device_pattern=True)
    S = ls.sys
    if scatter:
        S.set_assembly_mode(capi.ASSEMBLY_SCATTER)
    else:
        assert S.gather_plan_kind() == 1
    S.set_solution(sol)
    S.set_to_zero(3)
    S.assemble(3)
    ys = [S.spmv(x) for x in xs]
    rhs = S.get_rhs()
    again = None
    if not scatter:
        S.set_to_zero(3)
        S.assemble(3)
        again = ([S.spmv(x) for x in xs], S.get_rhs())
    S.close()
    return ys, rhs, again


@pytest.mark.parametrize("workload,n", [("t2d", 1024), ("t3d", 92)])
def test_three_assembly_paths_agree_at_bench_size(workload, n, monkeypatch):
    from feng_b200 import mesh as M, problems as PB
    if workload == "t2d":
        pb = PB.taylor_hood(M.square_mesh(n), "ns_div", 8, 1, 1 / 40., 1.0, build_pattern=False, with_source=False)
    else:
        pb = PB.taylor_hood(M.cube_mesh(n), "ns_div", 6, 3, 1 / 40., 1.0, build_pattern=False, with_source=False)
    sol = PB.perturb_unknowns(pb)
    rng = np.random.default_rng(20261017)
    xs = [rng.standard_normal(pb.n_inc) for _ in range(2)]
    ref_y, ref_r, ref_again = _probe(pb, sol, xs, None, monkeypatch)                   # default row-owner kernels
    for y, y2 in zip(ref_y, ref_again[0]):
        assert np.array_equal(y, y2)                                                  # write-once: bitwise repeatable
    assert np.array_equal(ref_r, ref_again[1])
    # CPU: the reference's element loops applied matrix-free (all host threads)
    from oracle import port
    if port.available():
        cy, cr, sec = port.PortProblem(pb, pattern_free=True).apply(sol, xs[:1])
        assert np.abs(cy[0] - ref_y[0]).max() <= 1e-12 * np.abs(cy[0]).max(), "GPU operator differs from the CPU element loops"
        assert np.abs(cr - ref_r).max() <= 1e-12 * np.abs(cr).max(), "GPU rhs differs from the CPU element loops"
    else:
        pytest.skip("oracle/_ref/libfeng_port.so not built")
    for name, kw in (("scatter", dict(kernel=None, scatter=True)),):
        ys, r, again = _probe(pb, sol, xs, kw.get("kernel"), monkeypatch, kw.get("scatter", False))
        for y, yr in zip(ys, ref_y):
            assert np.abs(y - yr).max() <= 1e-12 * np.abs(yr).max(), f"{name}: A x differs"
        assert np.abs(r - ref_r).max() <= 1e-12 * np.abs(ref_r).max(), f"{name}: rhs differs"
        if again is not None:
            assert all(np.array_equal(a, b) for a, b in zip(ys, again[0])) and np.array_equal(r, again[1])
