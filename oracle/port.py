"""TEST / BASELINE INFRASTRUCTURE ONLY -- ctypes front-end of oracle/_ref/libfeng_port.so (oracle/port_cpp.cpp): the C++/OpenMP
restatement of the reference's CPU assembly path, dim-generic, used as the timed CPU arm of the 3-D workload (the reference
has no vector-valued space on tetrahedra).  Only tests/ and bench.py's CPU legs may import this module."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libfeng_port.so")
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.port_assemble.restype = C.c_double
        L.port_apply.restype = C.c_double
        L.port_pattern.restype = C.c_int64
        L.port_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def max_threads() -> int:
    return lib().port_max_threads()


def set_threads(n: int):
    lib().port_set_threads(int(n))


def pattern(adrU, adrP, n_inc):
    """feEZCompressedRowStorage pattern (U x U, U x P, P x U couplings + forced diagonal) -> ia (int64), ja (int32)"""
    aU = np.ascontiguousarray(adrU, np.int64)
    aP = np.ascontiguousarray(adrP, np.int64)
    ia = np.zeros(n_inc + 1, np.int64)
    args = [C.c_int64(aU.shape[0]), aU.shape[1], aP.shape[1], _p(aU, C.c_int64), _p(aP, C.c_int64), C.c_int64(n_inc), _p(ia, C.c_int64)]
    nnz = lib().port_pattern(*args, None)
    ja = np.zeros(nnz, np.int32)
    lib().port_pattern(*args, _p(ja, C.c_int32))
    return ia, ja


class PortProblem:
    """Flat tables of one Taylor-Hood problem (feng_b200.problems.HostProblem) in the layout port_assemble takes."""

    def __init__(self, pb, colors=None, ia=None, ja=None, pattern_free=False):
        self.dim = pb.dim
        self.xyz = np.ascontiguousarray(pb.mesh.xyz, np.float64)
        self.cells = np.ascontiguousarray(pb.mesh.cells, np.int32)
        self.adrU = np.ascontiguousarray(pb.adrU, np.int64)
        self.adrP = np.ascontiguousarray(pb.adrP, np.int64)
        self.w = np.ascontiguousarray(pb.w, np.float64)
        self.LU = np.ascontiguousarray(pb.LU, np.float64)
        self.dLU = np.ascontiguousarray(pb.dLU, np.float64)
        self.LP = np.ascontiguousarray(pb.LP, np.float64)
        self.n_inc = int(pb.n_inc)
        if pattern_free:                      # matrix-free use only (apply)
            ia, ja = np.zeros(1, np.int64), np.zeros(0, np.int32)
        elif ia is None:
            ia, ja = pattern(self.adrU, self.adrP, self.n_inc)
        self.ia = np.ascontiguousarray(ia, np.int64)
        self.ja = np.ascontiguousarray(ja, np.int32)
        forms = [f for f in pb.forms if f.kind in (18, 22, 25, 26, 31)]
        assert len(forms) == len([f for f in pb.forms if f.source is None]), "source / transient forms are not restated"
        self.kinds = np.array([f.kind for f in forms], np.int32)
        self.coeff = np.array([f.coeff for f in forms], np.float64)
        self.param = np.array([f.param for f in forms], np.float64)
        if colors is None:
            colors = np.zeros(self.cells.shape[0], np.int64)
            assert pattern_free, "assemble() needs the element colours"
        colors = np.asarray(colors, np.int64)
        nc = int(colors.max()) + 1
        order = np.argsort(colors, kind="stable")
        self.color_elems = np.ascontiguousarray(order, np.int32)
        self.color_ptr = np.concatenate([[0], np.cumsum(np.bincount(colors, minlength=nc))]).astype(np.int64)
        self.n_colors = nc
        self.vals = np.zeros(self.ja.shape[0])
        self.rhs = np.zeros(self.n_inc)
        self.pattern_free = pattern_free

    def assemble(self, sol, matrix=True, residual=True):
        """-> (vals, rhs, seconds)"""
        sol = np.ascontiguousarray(sol, np.float64)
        what = (2 if matrix else 0) | (1 if residual else 0)
        sec = lib().port_assemble(self.dim, C.c_int64(self.cells.shape[0]), _p(self.xyz, C.c_double), _p(self.cells, C.c_int32),
                                  _p(self.adrU, C.c_int64), _p(self.adrP, C.c_int64), self.LU.shape[1], self.LP.shape[1],
                                  self.w.shape[0], _p(self.w, C.c_double), _p(self.LU, C.c_double), _p(self.dLU, C.c_double),
                                  _p(self.LP, C.c_double), C.c_int64(self.n_inc), _p(self.ia, C.c_int64), _p(self.ja, C.c_int32),
                                  _p(sol, C.c_double), int(self.kinds.shape[0]), _p(self.kinds, C.c_int32), _p(self.coeff, C.c_double),
                                  _p(self.param, C.c_double), self.n_colors, _p(self.color_ptr, C.c_int64),
                                  _p(self.color_elems, C.c_int32), what, _p(self.vals, C.c_double), _p(self.rhs, C.c_double))
        return self.vals, self.rhs, float(sec)

    def apply(self, sol, xs):
        """matrix-free y_v = A(sol) x_v for every vector of xs, and the rhs -> (ys, rhs, seconds)"""
        sol = np.ascontiguousarray(sol, np.float64)
        x = np.ascontiguousarray(np.stack(xs), np.float64)
        y = np.zeros_like(x)
        rhs = np.zeros(self.n_inc)
        sec = lib().port_apply(self.dim, C.c_int64(self.cells.shape[0]), _p(self.xyz, C.c_double), _p(self.cells, C.c_int32),
                               _p(self.adrU, C.c_int64), _p(self.adrP, C.c_int64), self.LU.shape[1], self.LP.shape[1], self.w.shape[0],
                               _p(self.w, C.c_double), _p(self.LU, C.c_double), _p(self.dLU, C.c_double), _p(self.LP, C.c_double),
                               C.c_int64(self.n_inc), _p(sol, C.c_double), int(self.kinds.shape[0]), _p(self.kinds, C.c_int32),
                               _p(self.coeff, C.c_double), _p(self.param, C.c_double), int(x.shape[0]), _p(x, C.c_double),
                               _p(y, C.c_double), _p(rhs, C.c_double))
        return [y[i] for i in range(x.shape[0])], rhs, float(sec)
