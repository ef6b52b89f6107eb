"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/_ref/libfeng_ref.so.

libfeng_ref.so is the UNMODIFIED reference (arthurbawin/feNG) compiled by oracle/Makefile plus
oracle/ref_harness.cpp.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product (feng_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libfeng_ref.so")
# the same harness linked with the product's C++ adapter (adapter/feLinearSystemB200.h) + libfeng_b200.so
LIB_B200_PATH = os.path.join(_HERE, "_ref", "libfeng_ref_b200.so")
DATA_DIR = os.path.join(_HERE, "_ref", "data")     # copies of the reference's regression meshes (oracle/Makefile: data)

KIND = {"diffusion": 0, "stokes_div": 1, "ns_div": 2, "ns_lap": 3, "stokes_lap": 4, "chns": 5, "poiseuille_div": 6,
        "poiseuille_lap": 7, "periodic_diffusion": 8, "var_diffusion": 9}


class Recipe(C.Structure):
    _fields_ = [("kind", C.c_int), ("order", C.c_int), ("quad_degree", C.c_int), ("field", C.c_int),
                ("mu", C.c_double), ("rho", C.c_double), ("transient", C.c_int), ("p_essential", C.c_int)]


def available() -> bool:
    return os.path.exists(LIB_PATH)


def available_b200() -> bool:
    return os.path.exists(LIB_B200_PATH)


_libs = {}


def lib(b200: bool = False):
    if b200 not in _libs:
        L = C.CDLL(LIB_B200_PATH if b200 else LIB_PATH)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_char_p, C.POINTER(Recipe)]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_max_threads.restype = C.c_int
        _libs[b200] = L
    return _libs[b200]


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


@dataclass
class FormInfo:
    M: int
    N: int
    has_matrix: bool
    sys_id: int
    transient: bool


class RefProblem:
    """One (mesh, recipe) instance of the reference CPU path."""

    def __init__(self, mesh_file: str, kind: str, order: int = 2, quad_degree: int = 8, field: int = 0,
                 mu: float = 1.0, rho: float = 1.0, transient: bool = False, p_essential: bool = True,
                 b200: bool = False, chns=None):
        self.L = lib(b200)
        if chns is not None:
            # parameter block of the CHNS recipe (oracle/ref_harness.cpp: g_chns), see oracle/chns_oracle.ChnsParams
            a = np.ascontiguousarray(chns, np.float64)
            self.L.ref_set_chns_params(_p(a), int(a.size))
        rc = Recipe(KIND[kind], order, quad_degree, field, mu, rho, int(transient), int(p_essential))
        self.h = self.L.ref_create(mesh_file.encode(), C.byref(rc))
        if not self.h:
            raise RuntimeError(f"reference could not build problem on {mesh_file}")
        self.h = C.c_void_p(self.h)
        info = np.zeros(16, np.int64)
        self.L.ref_info(self.h, _p(info, C.c_int64))
        (self.dim, self.n_vertices, self.n_elm, self.nv, self.n_dof, self.n_inc, self.nnz, self.n_quad,
         self.n_colors, self.n_spaces, self.n_forms, self.n_matrix_forms) = (int(x) for x in info[:12])

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- tables -------------------------------------------------------------------------
    def mesh(self):
        xyz = np.zeros((self.n_vertices, 3))
        conn = np.zeros((self.n_elm, self.nv), np.int32)
        self.L.ref_get_mesh(self.h, _p(xyz), _p(conn, C.c_int32))
        return xyz, conn

    def space_info(self, s):
        out = np.zeros(4, np.int64)
        self.L.ref_space_info(self.h, s, _p(out, C.c_int64))
        return int(out[0]), int(out[1])

    def adr(self, s):
        nF, _ = self.space_info(s)
        a = np.zeros((self.n_elm, nF), np.int64)
        self.L.ref_get_adr(self.h, s, _p(a, C.c_int64))
        return a

    def tables(self, s):
        nF, nC = self.space_info(s)
        shp = (self.n_quad, nF) if nC == 1 else (self.n_quad, nF, nC)
        arrs = [np.zeros(shp) for _ in range(4)]
        self.L.ref_get_tables(self.h, s, *[_p(a) for a in arrs])
        return arrs

    def quadrature(self):
        arrs = [np.zeros(self.n_quad) for _ in range(4)]
        self.L.ref_get_quadrature(self.h, *[_p(a) for a in arrs])
        return arrs

    def jacobians(self):
        J = np.zeros((self.n_elm, self.n_quad))
        self.L.ref_get_jacobians(self.h, _p(J))
        return J

    def colors(self):
        c = np.zeros(self.n_elm, np.int32)
        self.L.ref_get_colors(self.h, _p(c, C.c_int32))
        return c

    def pattern(self):
        ia = np.zeros(self.n_inc + 1, np.int64)
        ja = np.zeros(self.nnz, np.int64)
        self.L.ref_get_pattern(self.h, _p(ia, C.c_int64), _p(ja, C.c_int64))
        return ia, ja

    # ---- state ----------------------------------------------------------------------------
    def solution(self):
        s = np.zeros(self.n_dof)
        d = np.zeros(self.n_dof)
        self.L.ref_get_solution(self.h, _p(s), _p(d))
        return s, d

    def set_solution(self, sol, sol_dot=None, c0=0.0, t=0.0):
        sol = np.ascontiguousarray(sol, np.float64)
        dp = None
        if sol_dot is not None:
            sol_dot = np.ascontiguousarray(sol_dot, np.float64)
            dp = _p(sol_dot)
        self.L.ref_set_solution(self.h, _p(sol), dp, C.c_double(c0), C.c_double(t))

    def set_solution_n(self, sol_n, dt=0.0):
        """state at the previous time step (the reference's global solAtTimeN; None = the current solution) and the time
        step (feSolution::setTimeStep)"""
        if sol_n is None:
            self.L.ref_set_solution_n(self.h, None, C.c_double(dt))
        else:
            a = np.ascontiguousarray(sol_n, np.float64)
            self.L.ref_set_solution_n(self.h, _p(a), C.c_double(dt))

    # ---- hot path -------------------------------------------------------------------------
    def form_info(self, f) -> FormInfo:
        out = np.zeros(8, np.int64)
        self.L.ref_form_info(self.h, f, _p(out, C.c_int64))
        return FormInfo(int(out[0]), int(out[1]), bool(out[2]), int(out[3]), bool(out[4]))

    def element(self, f, e):
        fi = self.form_info(f)
        Ae = np.zeros((fi.M, fi.N))
        Be = np.zeros(fi.M)
        aI = np.zeros(fi.M, np.int64)
        aJ = np.zeros(fi.N, np.int64)
        self.L.ref_element(self.h, f, e, _p(Ae), _p(Be), _p(aI, C.c_int64), _p(aJ, C.c_int64))
        return Ae, Be, aI, aJ

    def assemble(self, matrix=True, residual=True):
        vals = np.zeros(self.nnz)
        rhs = np.zeros(self.n_inc)
        sec = np.zeros(2)
        what = (2 if matrix else 0) | (1 if residual else 0)
        self.L.ref_assemble(self.h, what, _p(vals), _p(rhs), _p(sec))
        return vals, rhs, sec

    def constrain(self):
        vals = np.zeros(self.nnz)
        rhs = np.zeros(self.n_inc)
        self.L.ref_constrain(self.h, _p(vals), _p(rhs))
        return vals, rhs

    def constraint_rows(self):
        n = C.c_int64(0)
        self.L.ref_constraint_rows(self.h, None, C.byref(n))
        rows = np.zeros(max(n.value, 1), np.int64)
        self.L.ref_constraint_rows(self.h, _p(rows, C.c_int64), C.byref(n))
        return rows[:n.value]

    def solve_current(self):
        du = np.zeros(self.n_inc)
        norms = np.zeros(3)
        rc = self.L.ref_solve_current(self.h, _p(du), _p(norms))
        if rc != 0:
            raise RuntimeError("reference SparseLU failed")
        return du, norms

    def newton(self, tol_res=1e-10, tol_cor=1e-10, max_iter=10):
        sol = np.zeros(self.n_dof)
        out = np.zeros(8)
        rc = self.L.ref_newton(self.h, C.c_double(tol_res), C.c_double(tol_cor), max_iter, _p(sol), _p(out))
        if rc != 0:
            raise RuntimeError(f"reference Newton failed rc={rc}")
        return sol, out

    def newton_b200(self, tol_res=1e-10, tol_cor=1e-10, max_iter=10, rel_tol=1e-8, pc=6, restart=30,
                    lin_max_iter=10000, scatter=0, device_pattern=False, abs_tol=None):
        """The unmodified reference Newton loop driving the CUDA backend through adapter/feLinearSystemB200.h.
        -> (solution, dict(errU, errP, n_solves, krylov_iterations, norm_axb, converged))"""
        sol = np.zeros(self.n_dof)
        out = np.zeros(8)
        opts = np.array([pc, restart, lin_max_iter, scatter, int(device_pattern)], np.int32)
        self.L.ref_set_b200_abs_tol(C.c_double(-1.0 if abs_tol is None else abs_tol))
        rc = self.L.ref_newton_b200(self.h, C.c_double(tol_res), C.c_double(tol_cor), max_iter, C.c_double(rel_tol),
                                    _p(opts, C.c_int32), _p(sol), _p(out))
        if rc != 0:
            raise RuntimeError(f"Newton with the B200 backend failed rc={rc}")
        return sol, dict(errU=out[0], errP=out[1], n_solves=int(out[2]), krylov_iterations=int(out[3]),
                         norm_axb=out[4], converged=bool(out[5]))

    def transient(self, b200=False, scheme=2, t0=0.0, t1=0.1, n_steps=3, tol_res=1e-10, tol_cor=1e-10, max_iter=20, rel_tol=1e-10,
                  pc=6, restart=30, lin_max_iter=10000):
        """n_steps of the reference's unmodified BDF1 / BDF2 integrator + Newton loop from the recipe's initial state, on the CPU
        stub backend or (b200=True) on the CUDA backend through the adapter -> (solution, dict)"""
        sol = np.zeros(self.n_dof)
        out = np.zeros(4)
        opts = np.array([pc, restart, lin_max_iter], np.int32)
        rc = self.L.ref_transient(self.h, int(b200), int(scheme), C.c_double(t0), C.c_double(t1), int(n_steps), C.c_double(tol_res),
                                  C.c_double(tol_cor), int(max_iter), C.c_double(rel_tol), _p(opts, C.c_int32), _p(sol), _p(out))
        if rc != 0:
            raise RuntimeError(f"transient run failed rc={rc}")
        return sol, dict(n_solves=int(out[0]), krylov_iterations=int(out[1]), converged=bool(out[2]))

    def assemble_b200(self, matrix=True, residual=True, device_pattern=False):
        """One assembly of the current state by the CUDA backend, driven through adapter/feLinearSystemB200.h from the
        reference's own host objects (b200=True instances only)."""
        vals = np.zeros(self.nnz)
        rhs = np.zeros(self.n_inc)
        what = (2 if matrix else 0) | (1 if residual else 0)
        rc = self.L.ref_assemble_b200(self.h, what, int(device_pattern), _p(vals), _p(rhs))
        if rc != 0:
            raise RuntimeError(f"assembly through the B200 adapter failed rc={rc}")
        return vals, rhs

    def constrain_b200(self, device_pattern=False):
        """assemble + constrainEssentialComponents + applyPeriodicity by the CUDA backend through the adapter"""
        vals = np.zeros(self.nnz)
        rhs = np.zeros(self.n_inc)
        rc = self.L.ref_constrain_b200(self.h, int(device_pattern), _p(vals), _p(rhs))
        if rc != 0:
            raise RuntimeError(f"constrained assembly through the B200 adapter failed rc={rc}")
        return vals, rhs

    def periodic_pairs(self):
        n = C.c_int64(0)
        self.L.ref_periodic_pairs(self.h, None, None, C.byref(n))
        m = np.zeros(max(n.value, 1), np.int64)
        s = np.zeros(max(n.value, 1), np.int64)
        self.L.ref_periodic_pairs(self.h, _p(m, C.c_int64), _p(s, C.c_int64), C.byref(n))
        return m[:n.value], s[:n.value]

    def error_norms(self, sol):
        sol = np.ascontiguousarray(sol, np.float64)
        out = np.zeros(2)
        self.L.ref_error_norms(self.h, _p(sol), _p(out))
        return out


def set_threads(n: int):
    lib().ref_set_threads(n)


def max_threads() -> int:
    return lib().ref_max_threads()
