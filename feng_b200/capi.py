"""ctypes binding of the C ABI (include/feng_b200.h) -- the same symbols the C++ adapter links against.

There is no fallback: if the CUDA library is missing this module raises at import of `lib()`, and every compute
entry point fails with B200_ERR_CUDA when no device is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libfeng_b200.so")

# every symbol include/feng_b200.h declares (tests/test_capi_symbols.py checks the list against the header)
SYMBOLS = [
    "b200_last_error", "b200_kernel_launches", "b200_reset_kernel_launches", "b200_create", "b200_destroy",
    "b200_set_mesh", "b200_set_quadrature", "b200_add_space", "b200_add_form", "b200_set_source", "b200_set_form_coefficient", "b200_set_pattern",
    "b200_build_pattern", "b200_get_pattern_size", "b200_get_pattern", "b200_set_colors", "b200_set_scatter_mode", "b200_set_assembly_mode", "b200_has_gather_plan", "b200_gather_kernel", "b200_error_norm", "b200_unique_edges", "b200_add_form_chns",
    "b200_set_constraints", "b200_set_periodic", "b200_set_blocks", "b200_finalize", "b200_system_size", "b200_set_solution",
    "b200_set_solution_n", "b200_set_essential", "b200_state_push", "b200_state_bdf", "b200_set_to_zero", "b200_assemble", "b200_rhs_max_norm", "b200_du_max_norm", "b200_constrain",
    "b200_apply_periodicity", "b200_solve", "b200_correct_solution", "b200_get_rhs", "b200_axpy_rhs",
    "b200_get_matrix_values", "b200_get_du", "b200_get_solution", "b200_spmv", "b200_last_assemble_ms",
    "b200_last_solve_ms", "b200_time_spmv", "b200_sync", "b200_time_begin", "b200_time_end",
    "b200_measure_fp64_peak", "b200_measure_dmma_peak", "b200_comm_unique_id", "b200_comm_init", "b200_set_halo", "b200_halo_exchange_host",
]

SCATTER_ATOMIC, SCATTER_COLORED = 0, 1
ASSEMBLY_AUTO, ASSEMBLY_SCATTER, ASSEMBLY_GATHER = 0, 1, 2
PC_NONE, PC_JACOBI, PC_BLOCK_JACOBI, PC_AMG, PC_SCHUR_AMG, PC_AUTO = 0, 1, 2, 4, 5, 6


class SolverOptions(C.Structure):
    _fields_ = [("rel_tol", C.c_double), ("abs_tol", C.c_double), ("div_tol", C.c_double), ("max_iter", C.c_int),
                ("restart", C.c_int), ("pc", C.c_int)]


class ChnsParams(C.Structure):
    """b200_chns_params (include/feng_b200.h)"""
    _fields_ = [("rho_a", C.c_double), ("rho_b", C.c_double), ("visc_a", C.c_double), ("visc_b", C.c_double),
                ("mobility", C.c_double), ("surface_tension", C.c_double), ("epsilon", C.c_double),
                ("force", C.c_double * 3), ("source_u", C.c_double * 3), ("source_p", C.c_double),
                ("source_phi", C.c_double), ("source_mu", C.c_double), ("limiter", C.c_int),
                ("degenerate_mobility", C.c_int), ("mass_alpha", C.c_double), ("khanwale", C.c_double * 7)]


FORM_CHNS_ABELS, FORM_CHNS_MASS_AVERAGED, FORM_CHNS_KHANWALE = 35, 36, 38


class SolveInfo(C.Structure):
    _fields_ = [("norm_dx", C.c_double), ("norm_rhs", C.c_double), ("norm_axb", C.c_double),
                ("iterations", C.c_int), ("converged", C.c_int), ("rel_residual", C.c_double)]


class B200Error(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(the engine is CUDA-only; there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.b200_last_error.restype = C.c_char_p
        L.b200_kernel_launches.restype = C.c_int64
        L.b200_system_size.restype = C.c_int64
        L.b200_system_size.argtypes = [C.c_void_p]
        L.b200_gather_kernel.argtypes = [C.c_void_p]
        L.b200_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.b200_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _d(a):
    return _ptr(a, C.c_double)


def _i32(a):
    return _ptr(a, C.c_int32)


def _i64(a):
    return _ptr(a, C.c_int64)


def unique_edges(n_vertices: int, pairs, device: int = 0):
    """Device version of numbering.build_edges' numpy.unique: (edge_of_pair int32[n], edges int32[nEdges, 2])."""
    pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
    n = pairs.shape[0]
    eop = np.empty(n, np.int32)
    edges = np.empty((n, 2), np.int32)
    ne = C.c_int64(0)
    L = lib()
    L.b200_unique_edges.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
    check(L.b200_unique_edges(int(device), int(n_vertices), int(n), pairs.ctypes.data_as(C.c_void_p), eop.ctypes.data_as(C.c_void_p),
                              edges.ctypes.data_as(C.c_void_p), C.byref(ne)), "b200_unique_edges")
    return eop, np.ascontiguousarray(edges[:ne.value])


def check(rc, what=""):
    if rc < 0:
        raise B200Error(f"{what}: rc={rc}: {lib().b200_last_error().decode()}")
    return rc


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(lib().b200_comm_unique_id(buf), "b200_comm_unique_id")
    return buf.raw


class System:
    """Thin object wrapper over b200_system*; one method per C entry point."""

    def __init__(self, device: int = 0):
        self.L = lib()
        h = C.c_void_p()
        check(self.L.b200_create(C.byref(h), device), "b200_create")
        self.h = h
        self.n_inc = self.n_dof = self.nnz = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- set-up -------------------------------------------------------------------------------------
    def set_mesh(self, dim, xyz, cells):
        xyz = np.ascontiguousarray(xyz, np.float64)
        cells = np.ascontiguousarray(cells, np.int32)
        check(self.L.b200_set_mesh(self.h, dim, C.c_int64(xyz.shape[0]), _d(xyz), C.c_int64(cells.shape[0]),
                                   cells.shape[1], _i32(cells)), "b200_set_mesh")

    def set_quadrature(self, w):
        w = np.ascontiguousarray(w, np.float64)
        check(self.L.b200_set_quadrature(self.h, w.shape[0], _d(w)), "b200_set_quadrature")

    def add_space(self, n_scalar, ncomp, adr, L, dL):
        adr = np.ascontiguousarray(adr, np.int32)
        L = np.ascontiguousarray(L, np.float64)
        dL = np.ascontiguousarray(dL, np.float64)
        return check(self.L.b200_add_space(self.h, n_scalar, ncomp, _i32(adr), _d(L), _d(dL)), "b200_add_space")

    def add_form(self, kind, su, sp=-1, coeff=1.0, param=1.0, source=None):
        src = None if source is None else np.ascontiguousarray(source, np.float64)
        return check(self.L.b200_add_form(self.h, kind, su, sp, C.c_double(coeff), C.c_double(param), _d(src)),
                     "b200_add_form")

    def add_form_chns(self, su, sp, sf, sm, model, kind=None):
        """model: any object with the attributes of feng_b200.problems.ChnsModel"""
        if kind is None:
            kind = {"abels": FORM_CHNS_ABELS, "mass_averaged": FORM_CHNS_MASS_AVERAGED,
                    "khanwale": FORM_CHNS_KHANWALE}[getattr(model, "formulation", "abels")]
        prm = ChnsParams(model.rhoA, model.rhoB, model.viscA, model.viscB, model.mobility, model.sigma, model.epsilon,
                         (C.c_double * 3)(model.force[0], model.force[1], 0.0),
                         (C.c_double * 3)(model.src_u[0], model.src_u[1], 0.0), model.src_p, model.src_phi,
                         model.src_mu, int(model.limiter), int(model.degenerate_mobility), float(getattr(model, "alpha", 0.0)),
                         (C.c_double * 7)(*[float(x) for x in getattr(model, "khanwale", (1.,) * 7)]))
        return check(self.L.b200_add_form_chns(self.h, kind, su, sp, sf, sm, C.byref(prm)), "b200_add_form_chns")

    def set_source(self, form_id, source):
        src = np.ascontiguousarray(source, np.float64)
        check(self.L.b200_set_source(self.h, form_id, _d(src)), "b200_set_source")

    def set_pattern(self, n_inc, n_dof, ia, ja):
        ia = np.ascontiguousarray(ia, np.int64)
        ja = np.ascontiguousarray(ja, np.int32)
        check(self.L.b200_set_pattern(self.h, C.c_int64(n_inc), C.c_int64(n_dof), _i64(ia), _i32(ja)),
              "b200_set_pattern")
        self.n_inc, self.n_dof, self.nnz = int(n_inc), int(n_dof), int(ia[-1])

    def build_pattern(self, n_inc, n_dof):
        """device-side EZCRS pattern from the spaces and forms registered so far"""
        check(self.L.b200_build_pattern(self.h, C.c_int64(n_inc), C.c_int64(n_dof)), "b200_build_pattern")
        ni, nz = C.c_int64(), C.c_int64()
        check(self.L.b200_get_pattern_size(self.h, C.byref(ni), C.byref(nz)), "b200_get_pattern_size")
        self.n_inc, self.n_dof, self.nnz = int(ni.value), int(n_dof), int(nz.value)

    def get_pattern(self):
        ia = np.zeros(self.n_inc + 1, np.int64)
        ja = np.zeros(self.nnz, np.int32)
        check(self.L.b200_get_pattern(self.h, _i64(ia), _i32(ja)), "b200_get_pattern")
        return ia, ja

    def set_colors(self, n_colors, colors):
        colors = np.ascontiguousarray(colors, np.int32)
        check(self.L.b200_set_colors(self.h, n_colors, _i32(colors)), "b200_set_colors")

    def set_scatter_mode(self, mode):
        check(self.L.b200_set_scatter_mode(self.h, mode), "b200_set_scatter_mode")

    def comm_init(self, ident: bytes, rank: int, world: int):
        check(self.L.b200_comm_init(self.h, C.c_char_p(ident), rank, world), "b200_comm_init")

    def set_halo(self, owned, neighbors, send_ptr, send_idx, recv_ptr, recv_idx):
        owned = np.ascontiguousarray(owned, np.uint8)
        nb = np.ascontiguousarray(neighbors, np.int32)
        sp, si = np.ascontiguousarray(send_ptr, np.int64), np.ascontiguousarray(send_idx, np.int32)
        rp, ri = np.ascontiguousarray(recv_ptr, np.int64), np.ascontiguousarray(recv_idx, np.int32)
        check(self.L.b200_set_halo(self.h, owned.ctypes.data_as(C.POINTER(C.c_uint8)), nb.shape[0], _i32(nb), _i64(sp),
                                   _i32(si), _i64(rp), _i32(ri)), "b200_set_halo")

    def halo_exchange_host(self, x):
        x = np.ascontiguousarray(x, np.float64).copy()
        check(self.L.b200_halo_exchange_host(self.h, _d(x)), "b200_halo_exchange_host")
        return x

    def set_assembly_mode(self, mode):
        check(self.L.b200_set_assembly_mode(self.h, mode), "b200_set_assembly_mode")

    def has_gather_plan(self) -> bool:
        return bool(self.L.b200_has_gather_plan(self.h))

    def gather_kernel(self) -> int:
        """velocity-row kernels of the gather plan: 0 none, 1 thread per node, 2 lane groups, 3 row lanes (gather_urow.cuh)"""
        return int(self.L.b200_gather_kernel(self.h))

    def error_norm(self, space: int, kind: int = 0, p: int = 2, exact=None) -> float:
        """feNorm on the device: kind 0 = Lp norm of (exact - uh), kind 1 = H1 seminorm; exact[nElm, nq, ncomp(, dim)] or None"""
        ex = None if exact is None else np.ascontiguousarray(exact, np.float64)
        out = C.c_double(0.)
        check(self.L.b200_error_norm(self.h, int(space), int(kind), int(p), None if ex is None else _d(ex), C.byref(out)),
              "b200_error_norm")
        return float(out.value)

    def gather_plan_kind(self) -> int:
        """0 scatter kernels, 1 row-owner gather plan, 2 patch plan"""
        return int(self.L.b200_has_gather_plan(self.h))

    def set_constraints(self, rows, master=None, slave=None):
        rows = np.ascontiguousarray(rows, np.int64)
        m = None if master is None else np.ascontiguousarray(master, np.int64)
        s = None if slave is None else np.ascontiguousarray(slave, np.int64)
        check(self.L.b200_set_constraints(self.h, C.c_int64(rows.shape[0]), _i64(rows),
                                          C.c_int64(0 if m is None else m.shape[0]), _i64(m), _i64(s)),
              "b200_set_constraints")

    def set_form_coefficient(self, form_id, table):
        t = np.ascontiguousarray(table, np.float64)
        check(self.L.b200_set_form_coefficient(self.h, int(form_id), _d(t)), "b200_set_form_coefficient")

    def set_periodic(self, master, slave):
        m = np.ascontiguousarray(master, np.int64)
        s = np.ascontiguousarray(slave, np.int64)
        check(self.L.b200_set_periodic(self.h, C.c_int64(m.shape[0]), _i64(m), _i64(s)), "b200_set_periodic")

    def set_essential(self, dofs, values):
        """essential-BC refresh on the device copy; pass the SAME dofs array every step (it is uploaded once)"""
        assert dofs.dtype == np.int64 and dofs.flags.c_contiguous and values.dtype == np.float64
        check(self.L.b200_set_essential(self.h, C.c_int64(dofs.shape[0]), _i64(dofs), _d(values)), "b200_set_essential")

    def state_push(self):
        check(self.L.b200_state_push(self.h), "b200_state_push")

    def state_bdf(self, coef, t=0.0, dt=0.0):
        c = np.ascontiguousarray(coef, np.float64)
        check(self.L.b200_state_bdf(self.h, int(c.shape[0]), _d(c), C.c_double(t), C.c_double(dt)), "b200_state_bdf")

    def set_blocks(self, block_ptr, block_rows):
        bp = np.ascontiguousarray(block_ptr, np.int64)
        br = np.ascontiguousarray(block_rows, np.int64)
        check(self.L.b200_set_blocks(self.h, C.c_int64(bp.shape[0] - 1), _i64(bp), _i64(br)), "b200_set_blocks")

    def finalize(self):
        check(self.L.b200_finalize(self.h), "b200_finalize")

    # ---- feLinearSystem virtuals ----------------------------------------------------------------------
    def set_solution(self, sol, sol_dot=None, c0=0.0, t=0.0):
        sol = np.ascontiguousarray(sol, np.float64)
        sd = None if sol_dot is None else np.ascontiguousarray(sol_dot, np.float64)
        check(self.L.b200_set_solution(self.h, _d(sol), _d(sd), C.c_double(c0), C.c_double(t)), "b200_set_solution")

    def set_solution_n(self, sol_n, dt=0.0):
        """state at the previous time step (the reference's global solAtTimeN; None = the current solution) and the time
        step (feSolution::getTimeStep)"""
        if sol_n is None:
            check(self.L.b200_set_solution_n(self.h, None, C.c_double(dt)), "b200_set_solution_n")
        else:
            a = np.ascontiguousarray(sol_n, np.float64)
            check(self.L.b200_set_solution_n(self.h, _d(a), C.c_double(dt)), "b200_set_solution_n")

    def set_to_zero(self, what=3):
        check(self.L.b200_set_to_zero(self.h, what), "b200_set_to_zero")

    def assemble(self, what=3, only_transient=False):
        check(self.L.b200_assemble(self.h, what, int(only_transient)), "b200_assemble")

    def rhs_max_norm(self):
        v = C.c_double()
        check(self.L.b200_rhs_max_norm(self.h, C.byref(v)), "b200_rhs_max_norm")
        return v.value

    def du_max_norm(self):
        v = C.c_double()
        check(self.L.b200_du_max_norm(self.h, C.byref(v)), "b200_du_max_norm")
        return v.value

    def constrain(self):
        check(self.L.b200_constrain(self.h), "b200_constrain")

    def apply_periodicity(self):
        check(self.L.b200_apply_periodicity(self.h), "b200_apply_periodicity")

    def solve(self, rel_tol=1e-8, abs_tol=1e-14, div_tol=1e6, max_iter=10000, restart=30, pc=PC_JACOBI,
              raise_on_fail=True):
        opt = SolverOptions(rel_tol, abs_tol, div_tol, max_iter, restart, pc)
        info = SolveInfo()
        rc = self.L.b200_solve(self.h, C.byref(opt), C.byref(info))
        if raise_on_fail:
            check(rc, "b200_solve")
        return info

    def correct_solution(self, sol_host=None, correct_dot=False):
        check(self.L.b200_correct_solution(self.h, _d(sol_host), int(correct_dot)), "b200_correct_solution")

    def get_rhs(self):
        out = np.zeros(self.n_inc)
        check(self.L.b200_get_rhs(self.h, _d(out)), "b200_get_rhs")
        return out

    def axpy_rhs(self, coeff, d):
        d = np.ascontiguousarray(d, np.float64)
        check(self.L.b200_axpy_rhs(self.h, C.c_double(coeff), _d(d)), "b200_axpy_rhs")

    def get_matrix_values(self):
        out = np.zeros(self.nnz)
        check(self.L.b200_get_matrix_values(self.h, _d(out)), "b200_get_matrix_values")
        return out

    def get_du(self):
        out = np.zeros(self.n_inc)
        check(self.L.b200_get_du(self.h, _d(out)), "b200_get_du")
        return out

    def get_solution(self):
        out = np.zeros(self.n_dof)
        check(self.L.b200_get_solution(self.h, _d(out)), "b200_get_solution")
        return out

    def spmv(self, x):
        x = np.ascontiguousarray(x, np.float64)
        y = np.zeros(self.n_inc)
        check(self.L.b200_spmv(self.h, _d(x), _d(y)), "b200_spmv")
        return y

    # ---- measurement ----------------------------------------------------------------------------------
    def last_assemble_ms(self):
        v = C.c_float()
        check(self.L.b200_last_assemble_ms(self.h, C.byref(v)), "b200_last_assemble_ms")
        return v.value

    def last_solve_ms(self):
        v = C.c_float()
        check(self.L.b200_last_solve_ms(self.h, C.byref(v)), "b200_last_solve_ms")
        return v.value

    def time_spmv(self, reps=20):
        v = C.c_float()
        check(self.L.b200_time_spmv(self.h, reps, C.byref(v)), "b200_time_spmv")
        return v.value

    def sync(self):
        check(self.L.b200_sync(self.h), "b200_sync")

    def time_begin(self):
        check(self.L.b200_time_begin(self.h), "b200_time_begin")

    def time_end(self):
        v = C.c_float()
        check(self.L.b200_time_end(self.h, C.byref(v)), "b200_time_end")
        return v.value


def measure_fp64_peak(device: int = 0) -> float:
    v = C.c_double()
    check(lib().b200_measure_fp64_peak(device, C.byref(v)), "b200_measure_fp64_peak")
    return v.value


def measure_dmma_peak(device: int = 0) -> float:
    v = C.c_double()
    check(lib().b200_measure_dmma_peak(device, C.byref(v)), "b200_measure_dmma_peak")
    return v.value


def kernel_launches() -> int:
    return int(lib().b200_kernel_launches())


def reset_kernel_launches():
    lib().b200_reset_kernel_launches()
