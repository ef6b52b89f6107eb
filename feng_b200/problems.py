"""Problem set-up on the host: mesh -> numbering -> tables -> form list -> initial state.

Mirrors what a reference driver does before it creates its linear system
(tests/withLinearSolver/navier_stokes.cpp:63-99, convergenceLaplace.cpp:56-77): spaces, feMetaNumber, feSolution
initialised node-wise from the analytic fields (src/feSolution.cpp:112-244), list of weak forms with their constant
coefficients.  The result is the set of flat tables the C ABI consumes (include/feng_b200.h).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import numbering as NB
from . import pattern as PT
from . import tables as T
from .mesh import Mesh

# elementSystemType values of the reference (src/feSysElm.h:13-77); the C ABI uses the same numbers.
SOURCE = 0
VECTOR_SOURCE = 2
TRANSIENT_MASS = 13
TRANSIENT_VECTOR_MASS = 15
DIFFUSION = 16
VECTOR_DIFFUSION = 18
VECTOR_CONVECTIVE_ACCELERATION = 22
DIV_NEWTONIAN_STRESS = 25
MIXED_GRADIENT = 26
MIXED_DIVERGENCE = 31

_LAYOUT = {DIV_NEWTONIAN_STRESS: (("U",), ("U", "P")), MIXED_DIVERGENCE: (("P",), ("U",)),
           MIXED_GRADIENT: (("U",), ("P",))}


def form_layout(kind):
    return _LAYOUT.get(kind, (("U",), ("U",)))


def has_matrix(kind):
    return kind not in (SOURCE, VECTOR_SOURCE)


@dataclass
class FormSpec:
    kind: int
    coeff: float = 1.0
    param: float = 1.0
    source: np.ndarray | None = None    # (nE, nq) or (nE, nq, dim) host tabulation of the source callback


# ---- analytic fields (numpy twins of the callbacks in oracle/ref_harness.cpp) -------------------------------
def u_exact(field_id, x, mu, rho):
    X, Y = x[..., 0], x[..., 1]
    if field_id == 0:     # tests/withLinearSolver/navier_stokes.cpp:19-25
        return np.stack([X ** 4 * Y ** 4, -4. / 5. * X ** 3 * Y ** 5], -1)
    if field_id == 1:     # Kovasznay, Re = rho/mu
        Re = rho / mu
        lam = Re / 2. - np.sqrt(Re * Re / 4. + 4. * np.pi ** 2)
        return np.stack([1. - np.exp(lam * X) * np.cos(2. * np.pi * Y),
                         lam / (2. * np.pi) * np.exp(lam * X) * np.sin(2. * np.pi * Y)], -1)
    if field_id == 3:     # smooth 3-D field of SURVEY.md section 8(d)
        Z = x[..., 2]
        return np.stack([np.sin(np.pi * Y) * np.cos(np.pi * Z), np.sin(np.pi * Z) * np.cos(np.pi * X),
                         np.sin(np.pi * X) * np.cos(np.pi * Y)], -1)
    return np.zeros(x.shape[:-1] + (2,))


def p_exact(field_id, x, mu, rho):
    X, Y = x[..., 0], x[..., 1]
    if field_id == 0:
        return X * X * Y * Y
    if field_id == 1:
        Re = rho / mu
        lam = Re / 2. - np.sqrt(Re * Re / 4. + 4. * np.pi ** 2)
        return 0.5 * rho * (1. - np.exp(2. * lam * X))
    if field_id == 3:
        return np.sin(np.pi * X) * np.sin(np.pi * Y) * np.sin(np.pi * x[..., 2])
    return np.zeros(x.shape[:-1])


def u_source(field_id, x, mu, rho, with_conv):
    """Momentum source of the reference's MMS driver generalised to (mu, rho):
    -( -rho (u.grad)u - grad p + mu lap u )   (tests/withLinearSolver/navier_stokes.cpp:33-52)."""
    X, Y = x[..., 0], x[..., 1]
    if field_id != 0:
        return np.zeros(x.shape[:-1] + (x.shape[-1] if x.shape[-1] == 3 and field_id == 3 else 2,))
    mdp = [-2. * X * Y * Y, -2. * X * X * Y]
    lap = [12. * (X * X * Y ** 4 + X ** 4 * Y * Y), -4. / 5. * (6. * X * Y ** 5 + 20. * X ** 3 * Y ** 3)]
    u = [X ** 4 * Y ** 4, -4. / 5. * X ** 3 * Y ** 5]
    gu = [[4. * X ** 3 * Y ** 4, -12. * X * X * Y ** 5 / 5.], [4. * X ** 4 * Y ** 3, -4. * X ** 3 * Y ** 4]]
    ugu = [u[0] * gu[0][0] + u[1] * gu[1][0], u[0] * gu[0][1] + u[1] * gu[1][1]]
    c = 1.0 if with_conv else 0.0
    return np.stack([-(-c * rho * ugu[i] + mdp[i] + mu * lap[i]) for i in range(2)], -1)


def s_exact(field_id, x):
    if field_id == 0:     # tests/withLinearSolver/convergenceLaplace.cpp:18-23 (+ z^6 in 3-D)
        return (x ** 6).sum(-1)
    return np.zeros(x.shape[:-1])


def s_source(field_id, x, k):
    if field_id == 0:
        return k * 30. * (x ** 4).sum(-1)
    return -np.ones(x.shape[:-1])


@dataclass
class HostProblem:
    mesh: Mesh
    dim: int
    ncomp: int                      # components of the primary field (1 scalar, dim vector)
    order: int
    num: NB.Numbering
    adrU: np.ndarray                # (nE, nS*ncomp) int64
    adrP: np.ndarray | None
    n_inc: int
    n_dof: int
    w: np.ndarray
    qpts: np.ndarray
    LU: np.ndarray
    dLU: np.ndarray
    LP: np.ndarray | None
    dLP: np.ndarray | None
    forms: list = field(default_factory=list)
    sol: np.ndarray | None = None
    ia: np.ndarray | None = None
    ja: np.ndarray | None = None
    meta: dict = field(default_factory=dict)
    # Cahn-Hilliard Navier-Stokes (config 5): phase marker / chemical potential spaces and the model parameters
    adrF: np.ndarray | None = None
    adrM: np.ndarray | None = None
    LF: np.ndarray | None = None
    dLF: np.ndarray | None = None
    chns: object | None = None

    def couplings(self):
        if self.chns is not None:
            a = np.concatenate([self.adrU, self.adrP, self.adrF, self.adrM], 1)
            return [(a, a)]
        adr = {"U": self.adrU, "P": self.adrP}
        out = []
        for f in self.forms:
            if has_matrix(f.kind):
                r, c = form_layout(f.kind)
                out.append((np.concatenate([adr[x] for x in r], 1), np.concatenate([adr[x] for x in c], 1)))
        return out

    def build_pattern(self):
        self.ia, self.ja = PT.build_pattern(self.n_inc, self.couplings())
        return self.ia, self.ja


def quad_points_physical(mesh: Mesh, qpts: np.ndarray) -> np.ndarray:
    """x[e, k, :] = sum_v L1_v(xi_k) x_v  (feSpace::interpolateVectorFieldAtQuadNode on the geometric space)."""
    L1, _ = T.basis(mesh.dim, 1, qpts)
    X = mesh.xyz[mesh.cells]                       # (nE, nv, 3)
    out = np.zeros((mesh.n_cells, qpts.shape[0], 3))
    for v in range(X.shape[1]):
        out += L1[None, :, v, None] * X[:, None, v, :]
    return out


def dof_coordinates(mesh: Mesh, num: NB.Numbering, fld: str, n_dof: int) -> np.ndarray:
    """Physical location of every DOF of a field (vertices, then edge mid-points), NaN elsewhere."""
    fn = num.fields[fld]
    xyz = np.full((n_dof, 3), np.nan)
    ok = fn.vertex_dof[:, 0] >= 0
    for c in range(fn.ncomp):
        xyz[fn.vertex_dof[ok, c]] = mesh.xyz[ok]
    ok = fn.edge_dof[:, 0] >= 0
    if ok.any():
        mid = 0.5 * mesh.xyz[num.edges[:, 0]] + 0.5 * mesh.xyz[num.edges[:, 1]]
        for c in range(fn.ncomp):
            xyz[fn.edge_dof[ok, c]] = mid[ok]
    return xyz


def dof_components(num: NB.Numbering, fld: str, n_dof: int) -> np.ndarray:
    fn = num.fields[fld]
    comp = np.full(n_dof, -1, np.int64)
    for c in range(fn.ncomp):
        v = fn.vertex_dof[:, c]
        comp[v[v >= 0]] = c
        e = fn.edge_dof[:, c]
        comp[e[e >= 0]] = c
    return comp


def taylor_hood(mesh: Mesh, kind: str = "ns_div", quad_degree: int = 8, field_id: int = 0, mu: float = 1.0,
                rho: float = 1.0, transient: bool = False, p_essential: bool = False,
                build_pattern: bool = True, with_source: bool = True) -> HostProblem:
    """P2/P1 (Navier-)Stokes with the form list and sign conventions of the reference's drivers
    (tests/withLinearSolver/navier_stokes.cpp:82-99): convU(-rho), divU(+1), source, then either
    divSigma(+1, mu) or diffU(-1, mu) + gradP(-1); optional transient mass(-rho)."""
    dim = mesh.dim
    with_conv = kind in ("ns_div", "ns_lap")
    div_form = kind in ("ns_div", "stokes_div")
    num = NB.build_numbering(mesh, NB.taylor_hood_spaces(dim, p_essential, mesh.point_pressure is not None))
    w, q = T.quadrature(dim, quad_degree)
    LU, dLU = T.basis(dim, 2, q)
    LP, dLP = T.basis(dim, 1, q)
    pb = HostProblem(mesh, dim, dim, 2, num, num.adr(mesh, "U", 2), num.adr(mesh, "P", 1), num.n_inc, num.n_dof,
                     w, q, LU, dLU, LP, dLP)
    forms = []
    if with_conv:
        forms.append(FormSpec(VECTOR_CONVECTIVE_ACCELERATION, -rho))
    forms.append(FormSpec(MIXED_DIVERGENCE, 1.0))
    if with_source:
        xq = quad_points_physical(mesh, q)
        src = u_source(field_id, xq[..., :dim] if dim == 2 else xq, mu, rho, with_conv)
        if src.shape[-1] != dim:
            src = np.zeros(xq.shape[:2] + (dim,))
        forms.append(FormSpec(VECTOR_SOURCE, 1.0, 0.0, np.ascontiguousarray(src)))
    if div_form:
        forms.append(FormSpec(DIV_NEWTONIAN_STRESS, 1.0, mu))
    else:
        forms.append(FormSpec(VECTOR_DIFFUSION, -1.0, mu))
        forms.append(FormSpec(MIXED_GRADIENT, -1.0))
    if transient:
        forms.append(FormSpec(TRANSIENT_VECTOR_MASS, -rho))
    pb.forms = forms
    # node-wise initialisation of every DOF (unknown and essential) from the analytic fields
    sol = np.zeros(num.n_dof)
    xu = dof_coordinates(mesh, num, "U", num.n_dof)
    cu = dof_components(num, "U", num.n_dof)
    isU = cu >= 0
    uval = u_exact(field_id, xu[isU], mu, rho)
    sol[isU] = uval[np.arange(uval.shape[0]), cu[isU]] if uval.shape[-1] > 1 else uval[:, 0]
    xp = dof_coordinates(mesh, num, "P", num.n_dof)
    isP = dof_components(num, "P", num.n_dof) >= 0
    sol[isP] = p_exact(field_id, xp[isP], mu, rho)
    pb.sol = sol
    pb.meta = dict(kind=kind, quad_degree=quad_degree, field=field_id, mu=mu, rho=rho, transient=transient,
                   p_essential=p_essential)
    if build_pattern:
        pb.build_pattern()
    return pb


def scalar_diffusion(mesh: Mesh, order: int = 2, quad_degree: int = 12, field_id: int = 0, k: float = 1.0,
                     transient: bool = False, rho: float = 1.0, build_pattern: bool = True) -> HostProblem:
    """Scalar diffusion + source (tests/withLinearSolver/convergenceLaplace.cpp:56-77, exe/example1.cpp:133-170)."""
    dim = mesh.dim
    num = NB.build_numbering(mesh, NB.scalar_spaces(order))
    w, q = T.quadrature(dim, quad_degree)
    LU, dLU = T.basis(dim, order, q)
    pb = HostProblem(mesh, dim, 1, order, num, num.adr(mesh, "U", order), None, num.n_inc, num.n_dof, w, q, LU, dLU,
                     None, None)
    xq = quad_points_physical(mesh, q)
    src = s_source(field_id, xq[..., :dim], k)
    pb.forms = [FormSpec(DIFFUSION, 1.0, k), FormSpec(SOURCE, 1.0, 0.0, np.ascontiguousarray(src))]
    if transient:
        pb.forms.append(FormSpec(TRANSIENT_MASS, rho))
    sol = np.zeros(num.n_dof)
    xu = dof_coordinates(mesh, num, "U", num.n_dof)
    ess = np.arange(num.n_dof) >= num.n_inc
    sol[ess] = s_exact(field_id, xu[ess][:, :dim])          # interior initialised to zero, boundary to the field
    pb.sol = sol
    pb.meta = dict(kind="diffusion", quad_degree=quad_degree, field=field_id, mu=k, transient=transient)
    if build_pattern:
        pb.build_pattern()
    return pb


@dataclass
class ChnsModel:
    """Parameters of the CHNS weak form (CHNS_Solver, src/CHNS_Solver.cpp:236-420; feSysElm CHNS_Abels,
    src/feSysElm.h:1269-1345).  Property laws as device enums: linear mixing in phi with optional clipping
    (density_f / densityLimiter_f / viscosity_f / viscosityLimiter_f), constant or degenerate mobility
    (degenerateMobility_f), src/CHNS_Solver.cpp:124-235."""
    rhoA: float = 1.0
    rhoB: float = 1.0
    viscA: float = 1.0
    viscB: float = 1.0
    mobility: float = 1.0
    sigma: float = 1.0
    epsilon: float = 0.1
    force: tuple = (0.0, 0.0)
    src_u: tuple = (0.0, 0.0)
    src_p: float = 0.0
    src_phi: float = 0.0
    src_mu: float = 0.0
    limiter: bool = False
    degenerate_mobility: bool = False
    phi_order: int = 1
    # "abels" = CHNS_Abels<2>; "mass_averaged" = CHNS_MassAveraged<2> with alpha = (rho_2 - rho_1) / (rho_1 + rho_2)
    # (src/feSysElm.h:1352-1430, src/CHNS_Solver.cpp:398-416)
    # "khanwale" = CHNS_Khanwale<2> with khanwale = (Re, Pe, Cn, We, Fr, rhoA, rhoB) (src/feSysElm.h:1434-1512,
    # src/CHNS_Solver.cpp:418-446)
    formulation: str = "abels"
    alpha: float = 0.0
    khanwale: tuple = (1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0)


def phi_init(x):
    """numpy twin of chnsPhiCb (oracle/ref_harness.cpp)"""
    return 1.2 * np.cos(np.pi * x[..., 0]) * np.cos(np.pi * x[..., 1])


def mu_init(x):
    """numpy twin of chnsMuCb (oracle/ref_harness.cpp)"""
    return 0.3 * np.sin(np.pi * x[..., 0]) * np.sin(2. * np.pi * x[..., 1]) + 0.1 * x[..., 0]


def chns(mesh: Mesh, model: ChnsModel | None = None, quad_degree: int = 8, field_id: int = 1, mu: float = 1.0,
         rho: float = 1.0, build_pattern: bool = True) -> HostProblem:
    """Monolithic CHNS system [U (P2 vector), P (P1), Phi, Mu (P1 or P2)] with the single CHNS_ABELS weak form, as
    CHNS_Solver sets it up (src/CHNS_Solver.cpp:236-420): velocity essential on the boundary, pressure pinned at
    `PointPression` when the mesh has one, natural conditions for Phi and Mu."""
    assert mesh.dim == 2, "the reference instantiates its CHNS forms for dim = 2 only (src/feSysElmCHNS.cpp:275)"
    model = model or ChnsModel()
    dim, fo = 2, model.phi_order
    sp = [NB.SpaceSpec("U", "domain", 2, dim), NB.SpaceSpec("U", "boundary", 2, dim, True),
          NB.SpaceSpec("P", "domain", 1, 1), NB.SpaceSpec("Phi", "domain", fo, 1), NB.SpaceSpec("Mu", "domain", fo, 1)]
    if mesh.point_pressure is not None:
        sp.append(NB.SpaceSpec("P", "point", 0, 1, True))
    num = NB.build_numbering(mesh, sp)
    w, q = T.quadrature(dim, quad_degree)
    LU, dLU = T.basis(dim, 2, q)
    LP, dLP = T.basis(dim, 1, q)
    LF, dLF = T.basis(dim, fo, q)
    pb = HostProblem(mesh, dim, dim, 2, num, num.adr(mesh, "U", 2), num.adr(mesh, "P", 1), num.n_inc, num.n_dof,
                     w, q, LU, dLU, LP, dLP)
    pb.adrF, pb.adrM, pb.LF, pb.dLF, pb.chns = num.adr(mesh, "Phi", fo), num.adr(mesh, "Mu", fo), LF, dLF, model
    pb.forms = []
    sol = np.zeros(num.n_dof)
    xu = dof_coordinates(mesh, num, "U", num.n_dof)
    cu = dof_components(num, "U", num.n_dof)
    isU = cu >= 0
    uval = u_exact(field_id, xu[isU], mu, rho)
    sol[isU] = uval[np.arange(uval.shape[0]), cu[isU]]
    for fld, fn in (("P", lambda x: p_exact(field_id, x, mu, rho)), ("Phi", phi_init), ("Mu", mu_init)):
        xf = dof_coordinates(mesh, num, fld, num.n_dof)
        ok = dof_components(num, fld, num.n_dof) >= 0
        sol[ok] = fn(xf[ok])
    pb.sol = sol
    pb.meta = dict(kind="chns_abels", quad_degree=quad_degree, field=field_id, mu=mu, rho=rho)
    if build_pattern:
        pb.build_pattern()
    return pb


def perturb_unknowns(pb: HostProblem, amplitude: float = 1e-2, seed: int = 20261017) -> np.ndarray:
    """Analytic state + i.i.d. uniform noise on every unknown DOF, so that the convective
    Jacobian has no accidental zeros."""
    rng = np.random.default_rng(seed)
    sol = pb.sol.copy()
    sol[:pb.n_inc] += rng.uniform(-amplitude, amplitude, pb.n_inc)
    return sol
