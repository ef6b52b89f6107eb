"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's monolithic Cahn-Hilliard Navier-Stokes weak form
(volume-averaged velocity, Abels et al.) and of its finite-difference Jacobian.

  * residual: CHNS_Abels<2>::computeBe, src/feSysElmCHNS.cpp:66-273 (fields [U, P, Phi, Mu], layout :13-14;
    lambda = 3/(2 sqrt 2) sigma epsilon, src/feSysElm.h:1338);
  * Jacobian: feBilinearForm::computeMatrixFiniteDifference, src/feBilinearForm.cpp:388-428 (forward differences,
    h0 = sqrt(DBL_EPSILON) (:170), delta = h0 max(|u_j|, 1), solDot perturbed by delta c0, Ae = -(Rh - R0)/delta);
  * property laws: src/CHNS_Solver.cpp:124-235 (linear mixing in phi, optional clipping of phi to [-1, 1], constant or
    degenerate mobility).

Pinned element by element on the compiled reference (tests/test_oracle_vs_reference.py::test_chns_*) and by the
committed fixture tests/golden/ref_square1_chns_abels.npz.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this module.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .fe_oracle import geometry, phys_grad

CHNS_ABELS = 35          # elementSystemType, src/feSysElm.h:59
CHNS_MASS_AVERAGED = 36  # src/feSysElm.h:60
CHNS_KHANWALE = 38       # src/feSysElm.h:62
H0 = float(np.sqrt(np.finfo(np.float64).eps))


@dataclass
class ChnsParams:
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
    formulation: str = "abels"     # or "mass_averaged" (CHNS_MassAveraged<2>, src/feSysElmCHNS.cpp:347-602),
    #                                   "khanwale" (CHNS_Khanwale<2>, src/feSysElmCHNS.cpp:678-938)
    alpha: float = 0.0             # CHNS_MassAveraged: (rho_2 - rho_1) / (rho_1 + rho_2), src/feSysElm.h:1419
    khanwale: tuple = (1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0)   # Re, Pe, Cn, We, Fr, rhoA, rhoB (src/feSysElm.h:1501-1507)

    def as_array(self):
        return np.array([self.rhoA, self.rhoB, self.viscA, self.viscB, self.mobility, self.sigma, self.epsilon,
                         self.force[0], self.force[1], self.src_u[0], self.src_u[1], self.src_p, self.src_phi,
                         self.src_mu, float(self.limiter), float(self.degenerate_mobility), float(self.phi_order),
                         {"abels": 0.0, "mass_averaged": 1.0, "khanwale": 2.0}[self.formulation], self.alpha,
                         *self.khanwale])

    @property
    def lam(self):
        return 3.0 / (2.0 * np.sqrt(2.0)) * self.sigma * self.epsilon


@dataclass
class ChnsProblem:
    dim: int
    xyz: np.ndarray
    cells: np.ndarray
    adr: list                       # [adrU (nE, nSU*dim), adrP, adrPhi, adrMu]
    w: np.ndarray
    L: list                         # scalar basis tables [LU, LP, LF, LM], (nq, nS)
    dL: list                        # (nq, nS, dim)
    n_inc: int
    prm: ChnsParams = field(default_factory=ChnsParams)


def residual(pb: ChnsProblem, geo, loc, dot, prm: ChnsParams, phi_n_loc=None, loc_n=None, dt=0.0):
    """Be[e, M] of CHNS_Abels / CHNS_MassAveraged / CHNS_Khanwale on every element; loc = [U (nE, nSU, d), P (nE, nSP),
    Phi, Mu], dot likewise (P, Mu entries unused); phi_n_loc = Phi DOFs at the previous time step (mass-averaged form),
    loc_n = all local DOFs at the previous time step and dt = time step (Khanwale form)."""
    d = pb.dim
    LU, LP, LF, LM = pb.L
    jw = geo.detJ[:, None] * pb.w[None, :]
    gU, gF, gM = phys_grad(pb.dL[0], geo), phys_grad(pb.dL[2], geo), phys_grad(pb.dL[3], geo)
    U, P, F, Mu = loc
    u = np.einsum("ka,eac->ekc", LU, U)
    p = np.einsum("kq,eq->ek", LP, P)
    phi = np.einsum("kq,eq->ek", LF, F)
    mu = np.einsum("kq,eq->ek", LM, Mu)
    dudt = np.einsum("ka,eac->ekc", LU, dot[0])
    dphidt = np.einsum("kq,eq->ek", LF, dot[2])
    gu = np.einsum("ekam,eac->ekmc", gU, U)                   # gu[m, n] = d_m u_n
    gphi = np.einsum("ekam,ea->ekm", gF, F)
    gmu = np.einsum("ekam,ea->ekm", gM, Mu)
    pc = np.clip(phi, -1.0, 1.0) if prm.limiter else phi
    rho = (prm.rhoA - prm.rhoB) / 2.0 * pc + (prm.rhoA + prm.rhoB) / 2.0
    drho = (prm.rhoA - prm.rhoB) / 2.0
    eta = (prm.viscA - prm.viscB) / 2.0 * pc + (prm.viscA + prm.viscB) / 2.0
    Mob = prm.mobility * np.abs(1.0 - phi * phi) if prm.degenerate_mobility else prm.mobility * np.ones_like(phi)
    f = np.asarray(prm.force, float)[:d]
    Su = np.asarray(prm.src_u, float)[:d]
    ugu = np.einsum("ekn,eknc->ekc", u, gu)
    gmgu = np.einsum("ekn,eknc->ekc", gmu, gu)
    S = gu + np.swapaxes(gu, 2, 3)
    divu = np.einsum("ekmm->ek", gu)
    ugphi = np.einsum("ekm,ekm->ek", u, gphi)
    lam = prm.lam
    if prm.formulation == "khanwale":
        return _residual_khanwale(pb, geo, prm, jw, loc, dot, loc if loc_n is None else loc_n, dt)
    if prm.formulation == "mass_averaged":
        return _residual_mass_averaged(pb, geo, prm, jw, (gU, gF, gM), (u, p, phi, mu, dudt, dphidt, gu, gphi, gmu),
                                       (rho, drho, eta, Mob, f, Su), (ugu, S, divu, ugphi), loc, phi_n_loc)
    # momentum: test function i = a*d + c
    vec = (rho[..., None] * (dudt + ugu - f) - (drho * Mob)[..., None] * gmgu + phi[..., None] * gmu + Su)   # . phi_a e_c
    Bu = np.einsum("ekc,ka,ek->eac", vec, LU, jw)
    Bu += np.einsum("ek,ekac,ek->eac", -p, gU, jw)                            # - p div(phi_i)
    Bu += np.einsum("ek,ekam,ekmc,ek->eac", eta, gU, S, jw)                   # eta S : grad(phi_i)
    Bp = np.einsum("ek,kq,ek->eq", divu + prm.src_p, LP, jw)
    Bf = np.einsum("ek,kq,ek->eq", dphidt + ugphi + prm.src_phi, LF, jw) + np.einsum("ek,ekm,ekqm,ek->eq", Mob, gmu, gF, jw)
    Bm = np.einsum("ek,kq,ek->eq", mu - lam / prm.epsilon ** 2 * phi * (phi * phi - 1.0) + prm.src_mu, LM, jw) \
        - lam * np.einsum("ekm,ekqm,ek->eq", gphi, gM, jw)
    nE = U.shape[0]
    return -np.concatenate([Bu.reshape(nE, -1), Bp, Bf, Bm], 1)


def _residual_mass_averaged(pb, geo, prm, jw, grads, flds, props, derived, loc, phi_n_loc):
    """CHNS_MassAveraged<2>::computeBe, src/feSysElmCHNS.cpp:347-602 (tau = lambda, beta = 3/(2 sqrt 2) sigma / epsilon,
    src/feSysElm.h:1425-1426)."""
    LU, LP, LF, LM = pb.L
    gU, gF, gM = grads
    u, p, phi, mu, dudt, dphidt, gu, gphi, gmu = flds
    rho, drho, eta, Mob, f, Su = props
    ugu, S, divu, ugphi = derived
    gP = phys_grad(pb.dL[1], geo)
    gp = np.einsum("ekam,ea->ekm", gP, loc[1])
    phi_n = np.einsum("kq,eq->ek", LF, loc[2] if phi_n_loc is None else phi_n_loc)
    alpha, tau = prm.alpha, prm.lam
    beta = 3.0 / (2.0 * np.sqrt(2.0)) * prm.sigma / prm.epsilon
    div_rho_u = rho * divu + drho * ugphi                                     # :458-461
    phi_avg = 0.5 * (phi + phi_n)
    well = (phi * (phi * phi - 1.0) + 4.0 * phi_avg * (phi_avg * phi_avg - 1.0) + phi_n * (phi_n * phi_n - 1.0)) * beta / 6.0
    # momentum (:486-513)
    vec = rho[..., None] * (dudt + ugu - f) + (0.5 * (drho * dphidt + div_rho_u))[..., None] * u + phi[..., None] * gmu + Su
    Bu = np.einsum("ekc,ka,ek->eac", vec, LU, jw)
    Bu += np.einsum("ek,ekac,ek->eac", -(p + eta * (2.0 / pb.dim) * divu), gU, jw)   # - p div(phi_i) - eta 2/dim div u div(phi_i)
    Bu += np.einsum("ek,ekam,ekmc,ek->eac", eta, gU, S, jw)
    flux = Mob[..., None] * (gmu + alpha * gp)
    # continuity (:518-541), tracer (:546-574), potential (:579-600)
    Bp = np.einsum("ek,kq,ek->eq", divu + prm.src_p, LP, jw) + alpha * np.einsum("ekm,ekqm,ek->eq", flux, gP, jw)
    Bf = np.einsum("ek,kq,ek->eq", dphidt + prm.src_phi, LF, jw) - np.einsum("ek,ekm,ekqm,ek->eq", phi, u, gF, jw) \
        + np.einsum("ekm,ekqm,ek->eq", flux, gF, jw)
    Bm = np.einsum("ek,kq,ek->eq", mu - well + prm.src_mu, LM, jw) - tau * np.einsum("ekm,ekqm,ek->eq", gphi, gM, jw)
    nE = loc[0].shape[0]
    return -np.concatenate([Bu.reshape(nE, -1), Bp, Bf, Bm], 1)


def _residual_khanwale(pb, geo, prm, jw, loc, dot, loc_n, dt):
    """CHNS_Khanwale<2>::computeBe, src/feSysElmCHNS.cpp:678-938: non-dimensional form on fields averaged with the
    previous time step; the volume force is the constant (0, -1) of :623."""
    d = pb.dim
    LU, LP, LF, LM = pb.L
    gU, gP, gF, gM = (phys_grad(pb.dL[i], geo) for i in range(4))
    Re, Pe, Cn, We, Fr, rA, rB = prm.khanwale

    def fields(l):
        U, P, F, Mu = l
        return (np.einsum("ka,eac->ekc", LU, U), np.einsum("kq,eq->ek", LP, P), np.einsum("kq,eq->ek", LF, F),
                np.einsum("kq,eq->ek", LM, Mu), np.einsum("ekam,eac->ekmc", gU, U), np.einsum("ekam,ea->ekm", gF, F),
                np.einsum("ekam,ea->ekm", gM, Mu))
    u, p, phi, mu, gu, gphi, gmu = fields(loc)
    un, pn, phin, mun, gun, gphin, gmun = fields(loc_n)
    ua, pa, fa, ma = 0.5 * (u + un), 0.5 * (p + pn), 0.5 * (phi + phin), 0.5 * (mu + mun)
    gua, gfa, gma = 0.5 * (gu + gun), 0.5 * (gphi + gphin), 0.5 * (gmu + gmun)
    dudt = np.einsum("ka,eac->ekc", LU, dot[0])
    dphidt = np.einsum("kq,eq->ek", LF, dot[2])

    def lin(x, a, b):
        c = np.clip(x, -1.0, 1.0) if prm.limiter else x
        return (a - b) / 2.0 * c + (a + b) / 2.0
    rho_n, rho_a, eta_a, rho = lin(phin, prm.rhoA, prm.rhoB), lin(fa, prm.rhoA, prm.rhoB), lin(fa, prm.viscA, prm.viscB), \
        lin(phi, prm.rhoA, prm.rhoB)
    drho = (prm.rhoA - prm.rhoB) / 2.0
    well = fa * (fa * fa - 1.0)
    divu, divua = np.einsum("ekmm->ek", gu), np.einsum("ekmm->ek", gua)
    ugu = np.einsum("ekn,eknc->ekc", ua, gua)
    S = gua + np.swapaxes(gua, 2, 3)
    jflux = (rB - rA) / (2.0 * rA * Cn) * gma
    jgu = np.einsum("ekn,eknc->ekc", jflux, gua)
    gg = gfa[..., :, None] * gfa[..., None, :]
    div_rau = drho * np.einsum("ekm,ekm->ek", gfa, ua) + rho_a * divua
    f = np.array([0.0, -1.0])[:d]
    Su = np.asarray(prm.src_u, float)[:d]
    vec = rho_a[..., None] * (dudt + ugu) + jgu / Pe - rho_a[..., None] * f / Fr + Su
    Bu = np.einsum("ekc,ka,ek->eac", vec, LU, jw)
    Bu += np.einsum("ekam,ekmc,ek->eac", gU, gg, jw) * (-Cn / We)
    Bu += np.einsum("ek,ekac,ek->eac", -pa / We, gU, jw)
    Bu += np.einsum("ek,ekam,ekmc,ek->eac", eta_a / Re, gU, S, jw)
    Bp = np.einsum("ek,kq,ek->eq", divu + (rho - rho_n) / dt + div_rau + prm.src_p, LP, jw) \
        - np.einsum("ekm,ekqm,ek->eq", jflux, gP, jw) / Pe
    Bf = np.einsum("ek,kq,ek->eq", dphidt + prm.src_phi, LF, jw) - np.einsum("ek,ekm,ekqm,ek->eq", fa, ua, gF, jw) \
        + np.einsum("ekm,ekqm,ek->eq", gma, gF, jw) / (Pe * Cn)
    Bm = np.einsum("ek,kq,ek->eq", ma - well + prm.src_mu, LM, jw) - Cn * Cn * np.einsum("ekm,ekqm,ek->eq", gfa, gM, jw)
    nE = loc[0].shape[0]
    return -np.concatenate([Bu.reshape(nE, -1), Bp, Bf, Bm], 1)


def element_systems(pb: ChnsProblem, sol, soldot=None, c0=0.0, matrix=True, sol_n=None, dt=0.0):
    """(Ae[e, M, M] by finite differences or None, Be[e, M], adr[e, M]).  sol_n: state at the previous time step (the
    reference's global solAtTimeN, never perturbed by the finite differences); None = the current solution."""
    geo = geometry(pb.xyz, pb.cells, pb.dim)
    nE = pb.cells.shape[0]
    d = pb.dim
    if soldot is None:
        soldot = np.zeros_like(sol)
    shapes = [(nE, pb.L[0].shape[1], d), None, None, None]

    def gather(vec):
        out = []
        for s, a in enumerate(pb.adr):
            v = vec[a]
            out.append(v.reshape(shapes[0]) if s == 0 else v)
        return out
    loc, dot = gather(sol), gather(soldot)
    phi_n = (sol if sol_n is None else sol_n)[pb.adr[2]].copy()
    loc_n = [v.copy() for v in gather(sol if sol_n is None else sol_n)]
    R0 = residual(pb, geo, loc, dot, pb.prm, phi_n, loc_n, dt)
    adr = np.concatenate(pb.adr, 1)
    if not matrix:
        return None, R0, adr
    M = R0.shape[1]
    Ae = np.zeros((nE, M, M))
    col = 0
    for s in range(4):
        flat = loc[s].reshape(nE, -1)
        flatd = dot[s].reshape(nE, -1)
        for j in range(flat.shape[1]):
            t, td = flat[:, j].copy(), flatd[:, j].copy()
            delta = H0 * np.maximum(np.abs(t), 1.0)
            flat[:, j] = t + delta
            flatd[:, j] = td + delta * c0
            Rh = residual(pb, geo, loc, dot, pb.prm, phi_n, loc_n, dt)
            Ae[:, :, col] = -(Rh - R0) * (1.0 / delta)[:, None]
            flat[:, j] = t
            flatd[:, j] = td
            col += 1
    return Ae, R0, adr


def assemble(pb: ChnsProblem, ia, ja, sol, soldot=None, c0=0.0, matrix=True, residual_=True, sol_n=None, dt=0.0):
    """Global CSR values and rhs (scatter of src/feLinearSystemMklPardiso.cpp:524-663, :699-741)."""
    Ae, Be, adr = element_systems(pb, sol, soldot, c0, matrix, sol_n, dt)
    n = np.int64(pb.n_inc)
    vals = np.zeros(ja.shape[0])
    rhs = np.zeros(pb.n_inc)
    aI = adr.astype(np.int64)
    if residual_:
        ok = aI < n
        np.add.at(rhs, aI[ok], Be[ok])
    if matrix:
        key = np.repeat(np.arange(pb.n_inc, dtype=np.int64), np.diff(ia)) * n + ja.astype(np.int64)
        I = aI[:, :, None]
        J = aI[:, None, :]
        ok = (I < n) & (J < n)
        pos = np.searchsorted(key, (I * n + J)[ok])
        np.add.at(vals, pos, Ae[ok])
    return vals, rhs
