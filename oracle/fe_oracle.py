"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's element kernels and scatter.

CPU oracle for the hot path: per-element quadrature loops of feSysElm::computeAe/computeBe,
the gather of feBilinearForm::initialize, and the colour-ordered scatter of feLinearSystemMklPardiso, restated
form by form with the reference's own per-form accumulation (one form at a time, quadrature-point major) and
generalised from dim=2 to dim=3 for the vector forms (the reference instantiates them for <2> only,
src/feVectorSysElm.cpp:132,240,507,582,753,1246,1536).

Pinned (tests/test_oracle_vs_reference.py, tests/golden/):
  * dim=2 vector forms and dim=2/3 scalar forms agree with the compiled reference (oracle/_ref) element by element;
  * committed golden fixtures generated from the compiled reference (tests/golden/make_golden.py).
The dim=3 vector forms have no reference implementation: parity for them is "restatement only", cross-checked by
(i) the dim=2 agreement of the same code path, (ii) their scalar sub-blocks against feSysElm_Diffusion<3>,
(iii) finite-difference Jacobian consistency (tests/test_oracle_fd.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# elementSystemType values of the reference (src/feSysElm.h:13-77)
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


@dataclass
class Form:
    kind: int
    coeff: float = 1.0          # the form's scalar coefficient callback, constant here
    param: float = 1.0          # viscosity / diffusivity where the form has one
    source: np.ndarray | None = None   # tabulated source at (e, k[, c]) for SOURCE / VECTOR_SOURCE


@dataclass
class Geometry:
    """Per-element affine map of straight simplices.  detJ as src/feCncGeo.cpp:332,340,385; inverse map as
    computeElementTransformation, src/feCncGeo.cpp:651-692: G[e, alpha, m] = d(xi_alpha)/d(x_m)."""
    detJ: np.ndarray
    G: np.ndarray


def geometry(xyz: np.ndarray, cells: np.ndarray, dim: int) -> Geometry:
    X = xyz[cells][:, :, :dim]                                   # (nE, nv, dim)
    # dx_m/dxi_alpha for the P1 geometric interpolant: columns = edge vectors from vertex 0
    F = np.stack([X[:, a + 1] - X[:, 0] for a in range(dim)], 2)  # F[e, m, alpha]
    if dim == 2:
        dxdr, dydr, dxds, dyds = F[:, 0, 0], F[:, 1, 0], F[:, 0, 1], F[:, 1, 1]
        J = dxdr * dyds - dydr * dxds
        G = np.empty_like(F)
        G[:, 0, 0] = dyds / J      # drdx[0]
        G[:, 1, 0] = -dydr / J     # drdx[1]
        G[:, 0, 1] = -dxds / J     # drdy[0]
        G[:, 1, 1] = dxdr / J      # drdy[1]
    else:
        J = np.linalg.det(F)
        G = np.linalg.inv(F)       # G[e, alpha, m]
    return Geometry(J, G)


def phys_grad(dL: np.ndarray, geo: Geometry) -> np.ndarray:
    """grad[e, k, a, m] = sum_alpha dL[k, a, alpha] * G[e, alpha, m]   (src/feSpace.cpp:669-710)"""
    return np.einsum("kaq,eqm->ekam", dL, geo.G)


def element_forms(form: Form, dim: int, geo: Geometry, w: np.ndarray, LU, dLU, LP=None, uloc=None, ploc=None,
                  udotloc=None, c0: float = 0.0):
    """Local matrix Ae[e, M, N] (None if the form has no matrix) and residual Be[e, M] of ONE form on every element.

    LU/dLU: scalar basis tables of the primary (velocity or scalar) space; LP: pressure basis table.
    uloc: (nE, nS, nc) local primary DOFs; ploc: (nE, nP); udotloc like uloc.
    Vector local index i = a*dim + c, (phi_i)_n = phi_a delta_cn   (src/feSpace_2D.cpp:41-55, :916-928).
    """
    nE = geo.detJ.shape[0]
    nS = LU.shape[1]
    jw = geo.detJ[:, None] * w[None, :]                           # (nE, nq)
    g = phys_grad(dLU, geo)                                       # (nE, nq, nS, dim)
    k = form.kind
    c = form.coeff
    d = dim
    eye = np.eye(d)

    if k == DIFFUSION:            # src/feSysElm.cpp:530-583
        Ae = form.param * np.einsum("ekam,ekbm,ek->eab", g, g, jw)
        gu = np.einsum("ekam,ea->ekm", g, uloc[:, :, 0])
        Be = -form.param * np.einsum("ekam,ekm,ek->ea", g, gu, jw)
        return Ae, Be
    if k == SOURCE:               # src/feSysElm.cpp:16-27
        return None, -np.einsum("ka,ek,ek->ea", LU, form.source, jw)
    if k == TRANSIENT_MASS:       # src/feSysElm.cpp:426-460
        Ae = c * c0 * np.einsum("ka,kb,ek->eab", LU, LU, jw)
        ud = np.einsum("ka,ea->ek", LU, udotloc[:, :, 0])
        return Ae, -c * np.einsum("ka,ek,ek->ea", LU, ud, jw)

    # ---- vector forms --------------------------------------------------------------------------------
    u = None if uloc is None else np.einsum("ka,eac->ekc", LU, uloc)          # (nE, nq, d)
    gu = None if uloc is None else np.einsum("ekam,eac->ekmc", g, uloc)       # gu[m, n] = d_m u_n  (src/feSpace.cpp:1391-1394)

    def vec(A4):   # (nE, nS, d, nS, d) -> (nE, nS*d, nS*d)
        return A4.reshape(nE, nS * d, nS * d)

    if k == VECTOR_CONVECTIVE_ACCELERATION:      # src/feVectorSysElm.cpp:1171-1242
        ugp = np.einsum("ekm,ekbm->ekb", u, g)                                     # u . grad(phi_b)
        A1 = np.einsum("ekb,ka,ek->eab", ugp, LU, jw)                              # (u.grad phi_b) phi_a  delta_ij
        A2 = np.einsum("kb,ka,ekji,ek->eaibj", LU, LU, gu, jw)                     # phi_b phi_a d_j u_i
        A = np.einsum("eab,ij->eaibj", A1, eye) + A2
        ugu = np.einsum("ekn,eknm->ekm", u, gu)                                    # (u.grad u)_m = u_n d_n u_m
        Be = -c * np.einsum("ekc,ka,ek->eac", ugu, LU, jw).reshape(nE, nS * d)
        return c * vec(A), Be
    if k == DIV_NEWTONIAN_STRESS:                # src/feVectorSysElm.cpp:1454-1532
        mu = form.param
        nP = LP.shape[1]
        gg = np.einsum("ekam,ekbm,ek->eab", g, g, jw)
        A1 = np.einsum("eab,ij->eaibj", gg, eye)
        A2 = np.einsum("ekaj,ekbi,ek->eaibj", g, g, jw)                            # d_j phi_a d_i phi_b
        Auu = -c * mu * vec(A1 + A2)
        Aup = c * np.einsum("kq,ekai,ek->eaiq", LP, g, jw).reshape(nE, nS * d, nP)
        p = np.einsum("kq,eq->ek", LP, ploc)
        S = gu + np.swapaxes(gu, 2, 3)
        Be = -c * (np.einsum("ek,ekai,ek->eai", p, g, jw) - mu * np.einsum("ekam,ekmi,ek->eai", g, S, jw))
        return np.concatenate([Auu, Aup], 2), Be.reshape(nE, nS * d)
    if k == MIXED_DIVERGENCE:                    # src/feVectorSysElm.cpp:685-749  (rows P, cols U)
        Ae = c * np.einsum("kq,ekbj,ek->eqbj", LP, g, jw).reshape(nE, LP.shape[1], nS * d)
        divu = np.einsum("ekmm->ek", gu)
        return Ae, -c * np.einsum("ek,kq,ek->eq", divu, LP, jw)
    if k == VECTOR_DIFFUSION:                    # src/feVectorSysElm.cpp:449-503
        kk = form.param
        gg = np.einsum("ekam,ekbm,ek->eab", g, g, jw)
        Ae = c * kk * vec(np.einsum("eab,ij->eaibj", gg, eye))
        Be = -c * kk * np.einsum("ekam,ekmi,ek->eai", g, gu, jw).reshape(nE, nS * d)
        return Ae, Be
    if k == MIXED_GRADIENT:                      # src/feVectorSysElm.cpp:528-578  (rows U, cols P)
        Ae = -c * np.einsum("kq,ekai,ek->eaiq", LP, g, jw).reshape(nE, nS * d, LP.shape[1])
        p = np.einsum("kq,eq->ek", LP, ploc)
        return Ae, c * np.einsum("ek,ekai,ek->eai", p, g, jw).reshape(nE, nS * d)
    if k == VECTOR_SOURCE:                       # src/feVectorSysElm.cpp:112-128
        return None, -np.einsum("ekc,ka,ek->eac", form.source, LU, jw).reshape(nE, nS * d)
    if k == TRANSIENT_VECTOR_MASS:               # src/feVectorSysElm.cpp:390-425
        mm = np.einsum("ka,kb,ek->eab", LU, LU, jw)
        Ae = c * c0 * vec(np.einsum("eab,ij->eaibj", mm, eye))
        ud = np.einsum("ka,eac->ekc", LU, udotloc)
        return Ae, -c * np.einsum("ekc,ka,ek->eac", ud, LU, jw).reshape(nE, nS * d)
    raise ValueError(f"form kind {k} not restated")


def form_layout(kind: int):
    """(rows, cols) field layout of the form's local matrix: 'U' primary space, 'P' pressure space
    (createElementarySystem of each weak form, e.g. src/feVectorSysElm.cpp:1428-1433)."""
    if kind == DIV_NEWTONIAN_STRESS:
        return ("U",), ("U", "P")
    if kind == MIXED_DIVERGENCE:
        return ("P",), ("U",)
    if kind == MIXED_GRADIENT:
        return ("U",), ("P",)
    return ("U",), ("U",)


def has_matrix(kind: int) -> bool:
    return kind not in (SOURCE, VECTOR_SOURCE)


@dataclass
class Problem:
    """Everything the hot path needs, as flat tables (what the C ABI receives)."""
    dim: int
    xyz: np.ndarray
    cells: np.ndarray
    adrU: np.ndarray                 # (nE, nS*nc)
    adrP: np.ndarray | None          # (nE, nP)
    ncomp: int
    w: np.ndarray
    LU: np.ndarray
    dLU: np.ndarray
    LP: np.ndarray | None
    n_inc: int
    forms: list = field(default_factory=list)


def assemble(pb: Problem, ia, ja, sol, soldot=None, c0=0.0, matrix=True, residual=True):
    """Global CSR values and rhs, form by form then element by element (src/feLinearSystemMklPardiso.cpp:524-663,
    :699-741); essential DOFs (>= nInc) filtered (:565-579, :733)."""
    geo = geometry(pb.xyz, pb.cells, pb.dim)
    nE = pb.cells.shape[0]
    nS = pb.LU.shape[1]
    uloc = sol[pb.adrU].reshape(nE, nS, pb.ncomp)
    udot = None if soldot is None else soldot[pb.adrU].reshape(nE, nS, pb.ncomp)
    ploc = None if pb.adrP is None else sol[pb.adrP]
    adr = {"U": pb.adrU, "P": pb.adrP}
    vals = np.zeros(ja.shape[0])
    rhs = np.zeros(pb.n_inc)
    n = np.int64(pb.n_inc)
    key = np.repeat(np.arange(pb.n_inc, dtype=np.int64), np.diff(ia)) * n + ja.astype(np.int64)
    for f in pb.forms:
        Ae, Be = element_forms(f, pb.dim, geo, pb.w, pb.LU, pb.dLU, pb.LP, uloc, ploc, udot, c0)
        rows, cols = form_layout(f.kind)
        aI = np.concatenate([adr[r] for r in rows], 1).astype(np.int64)
        aJ = np.concatenate([adr[r] for r in cols], 1).astype(np.int64)
        if residual:
            ok = aI < n
            np.add.at(rhs, aI[ok], Be[ok])
        if matrix and Ae is not None:
            I = aI[:, :, None]
            J = aJ[:, None, :]
            ok = (I < n) & (J < n)
            pos = np.searchsorted(key, (I * n + J)[ok])
            np.add.at(vals, pos, Ae[ok])
    return vals, rhs


def constrain(ia, ja, vals, rhs, rows):
    """constrainEssentialComponents (src/feLinearSystemMklPardiso.cpp:1092-1114): for every constrained row r the
    column r is zeroed in all rows, the row is zeroed, the diagonal set to 1 and the rhs entry to 0."""
    vals, rhs = vals.copy(), rhs.copy()
    n = ia.shape[0] - 1
    flag = np.zeros(n, bool)
    flag[np.asarray(rows, np.int64)] = True
    row_of = np.repeat(np.arange(n), np.diff(ia))
    vals[flag[ja]] = 0.0
    vals[flag[row_of]] = 0.0
    vals[flag[row_of] & (ja == row_of)] = 1.0
    rhs[flag] = 0.0
    return vals, rhs
