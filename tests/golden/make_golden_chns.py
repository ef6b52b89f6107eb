#!/usr/bin/env python
"""CHNS fixtures (tests/golden/ref_square1_chns_abels_p{1,2}.npz, ref_square1_chns_mass_averaged_p1.npz, ref_square1_chns_khanwale_p1.npz) from the
UNMODIFIED reference compiled into oracle/_ref: inputs exactly as the reference holds them on its own regression mesh
data/square1.msh, and the outputs of its own CPU path (CHNS_Abels<2> / CHNS_MassAveraged<2> / CHNS_Khanwale<2>::computeBe
+ computeMatrixFiniteDifference + the Pardiso-style scatter; the time-averaged fixtures have solAtTimeN != sol).

    python tests/golden/make_golden_chns.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import chns_oracle as CO, ref  # noqa: E402

for formulation, fo in (("abels", 1), ("abels", 2), ("mass_averaged", 1), ("khanwale", 1)):
    prm = CO.ChnsParams(rhoA=1.3, rhoB=0.7, viscA=0.05, viscB=0.02, mobility=0.01, sigma=0.5, epsilon=0.07,
                        force=(0.1, -0.98), src_u=(0.2, 0.3), src_p=0.1, src_phi=-0.2, src_mu=0.4, limiter=True,
                        degenerate_mobility=True, phi_order=fo, formulation=formulation,
                        alpha=-0.3 if formulation == "mass_averaged" else 0.0,
                        khanwale=(40., 25., 0.07, 3.0, 1.7, 1.3, 0.7) if formulation == "khanwale" else (1.,) * 7)
    P = ref.RefProblem("/root/reference/data/square1.msh", "chns", 2, 8, 1, 0.05, 1.3, chns=prm.as_array())
    xyz, conn = P.mesh()
    out = dict(kind="chns_" + formulation, dim=2, xyz=xyz, cells=conn, n_dof=P.n_dof, n_inc=P.n_inc, chns_params=prm.as_array())
    for s in range(4):
        out[f"adr{s}"] = P.adr(s)
        L, dr, ds, _ = P.tables(s)
        out[f"L{s}"], out[f"dLdr{s}"], out[f"dLds{s}"] = L, dr, ds
    w, qr, qs, _ = P.quadrature()
    out.update(w=w, qr=qr, qs=qs)
    ia, ja = P.pattern()
    out.update(ia=ia, ja=ja)
    sol0, _ = P.solution()
    rng = np.random.default_rng(20261017)
    sol = sol0.copy()
    sol[:P.n_inc] += rng.uniform(-1e-2, 1e-2, P.n_inc)
    sd = rng.standard_normal(P.n_dof)
    c0 = 3.5
    P.set_solution(sol, sd, c0, 0.0)
    if formulation != "abels":
        sol_n = sol + rng.uniform(-5e-2, 5e-2, P.n_dof)
        P.set_solution_n(sol_n, 0.02)
        out["sol_n"] = sol_n
        out["dt"] = 0.02
    elements = np.arange(0, P.n_elm, 5)
    Ae, Be = [], []
    for e in elements:
        a, b, _, _ = P.element(0, int(e))
        Ae.append(a)
        Be.append(b)
    v, r, _ = P.assemble()
    # the reference's OWN finite-difference noise: the same Jacobian from a state that differs by one unit in the last
    # place in every unknown (computeMatrixFiniteDifference divides residual differences by delta = sqrt(eps) max(|u|, 1),
    # src/feBilinearForm.cpp:404-422, so 1e-16-level differences of the residual come back multiplied by 6.7e7)
    sol_ulp = sol.copy()
    sol_ulp[:P.n_inc] = np.nextafter(sol[:P.n_inc], np.inf)
    P.set_solution(sol_ulp, sd, c0, 0.0)
    v_ulp, _, _ = P.assemble()
    P.set_solution(sol, sd, c0, 0.0)
    rowmax = np.maximum.reduceat(np.abs(v), ia[:-1])
    fd_noise = float((np.maximum.reduceat(np.abs(v_ulp - v), ia[:-1]) / rowmax).max())
    out["fd_noise"] = fd_noise
    print("   reference FD noise (1-ulp state change), relative to the row maximum: %.3e" % fd_noise)
    out.update(sol_init=sol0, sol=sol, sol_dot=sd, c0=c0, elements=elements, Ae=np.array(Ae), Be=np.array(Be), values=v,
               rhs=r)
    path = os.path.join(HERE, f"ref_square1_chns_{formulation}_p{fo}.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes; nElm", P.n_elm, "nInc", P.n_inc, "nnz", P.nnz)
    P.close()
