#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ from the UNMODIFIED reference compiled into oracle/_ref
(oracle/Makefile).  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

Every fixture is an .npz with the INPUTS of the hot path exactly as the reference holds them (vertices,
connectivity, element->DOF tables, quadrature + basis tables, state vector) and the OUTPUTS of the reference's own
CPU path on them (CSR pattern, colours, Jacobian determinants, per-element Ae/Be of every weak form, assembled CSR
values + rhs, constrained system, Newton solution and the error norms the reference's tests print).  The GPU box has
no /root/reference: the tests read only these files.

Cases
  ref_square1_*  the reference's own regression mesh data/square1.msh (42 unstructured triangles) -- the mesh of
                 tests/withLinearSolver/navier_stokes_MMS.output:4,11, stokes_MMS.output and
                 convergenceLaplace.output:10; the Newton error norms are stored next to those printed goldens.
  ref_cube1_*    data/cube1.msh, scalar P2 diffusion on tetrahedra (exe/example1.cpp:167)
  syn_*          synthetic meshes of feng_b200.mesh written with feng_b200.mesh.write_msh and read back by the
                 reference reader: pins the host-side numbering / pattern code bit-exactly.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from feng_b200 import mesh as M, problems as PB  # noqa: E402
from oracle import ref  # noqa: E402

REFDATA = "/root/reference/data"

# printed goldens of the reference's own tests (6 significant digits, tests/tests.h:4)
PRINTED = {
    # tests/withLinearSolver/navier_stokes_MMS.output:4 and :11 (divergence formulation, square1)
    ("ns_div", "square1"): (1.730601e-03, 1.064397e-02),
    ("ns_div", "square2"): (2.102148e-04, 2.316230e-03),     # :5, :12
    # navier_stokes_MMS.output:19,26 (Laplacian formulation)
    ("ns_lap", "square1"): (1.729202e-03, 1.064192e-02),
    # tests/withLinearSolver/stokes_MMS.output:4,11 (divergence) and :19,26 (Laplacian)
    ("stokes_div", "square1"): (1.730482e-03, 1.054679e-02),
    ("stokes_lap", "square1"): (1.729104e-03, 1.055463e-02),
    # tests/withLinearSolver/convergenceLaplace.output:10 (P2 Poisson; second number unused)
    ("diffusion", "square1"): (3.209814e-03, 0.0),
}


def dump(name, P, kind, perturb=1e-2, seed=20261017, newton=False, sol_dot=False, c0=0.0, elements=None,
         extra=None):
    rng = np.random.default_rng(seed)
    xyz, conn = P.mesh()
    out = dict(kind=kind, dim=P.dim, xyz=xyz, cells=conn, n_dof=P.n_dof, n_inc=P.n_inc)
    out.update(RECIPE)
    nsp = P.n_spaces
    for s in range(nsp):
        out[f"adr{s}"] = P.adr(s)
        L, dr, ds, dt = P.tables(s)
        out[f"L{s}"], out[f"dLdr{s}"], out[f"dLds{s}"], out[f"dLdt{s}"] = L, dr, ds, dt
    w, r, s_, t = P.quadrature()
    out.update(w=w, qr=r, qs=s_, qt=t, detJ=P.jacobians(), colors=P.colors())
    ia, ja = P.pattern()
    out.update(ia=ia, ja=ja.astype(np.int32))
    sol0, _ = P.solution()
    sol = sol0.copy()
    sol[:P.n_inc] += rng.uniform(-perturb, perturb, P.n_inc)
    sd = rng.standard_normal(P.n_dof) if sol_dot else None
    P.set_solution(sol, sd, c0, 0.0)
    out.update(sol_init=sol0, sol=sol, c0=c0)
    if sd is not None:
        out["sol_dot"] = sd
    vals, rhs, _ = P.assemble()
    out.update(vals=vals, rhs=rhs)
    if elements is None:
        elements = sorted(set([0, 1, P.n_elm // 2, P.n_elm - 1]))
    out["elements"] = np.array(elements)
    finfo = []
    for f in range(P.n_forms):
        fi = P.form_info(f)
        finfo.append([fi.M, fi.N, int(fi.has_matrix), fi.sys_id, int(fi.transient)])
        Ae = np.zeros((len(elements), fi.M, fi.N))
        Be = np.zeros((len(elements), fi.M))
        aI = np.zeros((len(elements), fi.M), np.int64)
        aJ = np.zeros((len(elements), fi.N), np.int64)
        for n, e in enumerate(elements):
            Ae[n], Be[n], aI[n], aJ[n] = P.element(f, e)
        out[f"Ae{f}"], out[f"Be{f}"], out[f"adrI{f}"], out[f"adrJ{f}"] = Ae, Be, aI, aJ
    out["form_info"] = np.array(finfo)
    rows = P.constraint_rows()
    out["constraint_rows"] = rows
    cv, cr = P.constrain()
    out.update(vals_constrained=cv, rhs_constrained=cr)
    P.set_solution(sol, sd, c0, 0.0)
    P.assemble()
    if newton:
        P.set_solution(sol0, None, 0.0, 0.0)
        nsol, info = P.newton(1e-10, 1e-10, 10)
        out.update(newton_sol=nsol, newton_info=info, error_norms=P.error_norms(nsol))
    if extra:
        out.update(extra)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: nElm={P.n_elm} nDOF={P.n_dof} nInc={P.n_inc} nnz={P.nnz} -> {os.path.getsize(path) / 1e3:.1f} kB",
          (out.get("error_norms")))
    return out


RECIPE = {}
_RefProblem = ref.RefProblem


def RefProblemRec(mesh_file, kind, order=2, quad_degree=8, field=0, mu=1.0, rho=1.0, transient=False,
                  p_essential=True):
    """ref.RefProblem that also remembers its recipe for the fixture"""
    RECIPE.clear()
    RECIPE.update(order=order, quad_degree=quad_degree, field=field, mu=mu, rho=rho, transient=transient,
                  p_essential=p_essential)
    return _RefProblem(mesh_file, kind, order, quad_degree, field, mu, rho, transient, p_essential)


def main():
    assert ref.available(), "build oracle/_ref first (make -C oracle)"
    ref.set_threads(1)   # per-row summation order of the serial colour loop (src/feLinearSystemMklPardiso.cpp:537-663)
    # ---- the reference's own regression meshes -------------------------------------------------------
    sq1 = os.path.join(REFDATA, "square1.msh")
    for kind in ("ns_div", "ns_lap", "stokes_div", "stokes_lap"):
        P = RefProblemRec(sq1, kind, 2, 8, field=0, mu=1.0, rho=1.0, p_essential=True)
        o = dump(f"ref_square1_{kind}", P, kind, newton=True)
        if (kind, "square1") in PRINTED:
            eu, ep = PRINTED[(kind, "square1")]
            got = o["error_norms"]
            assert f"{got[0]:.6e}" == f"{eu:.6e}" and f"{got[1]:.6e}" == f"{ep:.6e}", (got, eu, ep)
        P.close()
    P = RefProblemRec(sq1, "ns_div", 2, 8, field=0, mu=0.05, rho=1.3, transient=True, p_essential=True)
    dump("ref_square1_ns_div_transient", P, "ns_div", sol_dot=True, c0=3.5)
    P.close()
    P = RefProblemRec(sq1, "diffusion", 2, 12, field=0, mu=1.0)
    o = dump("ref_square1_diffusion_p2", P, "diffusion", newton=True)
    assert f"{o['error_norms'][0]:.6e}" == f"{PRINTED[('diffusion', 'square1')][0]:.6e}"
    P.close()
    P = RefProblemRec(sq1, "diffusion", 1, 4, field=0, mu=0.7, rho=1.3, transient=True)
    dump("ref_square1_diffusion_p1_transient", P, "diffusion", sol_dot=True, c0=2.5)
    P.close()
    P = RefProblemRec(os.path.join(REFDATA, "cube1.msh"), "diffusion", 2, 4, field=0, mu=1.0)
    dump("ref_cube1_diffusion_p2", P, "diffusion")
    P.close()
    # second mesh of the printed NS golden: only the scalars (the mesh itself is not stored)
    P = RefProblemRec(os.path.join(REFDATA, "square2.msh"), "ns_div", 2, 8, field=0, mu=1.0, rho=1.0,
                       p_essential=True)
    nsol, info = P.newton(1e-10, 1e-10, 10)
    en = P.error_norms(nsol)
    eu, ep = PRINTED[("ns_div", "square2")]
    assert f"{en[0]:.6e}" == f"{eu:.6e}" and f"{en[1]:.6e}" == f"{ep:.6e}", (en, eu, ep)
    P.close()
    # ---- synthetic meshes through the reference reader -----------------------------------------------
    tmp = tempfile.mkdtemp()
    m = M.square_mesh(5)
    path = os.path.join(tmp, "t2d5.msh")
    M.write_msh(m, path)
    for kind, pe in (("ns_div", True), ("ns_div", False), ("stokes_lap", True)):
        P = RefProblemRec(path, kind, 2, 8, field=0, mu=0.5, rho=1.3, p_essential=pe)
        dump(f"syn_t2d5_{kind}_{'pbord' if pe else 'ppoint'}", P, kind, newton=pe,
             extra=dict(bfacets=m.bfacets, point_pressure=m.point_pressure, mu=0.5, rho=1.3, p_essential=pe))
        P.close()
    P = RefProblemRec(path, "diffusion", 2, 12, field=0, mu=0.7)
    dump("syn_t2d5_diffusion_p2", P, "diffusion", extra=dict(bfacets=m.bfacets, point_pressure=m.point_pressure,
                                                             mu=0.7))
    P.close()
    m3 = M.cube_mesh(2)
    path3 = os.path.join(tmp, "t3d2.msh")
    M.write_msh(m3, path3)
    for order, deg in ((2, 4), (1, 2)):
        P = RefProblemRec(path3, "diffusion", order, deg, field=0, mu=0.7, rho=1.3, transient=True)
        dump(f"syn_t3d2_diffusion_p{order}", P, "diffusion", sol_dot=True, c0=2.5,
             extra=dict(bfacets=m3.bfacets, point_pressure=m3.point_pressure, mu=0.7, rho=1.3, order=order, deg=deg))
        P.close()


if __name__ == "__main__":
    main()
