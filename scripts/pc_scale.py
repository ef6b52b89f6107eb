#!/usr/bin/env python
"""GMRES(30) + Schur/AMG on the T3D(n) / T2D(n) Navier-Stokes Jacobian of the bench state: iterations and time versus size.

    python scripts/pc_scale.py [--dim 3] [--n 16 32 48] [--pc 5] [--restart 30]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--n", type=int, nargs="+", default=[16, 32])
    ap.add_argument("--pc", type=int, default=5)
    ap.add_argument("--restart", type=int, default=30)
    ap.add_argument("--maxit", type=int, default=600)
    ap.add_argument("--kind", default="ns_div")
    ap.add_argument("--box", type=int, default=0, help="3-D: n x n x (box * n) cells on [0,1]^2 x [0, box] instead of the cube (the weak-scaling domain of "
                    "`box` GPUs, undecomposed)")
    args = ap.parse_args()
    from feng_b200 import mesh as M, problems as PB
    from feng_b200.linear_system import LinearSystemB200
    for n in args.n:
        m = M.cube_mesh(n) if args.dim == 3 else M.square_mesh(n)
        if args.dim == 3 and args.box > 1:
            m = M.box_mesh(n, n, n * args.box, float(args.box))
            m.point_pressure = 0
        pb = PB.taylor_hood(m, args.kind, 6 if args.dim == 3 else 8, 3 if args.dim == 3 else 1, 1. / 40., 1.0, with_source=False,
                            build_pattern=False)
        sol = PB.perturb_unknowns(pb)
        ls = LinearSystemB200(pb, device_pattern=True)
        S = ls.sys
        S.set_solution(sol)
        S.set_to_zero(3)
        S.assemble(3)
        S.constrain()
        out = {"dim": args.dim, "n": n, "n_inc": int(pb.n_inc), "nnz": int(S.nnz)}
        for rep in range(2):
            S.sync()
            t0 = time.perf_counter()
            info = S.solve(1e-8, 1e-14, 1e6, args.maxit, args.restart, args.pc, raise_on_fail=False)
            S.sync()
            out[f"solve{rep}"] = {"wall_ms": (time.perf_counter() - t0) * 1e3, "solve_ms": S.last_solve_ms(), "its": info.iterations,
                                  "converged": bool(info.converged), "rel": info.rel_residual, "axb": info.norm_axb,
                                  "rhs": info.norm_rhs}
        print(json.dumps(out), flush=True)
        del ls, S


if __name__ == "__main__":
    main()
