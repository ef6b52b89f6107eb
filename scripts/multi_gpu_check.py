#!/usr/bin/env python
"""Multi-GPU parity check, launched by torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        scripts/multi_gpu_check.py [--size 16]

Every rank assembles its strip on its GPU and the ranks solve the Newton linear system together (halo exchange over
NCCL before every SpMV, all-reduced dot products).  Rank 0 also solves the undecomposed problem on its GPU alone; the
distributed Newton solution must agree with it within the solver tolerance on every rank's owned rows, and the halo
exchange must reproduce the owners' values bit-for-bit."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=6)
    ap.add_argument("--dim", type=int, default=2, help="2: strips of triangles cut along y; 3: slabs of tetrahedra cut along z")
    ap.add_argument("--partition", default="structured", choices=["structured", "rcb"],
                    help="rcb: recursive coordinate bisection of an unstructured mesh (jittered vertices, shuffled element "
                         "order) held by every rank, rows matched through topological keys")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from feng_b200 import mesh as M, partition as PT, problems as PB
    from feng_b200.linear_system import LinearSystemB200, NLSolverOptions, solve_newton_raphson
    n = args.n
    mu, rho = 1.0, 1.0
    mg = None
    if args.partition == "rcb":
        # the same unstructured mesh on every rank: interior vertices jittered, elements in random order
        mg = M.rect_mesh(n, n * world, 1.0, float(world)) if args.dim == 2 else M.box_mesh(n, n, n * world, float(world))
        r0 = np.random.default_rng(11)
        interior = np.setdiff1d(np.arange(mg.n_vertices), np.unique(mg.bfacets))
        mg.xyz[interior, :args.dim] += r0.uniform(-0.15 / n, 0.15 / n, (interior.size, args.dim))
        mg.cells = np.ascontiguousarray(mg.cells[r0.permutation(mg.n_cells)])
        mg.bfacets = M.boundary_facets(mg.cells)
        mg.point_pressure = 0
        deg, fld = (8, 1) if args.dim == 2 else (6, 3)
        pb, part, gvert = PT.submesh_problem(mg, PT.rcb_partition(mg, world), rank, world, "ns_div", deg, fld, mu, rho)
    elif args.dim == 2:
        pb, part = PT.strip_problem(n, rank, world, "ns_div", 8, 1, mu, rho, with_source=False)
    else:
        pb, part = PT.slab_problem(n, rank, world, "ns_div", 6, 3, mu, rho, with_source=False)
    ls = LinearSystemB200(pb, device=local, device_pattern=True, partition=part)
    # halo exchange: ghost entries must come back as the owners' values
    rng = np.random.default_rng(5)
    if args.partition == "rcb":
        pg = PB.taylor_hood(mg, "ns_div", deg, fld, mu, rho, with_source=False)
        gkeys = PT.topological_keys(pg, np.arange(mg.n_vertices), mg.n_vertices)[:pg.n_inc]
    elif args.dim == 2:
        mg = M.rect_mesh(n, n * world, 1.0, float(world))
        mg.point_pressure = 0
        pg = PB.taylor_hood(mg, "ns_div", 8, 1, mu, rho, with_source=False)
        gkeys, _ = PT.dof_keys_and_owner(pg, n, 1)
    else:
        mg = M.box_mesh(n, n, n * world, float(world))
        mg.point_pressure = 0
        pg = PB.taylor_hood(mg, "ns_div", 6, 3, mu, rho, with_source=False)
        gkeys, _ = PT.dof_keys_and_owner(pg, n, 1)
    order = np.argsort(gkeys)
    pos = order[np.searchsorted(gkeys[order], part.keys)]
    xg = rng.standard_normal(pg.n_inc)
    x = xg[pos].copy()
    x[part.owned == 0] = -777.0
    x = ls.sys.halo_exchange_host(x)
    assert np.array_equal(x, xg[pos]), f"rank {rank}: halo exchange mismatch"
    # distributed Newton solve (Kovasznay boundary data, zero initial guess inside)
    ls.setRelativeTol(1e-9)        # restart length, iteration cap and preconditioner: the defaults (GMRES(30), 1e4, AUTO)
    sol = pb.sol.copy()
    sol[:pb.n_inc] = 0.0
    status, hist = solve_newton_raphson(ls, sol, NLSolverOptions(1e-7, 1e-7, 1e4, 20, 3, 1e-1))
    assert status == 0, (rank, status, hist[-1:] if hist else None)
    # reference: the undecomposed problem on one GPU (every rank solves it redundantly on its own GPU; small mesh)
    lg = LinearSystemB200(pg, device=local, device_pattern=True)
    lg.setRelativeTol(1e-9)
    solg = pg.sol.copy()
    solg[:pg.n_inc] = 0.0
    st2, hist2 = solve_newton_raphson(lg, solg, NLSolverOptions(1e-7, 1e-7, 1e4, 20, 3, 1e-1))
    assert st2 == 0
    own = np.nonzero(part.owned)[0]
    err = np.abs(sol[own] - solg[pos[own]]).max()
    scale = np.abs(solg[:pg.n_inc]).max()
    # ghost rows must hold the owners' corrections too (the local state feeds the next assembly)
    errg = np.abs(sol[:pb.n_inc] - solg[pos]).max()
    t = torch.tensor([err, errg], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"multi_gpu_check dim={args.dim} partition={args.partition} world={world} n={n}: newton its {len(hist)} (single GPU {len(hist2)}), "
              f"krylov its {[h['linearIter'] for h in hist]} vs {[h['linearIter'] for h in hist2]}, "
              f"max |du_dist - du_single| owned {t[0].item():.3e} all {t[1].item():.3e} (scale {scale:.3e})")
    assert t[0].item() <= 1e-5 * scale and t[1].item() <= 1e-5 * scale, t
    if rank == 0:
        print("MULTI_GPU_CHECK_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
