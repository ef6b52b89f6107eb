"""Host logic of the multi-GPU path on CPU: two processes over gloo (world_size 2) build their strips, the halo plan
and check, against the undecomposed problem assembled by the oracle, that
  * every rank's OWNED rows of the locally assembled matrix / rhs are the global rows (owner-computes with one ghost
    layer needs no exchange of matrix entries),
  * the halo exchange + owned-row SpMV reproduces the global product, and owned-row dot products sum to the global one.
The same Partition object drives the NCCL path on GPUs (feng_b200/csrc/comm.cu, tests/test_gpu_multi.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, out, dim=2, strong=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scipy.sparse as sp
    import torch
    import torch.distributed as dist
    from conftest import to_oracle_problem
    from feng_b200 import mesh as M, partition as PT, problems as PB
    from oracle import fe_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if dim == 2:
            pb, part = PT.strip_problem(n, rank, world, "ns_div", 8, 1, 0.05, 1.0, build_pattern=True, with_source=True)
            # undecomposed problem, identical on every rank
            mg = M.rect_mesh(n, n * world, 1.0, float(world))
            mg.point_pressure = 0
            pg = PB.taylor_hood(mg, "ns_div", 8, 1, 0.05, 1.0, with_source=True)
        elif strong:
            # strong scaling: ONE T3D(n) cube cut into `world` slabs of (nearly) equal thickness
            pb, part = PT.slab_problem(n, rank, world, "ns_div", 6, 3, 0.05, 1.0, build_pattern=True, with_source=False, strong=True)
            mg = M.box_mesh(n, n, n, 1.0)
            mg.point_pressure = 0
            pg = PB.taylor_hood(mg, "ns_div", 6, 3, 0.05, 1.0, with_source=False)
        else:
            pb, part = PT.slab_problem(n, rank, world, "ns_div", 6, 3, 0.05, 1.0, build_pattern=True, with_source=False)
            mg = M.box_mesh(n, n, n * world, float(world))
            mg.point_pressure = 0
            pg = PB.taylor_hood(mg, "ns_div", 6, 3, 0.05, 1.0, with_source=False)

        class G:     # global keys of the undecomposed numbering (same key function, one "strip" of height world)
            pass
        gkeys, _ = PT.dof_keys_and_owner(pg, n, 1)
        order = np.argsort(gkeys)
        pos = order[np.searchsorted(gkeys[order], part.keys)]          # local row -> global row
        assert np.array_equal(gkeys[pos], part.keys)
        # every global row is owned by exactly one rank
        cnt = torch.zeros(pg.n_inc, dtype=torch.int64)
        cnt[torch.from_numpy(pos[part.owned == 1])] += 1
        dist.all_reduce(cnt)
        assert bool((cnt == 1).all())
        # state: a global random vector restricted to the strip
        rng = np.random.default_rng(7)
        solg = pg.sol.copy()
        solg[:pg.n_inc] += rng.uniform(-1e-2, 1e-2, pg.n_inc)
        sol = pb.sol.copy()
        sol[:pb.n_inc] = solg[pos]
        vg, rg = O.assemble(to_oracle_problem(pg), pg.ia, pg.ja, solg)
        v, r = O.assemble(to_oracle_problem(pb), pb.ia, pb.ja, sol)
        Ag = sp.csr_matrix((vg, pg.ja, pg.ia), shape=(pg.n_inc, pg.n_inc))
        A = sp.csr_matrix((v, pb.ja, pb.ia), shape=(pb.n_inc, pb.n_inc))
        own = np.nonzero(part.owned)[0]
        assert np.abs(r[own] - rg[pos[own]]).max() <= 1e-13 * np.abs(rg).max()
        # owned rows, column by column through the key map
        P = sp.csr_matrix((np.ones(pb.n_inc), (np.arange(pb.n_inc), pos)), shape=(pb.n_inc, pg.n_inc))
        D = (A[own] - (P @ Ag @ P.T)[own])
        assert abs(D).max() <= 1e-13 * abs(Ag).max()
        assert (P @ Ag)[own].nnz == (P @ Ag @ P.T)[own].nnz            # no owned row couples outside the strip
        # halo exchange of a vector whose ghost entries are wrong, then the owned-row product
        xg = rng.standard_normal(pg.n_inc)
        x = xg[pos].copy()
        x[part.owned == 0] = np.nan
        reqs, bufs = [], []
        for k, nb in enumerate(part.neighbors):
            s = torch.from_numpy(x[part.send_idx[part.send_ptr[k]:part.send_ptr[k + 1]]].copy())
            rcv = torch.zeros(int(part.recv_ptr[k + 1] - part.recv_ptr[k]), dtype=torch.float64)
            reqs += [dist.isend(s, int(nb)), dist.irecv(rcv, int(nb))]
            bufs.append((k, rcv))
        for q in reqs:
            q.wait()
        for k, rcv in bufs:
            x[part.recv_idx[part.recv_ptr[k]:part.recv_ptr[k + 1]]] = rcv.numpy()
        assert np.array_equal(x, xg[pos])
        y = A @ x
        assert np.abs(y[own] - (Ag @ xg)[pos[own]]).max() <= 1e-12 * np.abs(Ag @ xg).max()
        d = torch.tensor([float(np.dot(x[own], y[own]))], dtype=torch.float64)
        dist.all_reduce(d)
        assert abs(d.item() - float(xg @ (Ag @ xg))) <= 1e-10 * abs(float(xg @ (Ag @ xg)))
        out[rank] = "ok"
    except Exception as ex:     # noqa: BLE001
        import traceback
        out[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_strip_partition_over_gloo(world):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() + world) % 2000
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, 4, out), nprocs=world, join=True)
    for r in range(world):
        assert out.get(r) == "ok", out.get(r)


def test_slab_partition_over_gloo():
    """3-D counterpart: slabs of the tetrahedral box cut along z (the partition bench.py uses for T3D at N > 1)."""
    import torch.multiprocessing as mp
    world = 2
    port = 31500 + os.getpid() % 2000
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, 2, out, 3), nprocs=world, join=True)
    for r in range(world):
        assert out.get(r) == "ok", out.get(r)


@pytest.mark.parametrize("world", [2, 3])
def test_strong_scaling_slabs_over_gloo(world):
    """strong scaling (bench.py --scaling strong): one T3D(5) cube dealt out in slabs of 2+3 cell layers, one T3D(6) cube in
    2+2+2 (a part needs two layers: its ghost layer must not reach the physical boundary of the neighbour's side)"""
    import torch.multiprocessing as mp
    port = 33500 + (os.getpid() + world) % 2000
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, 5 if world == 2 else 6, out, 3, True), nprocs=world, join=True)
    for r in range(world):
        assert out.get(r) == "ok", out.get(r)


def _worker_rcb(rank, world, port, dim, out):
    """General meshes: recursive coordinate bisection + topological row keys (feng_b200.partition.submesh_problem) on the
    reference's unstructured data/square1.msh (from the committed fixture) or on a small tetrahedral cube."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scipy.sparse as sp
    import torch
    import torch.distributed as dist
    from conftest import to_oracle_problem
    from feng_b200 import mesh as M, partition as PT, problems as PB
    from oracle import fe_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if dim == 2:
            g = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_square1_ns_div.npz"), allow_pickle=False))
            cells = g["cells"].astype(np.int32)
            mg = M.Mesh(2, g["xyz"], cells, M.boundary_facets(cells), 0)
            deg, fld = 8, 1
        else:
            mg = M.cube_mesh(3)
            deg, fld = 6, 3
        part = PT.rcb_partition(mg, world)
        pb, prt, gvert = PT.submesh_problem(mg, part, rank, world, "ns_div", deg, fld, 0.05, 1.0, build_pattern=True)
        pg = PB.taylor_hood(mg, "ns_div", deg, fld, 0.05, 1.0, with_source=False)
        gkeys = PT.topological_keys(pg, np.arange(mg.n_vertices), mg.n_vertices)[:pg.n_inc]
        order = np.argsort(gkeys)
        pos = order[np.searchsorted(gkeys[order], prt.keys)]
        assert np.array_equal(gkeys[pos], prt.keys)                     # every local unknown is a global unknown
        cnt = torch.zeros(pg.n_inc, dtype=torch.int64)
        cnt[torch.from_numpy(pos[prt.owned == 1])] += 1
        dist.all_reduce(cnt)
        assert bool((cnt == 1).all())                                   # exactly one owner per row
        rng = np.random.default_rng(7)
        solg = pg.sol.copy()
        solg[:pg.n_inc] += rng.uniform(-1e-2, 1e-2, pg.n_inc)
        sol = pb.sol.copy()
        sol[:pb.n_inc] = solg[pos]
        vg, rg = O.assemble(to_oracle_problem(pg), pg.ia, pg.ja, solg)
        v, r = O.assemble(to_oracle_problem(pb), pb.ia, pb.ja, sol)
        Ag = sp.csr_matrix((vg, pg.ja, pg.ia), shape=(pg.n_inc, pg.n_inc))
        A = sp.csr_matrix((v, pb.ja, pb.ia), shape=(pb.n_inc, pb.n_inc))
        own = np.nonzero(prt.owned)[0]
        assert np.abs(r[own] - rg[pos[own]]).max() <= 1e-13 * np.abs(rg).max()
        P = sp.csr_matrix((np.ones(pb.n_inc), (np.arange(pb.n_inc), pos)), shape=(pb.n_inc, pg.n_inc))
        assert abs(A[own] - (P @ Ag @ P.T)[own]).max() <= 1e-13 * abs(Ag).max()
        assert (P @ Ag)[own].nnz == (P @ Ag @ P.T)[own].nnz            # no owned row couples outside the sub-mesh
        # halo exchange, then the owned-row product
        xg = rng.standard_normal(pg.n_inc)
        x = xg[pos].copy()
        x[prt.owned == 0] = np.nan
        reqs, bufs = [], []
        for k, nb in enumerate(prt.neighbors):
            s = torch.from_numpy(x[prt.send_idx[prt.send_ptr[k]:prt.send_ptr[k + 1]]].copy())
            rcv = torch.zeros(int(prt.recv_ptr[k + 1] - prt.recv_ptr[k]), dtype=torch.float64)
            reqs += [dist.isend(s, int(nb)), dist.irecv(rcv, int(nb))]
            bufs.append((k, rcv))
        for q in reqs:
            q.wait()
        for k, rcv in bufs:
            x[prt.recv_idx[prt.recv_ptr[k]:prt.recv_ptr[k + 1]]] = rcv.numpy()
        assert np.array_equal(x, xg[pos])
        y = A @ x
        assert np.abs(y[own] - (Ag @ xg)[pos[own]]).max() <= 1e-12 * np.abs(Ag @ xg).max()
        out[rank] = "ok"
    except Exception:     # noqa: BLE001
        import traceback
        out[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dim", [(2, 2), (3, 2), (2, 3)])
def test_rcb_partition_of_general_meshes_over_gloo(world, dim):
    import torch.multiprocessing as mp
    port = 33500 + (os.getpid() + 7 * world + dim) % 2000
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_rcb, args=(world, port, dim, out), nprocs=world, join=True)
    for r in range(world):
        assert out.get(r) == "ok", out.get(r)
