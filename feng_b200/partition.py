"""Element / row partition for the multi-GPU path: strips (2-D) and slabs (3-D) of the structured
generators.

The reference offers no domain decomposition (every MPI rank holds the whole mesh and keeps its row range,
src/feLinearSystemPETSc.cpp:626-640); METIS is a CMake option no source file uses.  Here every rank builds ONLY its own
strip of the mesh plus one ghost layer of elements towards each neighbour:

  * assembly is owner-computes: every row a rank owns sees all of its elements locally, no exchange of matrix entries;
  * each row (unknown DOF) is owned by exactly one rank; rows on a cut line belong to the lower rank;
  * ghost rows are incomplete locally and never used: dot products run over owned rows, the SpMV input gets its ghost
    entries from the owners (halo exchange) before every product.

Rows are matched across ranks through a geometric key (field, component, position on the fine lattice), so the
halo plan needs one all-gather of the ghost keys at set-up and nothing else.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import mesh as M
from . import problems as PB


@dataclass
class Partition:
    rank: int
    world: int
    owned: np.ndarray            # (n_inc,) uint8, 1 if this rank owns the row
    keys: np.ndarray             # (n_inc,) int64 global key of every unknown
    neighbors: np.ndarray        # (n_nbr,) int32
    send_ptr: np.ndarray         # (n_nbr + 1,) int64
    send_idx: np.ndarray         # local rows sent to the neighbours (this rank owns them)
    recv_ptr: np.ndarray
    recv_idx: np.ndarray         # local ghost rows received from the neighbours
    owned_cells: int = 0

    @property
    def n_owned(self):
        return int(self.owned.sum())


def strip_mesh(n: int, rank: int, world: int) -> tuple[M.Mesh, int]:
    """Strip `rank` of the [0,1] x [0,world] domain: n x n owned cells plus one ghost row of cells towards each
    neighbour; the artificial cuts carry no physical boundary."""
    gb, gt = (1 if rank > 0 else 0), (1 if rank < world - 1 else 0)
    ny = n + gb + gt
    m = M.rect_mesh(n, ny, 1.0, ny / n, 0.0, rank - gb / n, cut_bottom=rank > 0, cut_top=rank < world - 1)
    if rank == 0:
        m.point_pressure = 0
    return m, 2 * n * n


def layer_bounds(n: int, world: int, strong: bool) -> np.ndarray:
    """cell-layer boundaries of the parts along the cut axis: weak scaling = n layers per part of a box of height `world`;
    strong scaling = the n layers of the unit cube dealt out as evenly as possible"""
    if not strong:
        return np.arange(world + 1, dtype=np.int64) * n
    return (np.arange(world + 1, dtype=np.int64) * n) // world


def slab_mesh(n: int, rank: int, world: int, strong: bool = False) -> tuple[M.Mesh, int]:
    """Slab `rank` of the [0,1]^2 x [0,world] box (weak scaling: n^3 owned cells = 6 n^3 tetrahedra per rank) or of the unit
    cube T3D(n) cut into `world` slabs (strong scaling), plus one ghost layer of cells towards each neighbour; the artificial
    cuts carry no physical boundary."""
    b = layer_bounds(n, world, strong)
    k0, k1 = int(b[rank]), int(b[rank + 1])
    assert world == 1 or np.diff(b).min() >= 2, "every slab needs at least two cell layers"
    gb, gt = (1 if rank > 0 else 0), (1 if rank < world - 1 else 0)
    nz = (k1 - k0) + gb + gt
    m = M.box_mesh(n, n, nz, nz / n, (k0 - gb) / n, cut_bottom=rank > 0, cut_top=rank < world - 1)
    if rank == 0:
        m.point_pressure = 0
    return m, 6 * n * n * (k1 - k0)


def dof_keys_and_owner(pb: PB.HostProblem, n: int, world: int, bounds: np.ndarray | None = None):
    """Global key and owning rank of every unknown of a strip (2-D, cut along y) or slab (3-D, cut along z) problem:
    unit cells of size 1/n; parts of height 1 (n layers each) unless `bounds` gives the cell-layer boundaries."""
    num, mesh, n_dof = pb.num, pb.mesh, pb.n_dof
    axis = pb.dim - 1
    keys = np.full(n_dof, -1, np.int64)
    own = np.zeros(n_dof, np.int64)
    for f, fld in enumerate(num.fields):
        xyz = PB.dof_coordinates(mesh, num, fld, n_dof)
        comp = PB.dof_components(num, fld, n_dof)
        sel = comp >= 0
        ix = np.rint(xyz[sel, 0] * 2 * n).astype(np.int64)
        iy = np.rint(xyz[sel, 1] * 2 * n).astype(np.int64)
        if pb.dim == 2:
            keys[sel] = ((f * 4 + comp[sel]) << 56) | (iy << 28) | ix
            ic = iy
        else:
            iz = np.rint(xyz[sel, 2] * 2 * n).astype(np.int64)
            keys[sel] = ((f * 4 + comp[sel]) << 58) | (iz << 38) | (iy << 19) | ix
            ic = iz
        # part r owns the coordinate range (r, r + 1] along the cut axis; the bottom side belongs to rank 0
        if bounds is None:
            r = np.ceil(ic / (2.0 * n)).astype(np.int64) - 1
        else:
            r = np.searchsorted(2 * np.asarray(bounds, np.int64), ic, side="left") - 1     # (2 b_r, 2 b_{r+1}]
        own[sel] = np.clip(r, 0, world - 1)
    return keys[:pb.n_inc], own[:pb.n_inc]


def build_halo(keys: np.ndarray, owner: np.ndarray, rank: int, world: int, allgather, owned_cells: int = 0) -> Partition:
    """allgather(obj) -> list of every rank's obj (torch.distributed.all_gather_object or a test double)."""
    owned = (owner == rank)
    ghost = np.nonzero(~owned)[0]
    need = {}
    for r in np.unique(owner[ghost]):
        idx = ghost[owner[ghost] == r]
        order = np.argsort(keys[idx], kind="stable")
        need[int(r)] = (keys[idx][order], idx[order])
    everyone = allgather({r: k for r, (k, _) in need.items()})
    own_idx = np.nonzero(owned)[0]
    order = np.argsort(keys[own_idx], kind="stable")
    own_keys, own_loc = keys[own_idx][order], own_idx[order]
    nbrs = sorted(set(need) | {r for r, req in enumerate(everyone) if rank in req and r != rank})
    send_ptr, recv_ptr, send_idx, recv_idx = [0], [0], [], []
    for r in nbrs:
        req = everyone[r].get(rank)
        if req is not None and len(req):
            pos = np.searchsorted(own_keys, req)
            if (pos >= own_keys.size).any() or (own_keys[np.minimum(pos, own_keys.size - 1)] != req).any():
                raise RuntimeError(f"rank {rank}: neighbour {r} asks for rows this rank does not own")
            send_idx.append(own_loc[pos])
        send_ptr.append(send_ptr[-1] + (0 if req is None else len(req)))
        if r in need:
            recv_idx.append(need[r][1])
        recv_ptr.append(recv_ptr[-1] + (len(need[r][1]) if r in need else 0))
    cat = lambda parts: np.concatenate(parts).astype(np.int32) if parts else np.zeros(0, np.int32)
    return Partition(rank, world, owned.astype(np.uint8), keys, np.array(nbrs, np.int32), np.array(send_ptr, np.int64),
                     cat(send_idx), np.array(recv_ptr, np.int64), cat(recv_idx), owned_cells)


def strip_problem(n: int, rank: int, world: int, kind: str = "ns_div", quad_degree: int = 8, field_id: int = 1,
                  mu: float = 1.0 / 40.0, rho: float = 1.0, allgather=None, build_pattern: bool = False,
                  with_source: bool = False):
    """(HostProblem, Partition) of strip `rank`."""
    m, owned_cells = strip_mesh(n, rank, world)
    pb = PB.taylor_hood(m, kind, quad_degree, field_id, mu, rho, build_pattern=build_pattern, with_source=with_source)
    if world == 1:
        return pb, None
    keys, owner = dof_keys_and_owner(pb, n, world)
    if allgather is None:
        import torch.distributed as dist

        def allgather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
    return pb, build_halo(keys, owner, rank, world, allgather, owned_cells)


def slab_problem(n: int, rank: int, world: int, kind: str = "ns_div", quad_degree: int = 6, field_id: int = 3,
                 mu: float = 1.0 / 40.0, rho: float = 1.0, allgather=None, build_pattern: bool = False,
                 with_source: bool = False, strong: bool = False):
    """(HostProblem, Partition) of slab `rank` of the tetrahedral box (the 3-D counterpart of strip_problem); strong = the
    unit cube T3D(n) cut into `world` slabs instead of one T3D(n) cube per rank."""
    m, owned_cells = slab_mesh(n, rank, world, strong)
    pb = PB.taylor_hood(m, kind, quad_degree, field_id, mu, rho, build_pattern=build_pattern, with_source=with_source)
    if world == 1:
        return pb, None
    keys, owner = dof_keys_and_owner(pb, n, world, layer_bounds(n, world, strong) if strong else None)
    if allgather is None:
        import torch.distributed as dist

        def allgather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
    return pb, build_halo(keys, owner, rank, world, allgather, owned_cells)


# ---------------------------------------------------------------------------------------------------------------
# General meshes: recursive coordinate bisection of the elements (METIS is not in this image; any element -> part map
# can be passed instead), topological row keys
# ---------------------------------------------------------------------------------------------------------------
def rcb_partition(mesh: M.Mesh, world: int) -> np.ndarray:
    """part[nE]: recursive coordinate bisection of the element centroids into `world` parts of (almost) equal size;
    every cut is perpendicular to the longest extent of the piece it splits."""
    cen = mesh.xyz[mesh.cells].mean(1)[:, :mesh.dim]
    part = np.zeros(mesh.n_cells, np.int32)

    def split(idx, p0, n):
        if n == 1:
            part[idx] = p0
            return
        c = cen[idx]
        ax = int(np.argmax(c.max(0) - c.min(0)))
        nl = n // 2
        k = (idx.size * nl) // n
        order = np.argsort(c[:, ax], kind="stable")
        split(idx[order[:k]], p0, nl)
        split(idx[order[k:]], p0 + nl, n - nl)
    split(np.arange(mesh.n_cells), 0, world)
    return part


def _sorted_key(v: np.ndarray, n: int) -> np.ndarray:
    s = np.sort(v.astype(np.int64), axis=1)
    key = s[:, 0]
    for c in range(1, s.shape[1]):
        key = key * np.int64(n) + s[:, c]
    return key


def topological_keys(pb: PB.HostProblem, gvert: np.ndarray, n_vertices_global: int) -> np.ndarray:
    """Global identity of every DOF of a (sub-)problem: (field, component, vertex | edge of the WHOLE mesh).  gvert maps
    the vertices of pb.mesh to vertices of the whole mesh."""
    num = pb.num
    keys = np.full(pb.n_dof, -1, np.int64)
    ge = gvert[num.edges]
    ekey = np.minimum(ge[:, 0], ge[:, 1]).astype(np.int64) * np.int64(n_vertices_global) + np.maximum(ge[:, 0], ge[:, 1])
    for f, fn in enumerate(num.fields.values()):
        for c in range(fn.ncomp):
            tag = np.int64((f * 4 + c)) << np.int64(58)
            vd = fn.vertex_dof[:, c]
            ok = vd >= 0
            keys[vd[ok]] = tag | gvert[ok].astype(np.int64)
            ed = fn.edge_dof[:, c]
            ok = ed >= 0
            keys[ed[ok]] = tag | (np.int64(1) << np.int64(57)) | ekey[ok]
    return keys


def submesh_problem(mesh: M.Mesh, part: np.ndarray, rank: int, world: int, kind: str = "ns_div", quad_degree: int = 8,
                    field_id: int = 1, mu: float = 1.0 / 40.0, rho: float = 1.0, allgather=None,
                    build_pattern: bool = False, with_source: bool = False):
    """(HostProblem, Partition, gvert) of part `rank` of an arbitrary simplicial mesh held by every rank (as the reference
    does, src/feLinearSystemPETSc.cpp:626-640).  A node (vertex or edge) belongs to the lowest part among its adjacent
    elements; the sub-mesh holds every element adjacent to an owned node (owner-computes, one ghost layer), its physical
    boundary facets, and the ghost-layer vertices / edges of the physical boundary as essential entities."""
    from . import numbering as NB
    nVg = mesh.n_vertices
    edges_g, cell_edges_g = NB.build_edges(mesh)
    vo = np.full(nVg, world, np.int64)
    np.minimum.at(vo, mesh.cells.reshape(-1), np.repeat(part.astype(np.int64), mesh.cells.shape[1]))
    eo = np.full(edges_g.shape[0], world, np.int64)
    np.minimum.at(eo, cell_edges_g.reshape(-1), np.repeat(part.astype(np.int64), cell_edges_g.shape[1]))
    local = np.nonzero((vo[mesh.cells] == rank).any(1) | (eo[cell_edges_g] == rank).any(1))[0]
    gvert = np.unique(mesh.cells[local])
    g2l = np.full(nVg, -1, np.int64)
    g2l[gvert] = np.arange(gvert.size)
    cells = g2l[mesh.cells[local]].astype(np.int32)
    # physical boundary facets of the sub-mesh = its one-sided facets that are boundary facets of the whole mesh
    bkey_g = np.sort(_sorted_key(mesh.bfacets, nVg))
    bf = M.boundary_facets(cells)
    bf = np.ascontiguousarray(bf[np.isin(_sorted_key(gvert[bf], nVg), bkey_g)]).astype(np.int32)
    pp = None
    if mesh.point_pressure is not None and g2l[mesh.point_pressure] >= 0:
        pp = int(g2l[mesh.point_pressure])
    sub = M.Mesh(mesh.dim, np.ascontiguousarray(mesh.xyz[gvert]), cells, bf, pp)
    # ghost-layer entities on the physical boundary of the whole mesh
    bv_g = np.unique(mesh.bfacets)
    sub.extra_boundary_vertices = np.nonzero(np.isin(gvert, bv_g))[0]
    be_g = mesh.bfacets if mesh.dim == 2 else mesh.bfacets[:, [[0, 1], [1, 2], [2, 0]]].reshape(-1, 2)
    bekey_g = np.unique(_sorted_key(be_g, nVg))
    edges_l, _ = NB.build_edges(sub)
    on_b = np.isin(_sorted_key(gvert[edges_l], nVg), bekey_g)
    sub.extra_boundary_edges = edges_l[on_b]
    pb = PB.taylor_hood(sub, kind, quad_degree, field_id, mu, rho, build_pattern=build_pattern, with_source=with_source)
    keys = topological_keys(pb, gvert, nVg)
    # owner of every DOF = owner of its node
    owner = np.zeros(pb.n_dof, np.int64)
    ekey_sorted_idx = np.argsort(_sorted_key(edges_g, nVg))
    ekey_sorted = _sorted_key(edges_g, nVg)[ekey_sorted_idx]
    ge = ekey_sorted_idx[np.searchsorted(ekey_sorted, _sorted_key(gvert[pb.num.edges], nVg))]   # local edge -> global edge
    for fn in pb.num.fields.values():
        for c in range(fn.ncomp):
            vd = fn.vertex_dof[:, c]
            ok = vd >= 0
            owner[vd[ok]] = vo[gvert[ok]]
            ed = fn.edge_dof[:, c]
            ok = ed >= 0
            owner[ed[ok]] = eo[ge[ok]]
    if world == 1:
        return pb, None, gvert
    if allgather is None:
        import torch.distributed as dist

        def allgather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
    owned_cells = int((part == rank).sum())
    return pb, build_halo(keys[:pb.n_inc], owner[:pb.n_inc], rank, world, allgather, owned_cells), gvert
