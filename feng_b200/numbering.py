"""DOF numbering of Lagrange spaces on straight (P1-geometry) triangle / tetrahedron meshes.

Host-side set-up code (numpy).  It reproduces, bit for bit, the numbering the reference produces for the same
mesh and the same list of spaces, so that element->DOF tables, CSR patterns and solution vectors can be compared
entry by entry (tests/test_numbering.py does so against the compiled reference):

  * one numbering per field, fields taken in order of first appearance        src/feNumber.cpp:698-703
  * all UNKNOWN DOFs of all fields first, then all ESSENTIAL ones             src/feNumber.cpp:776-784
  * inside a field: vertices (mesh vertex order, components consecutive),
    then element DOFs, then edges (edge-tag order), then faces               src/feNumber.cpp:370-483
  * edge tags = order of first appearance while sweeping the cells:
      triangles: local edges (0,1),(1,2),(2,0)                                src/feMeshRead.cpp:1412-1456
      tetrahedra: the 4 faces {0,2,1},{0,1,3},{0,3,2},{3,1,2}, each giving its
      3 edges, i.e. first appearances (0,2),(2,1),(1,0),(1,3),(3,0),(3,2)     src/feTetrahedron.cpp:11-34,
                                                                              src/feTriangle.cpp:3-26, src/feTetrahedron.h:30-31
  * P2 boundary *line* spaces carry their mid-side DOF as an element DOF that aliases the edge DOF
    (src/feSpace_1D.cpp:490-533, src/feNumber.cpp:393-401): essential mid-side DOFs of a 2-D problem are therefore
    numbered in boundary-element order, before any remaining essential edge.
  * element->DOF tables ("adr"):  TriP1 src/feSpace_2D.cpp:154-159, TriP2 :772-788, VecTriP2 :1017-1041,
    TetPn src/feSpace_3D.cpp:293-316 (vertices, then the 6 edges in the order of src/feTetrahedron.h:31).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .mesh import Mesh, TET_EDGES, TRI_EDGES

UNASSIGNED, UNKNOWN, ESSENTIAL = -3, -1, -2


@dataclass
class SpaceSpec:
    field: str            # field name, e.g. "U", "P"
    where: str            # "domain", "boundary" or "point"
    order: int            # Lagrange order 0 (point), 1 or 2
    ncomp: int = 1        # 1 = scalar Lagrange, dim = vector Lagrange
    essential: bool = False


@dataclass
class FieldNumbering:
    ncomp: int
    vertex_dof: np.ndarray            # (nV, ncomp) int64, UNASSIGNED where the field has no DOF
    edge_dof: np.ndarray              # (nEdges, ncomp) int64


@dataclass
class Numbering:
    n_inc: int
    n_dof: int
    fields: dict = field(default_factory=dict)     # name -> FieldNumbering
    edges: np.ndarray | None = None                # (nEdges, 2) vertices, first-appearance orientation
    cell_edges: np.ndarray | None = None           # (nE, 3|6) edge index of each local edge

    def adr(self, mesh: Mesh, fld: str, order: int) -> np.ndarray:
        """(nE, nF) element->global DOF table of the domain space of field `fld`."""
        fn = self.fields[fld]
        nc = fn.ncomp
        v = fn.vertex_dof[mesh.cells]                          # (nE, nv, nc)
        parts = [v.reshape(mesh.n_cells, -1)]
        if order == 2:
            e = fn.edge_dof[self.cell_edges]                   # (nE, ne, nc)
            parts.append(e.reshape(mesh.n_cells, -1))
        return np.ascontiguousarray(np.concatenate(parts, 1))


def _edge_device(n_pairs: int):
    """Device index for the edge tags, or None: the GPU path is taken for meshes where numpy.unique is the slowest step of the
    set-up (>= 2 M vertex pairs) when a CUDA device is visible; B200_EDGES=host | device overrides."""
    import os
    mode = os.environ.get("B200_EDGES", "auto")
    if mode == "host" or (mode == "auto" and n_pairs < 2_000_000):
        return None
    try:
        import torch
        if not torch.cuda.is_available():
            return None
        return int(os.environ.get("LOCAL_RANK", "0")) % max(torch.cuda.device_count(), 1)
    except Exception:
        return None


def build_edges(mesh: Mesh, device="auto"):
    """Unique edges in order of first appearance and the (nE, n_local_edges) cell->edge table."""
    loc = TRI_EDGES if mesh.dim == 2 else TET_EDGES
    ev = mesh.cells[:, loc].astype(np.int64)                   # (nE, ne, 2)
    flat = ev.reshape(-1, 2)
    n_pre = 0
    if mesh.dim == 3 and mesh.bfacets.size:
        # In a .msh file the boundary triangles (dim-2 block) precede the tetrahedra, and the reader creates the
        # edges of every 3-node triangle it meets (src/feMeshRead.cpp:1317-1338) in the same edge set the
        # tetrahedra use (:1498, :1613-1614): boundary edges get the first tags, in boundary-triangle order.
        pre = mesh.bfacets[:, TRI_EDGES].astype(np.int64).reshape(-1, 2)
        n_pre = pre.shape[0]
        flat = np.concatenate([pre, flat], 0)
    if device == "auto":
        device = _edge_device(flat.shape[0])
    if device is not None and flat.shape[0] < 2 ** 31 - 1:
        # same tags from the device (csrc/numbering.cu: sort / segment / rank on the GPU); bit-identical, tests/test_gpu_numbering.py
        from . import capi
        eop, edges = capi.unique_edges(mesh.n_vertices, flat, device)
        cell_edges = eop[n_pre:].reshape(mesh.n_cells, loc.shape[0]).astype(np.int64)
        return edges.astype(np.int64), cell_edges
    lo = flat.min(1)
    hi = flat.max(1)
    key = lo * np.int64(mesh.n_vertices) + hi
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")                   # unique ids sorted by first appearance
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    cell_edges = rank[inv.reshape(-1)[n_pre:]].reshape(mesh.n_cells, loc.shape[0])
    edges = flat[first[order]]
    return edges, cell_edges


def _edge_lookup(edges: np.ndarray, n_vertices: int, pairs: np.ndarray) -> np.ndarray:
    key = np.minimum(edges[:, 0], edges[:, 1]) * np.int64(n_vertices) + np.maximum(edges[:, 0], edges[:, 1])
    srt = np.argsort(key)
    q = np.minimum(pairs[:, 0], pairs[:, 1]).astype(np.int64) * np.int64(n_vertices) + \
        np.maximum(pairs[:, 0], pairs[:, 1]).astype(np.int64)
    pos = np.searchsorted(key[srt], q)
    return srt[pos]


def build_numbering(mesh: Mesh, spaces: list[SpaceSpec]) -> Numbering:
    nV = mesh.n_vertices
    edges, cell_edges = build_edges(mesh)
    nEd = edges.shape[0]
    bverts = np.unique(mesh.bfacets)
    if mesh.dim == 2:
        bedges = _edge_lookup(edges, nV, mesh.bfacets)                    # one per boundary line, in line order
    else:
        tri_e = mesh.bfacets[:, [[0, 1], [1, 2], [2, 0]]].reshape(-1, 2)
        bedges = _edge_lookup(edges, nV, tri_e)
    # sub-meshes of a partition (feng_b200/partition.py): vertices / edges of the ghost layer that lie on the physical
    # boundary of the WHOLE mesh although no boundary facet of this sub-mesh holds them
    xv = getattr(mesh, "extra_boundary_vertices", None)
    if xv is not None and len(xv):
        bverts = np.union1d(bverts, np.asarray(xv, np.int64))
    xe = getattr(mesh, "extra_boundary_edges", None)
    if xe is not None and len(xe):
        extra = _edge_lookup(edges, nV, np.asarray(xe, np.int64).reshape(-1, 2))
        bedges = np.concatenate([bedges, extra[~np.isin(extra, bedges)]])

    names = []
    for s in spaces:
        if s.field not in names:
            names.append(s.field)

    codes = {}
    for name in names:
        nc = max(s.ncomp for s in spaces if s.field == name)
        codes[name] = dict(nc=nc, v=np.full(nV, UNASSIGNED, np.int8), e=np.full(nEd, UNASSIGNED, np.int8),
                           belem=False)
    # unknown marks of every space first, then essential marks (src/feNumber.cpp:747-753)
    for s in spaces:
        c = codes[s.field]
        if s.where == "domain":
            c["v"][np.unique(mesh.cells)] = UNKNOWN
            if s.order == 2:
                c["e"][:] = UNKNOWN
        elif s.where == "boundary":
            c["v"][bverts] = UNKNOWN
            if s.order == 2:
                c["e"][bedges] = UNKNOWN
        elif s.where == "point":
            c["v"][mesh.point_pressure] = UNKNOWN
    for s in spaces:
        if not s.essential:
            continue
        c = codes[s.field]
        if s.where == "boundary":
            c["v"][bverts] = ESSENTIAL
            if s.order == 2:
                c["e"][bedges] = ESSENTIAL
                c["belem"] = mesh.dim == 2      # mid-side DOFs of boundary lines are element DOFs
        elif s.where == "point":
            c["v"][mesh.point_pressure] = ESSENTIAL
        elif s.where == "domain":
            c["v"][np.unique(mesh.cells)] = ESSENTIAL
            if s.order == 2:
                c["e"][:] = ESSENTIAL

    num = Numbering(0, 0, {}, edges, cell_edges)
    for name in names:
        nc = codes[name]["nc"]
        num.fields[name] = FieldNumbering(nc, np.full((nV, nc), UNASSIGNED, np.int64),
                                          np.full((nEd, nc), UNASSIGNED, np.int64))
    comp = lambda nc: np.arange(nc, dtype=np.int64)[None, :]
    g = 0
    for code in (UNKNOWN, ESSENTIAL):
        for name in names:
            c, fn = codes[name], num.fields[name]
            nc = fn.ncomp
            idx = np.nonzero(c["v"] == code)[0]
            fn.vertex_dof[idx] = g + nc * np.arange(idx.size, dtype=np.int64)[:, None] + comp(nc)
            g += nc * idx.size
            done = np.zeros(nEd, bool)
            if code == ESSENTIAL and c["belem"]:
                # element DOFs of the boundary lines, in boundary-element order, alias the edge DOFs
                fn.edge_dof[bedges] = g + nc * np.arange(bedges.size, dtype=np.int64)[:, None] + comp(nc)
                g += nc * bedges.size
                done[bedges] = True
            idx = np.nonzero((c["e"] == code) & ~done)[0]
            fn.edge_dof[idx] = g + nc * np.arange(idx.size, dtype=np.int64)[:, None] + comp(nc)
            g += nc * idx.size
        if code == UNKNOWN:
            num.n_inc = g
    num.n_dof = g
    return num


def taylor_hood_spaces(dim: int, p_essential_boundary: bool = False, point_pressure: bool = True):
    """The space list of the reference's (Navier-)Stokes drivers (tests/withLinearSolver/navier_stokes.cpp:74-80)."""
    sp = [SpaceSpec("U", "domain", 2, dim), SpaceSpec("U", "boundary", 2, dim, True), SpaceSpec("P", "domain", 1, 1)]
    if p_essential_boundary:
        sp.append(SpaceSpec("P", "boundary", 1, 1, True))
    elif point_pressure:
        sp.append(SpaceSpec("P", "point", 0, 1, True))
    return sp


def scalar_spaces(order: int = 2):
    """tests/withLinearSolver/convergenceLaplace.cpp:64-69"""
    return [SpaceSpec("U", "domain", order, 1), SpaceSpec("U", "boundary", order, 1, True)]
