"""Synthetic refined meshes generated in code and a Gmsh 4.1 ASCII writer.

T2D(N): unit square, N x N cells, each split along the same diagonal into two CCW triangles.
T3D(N): unit cube, N^3 cells, each split into 6 Kuhn tetrahedra.

The writer exists so that the unmodified reference reader (src/feMeshRead.cpp:645-, contract summarised in
SURVEY.md section 7-1) ingests the IDENTICAL mesh: the vertex order of the file is the host vertex index
(src/feMeshRead.cpp:942-958), element order inside the single (dim, entity) block is the local element index,
and boundary facets are oriented like the adjacent cell's facet (src/feMeshRead.cpp:1836-1898).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# local facets of a cell, in the reference's order:
#   triangle edges (0,1),(1,2),(2,0)                                  src/feMeshRead.cpp:1412-1431
#   tetrahedron faces {0,2,1},{0,1,3},{0,3,2},{3,1,2}                 src/feTetrahedron.h:30
TRI_EDGES = np.array([[0, 1], [1, 2], [2, 0]], np.int32)
TET_FACES = np.array([[0, 2, 1], [0, 1, 3], [0, 3, 2], [3, 1, 2]], np.int32)
#   tetrahedron edges {0,2},{2,1},{1,0},{1,3},{3,0},{3,2}             src/feTetrahedron.h:31
TET_EDGES = np.array([[0, 2], [2, 1], [1, 0], [1, 3], [3, 0], [3, 2]], np.int32)


@dataclass
class Mesh:
    dim: int
    xyz: np.ndarray                 # (nV, 3) float64
    cells: np.ndarray               # (nE, dim+1) int32   physical "Domaine"
    bfacets: np.ndarray             # (nB, dim) int32     physical "Bord", oriented like the adjacent cell
    point_pressure: int | None = 0  # vertex carrying the 0-D physical "PointPression"
    names: dict = field(default_factory=lambda: {"domain": "Domaine", "boundary": "Bord",
                                                 "point": "PointPression"})

    @property
    def n_vertices(self):
        return self.xyz.shape[0]

    @property
    def n_cells(self):
        return self.cells.shape[0]


def boundary_facets(cells: np.ndarray) -> np.ndarray:
    """Facets that belong to exactly one cell, with the node order of that cell's local facet, listed in
    (cell, local facet) order."""
    nv = cells.shape[1]
    loc = TRI_EDGES if nv == 3 else TET_FACES
    fac = cells[:, loc]                                  # (nE, nf, dim)
    flat = fac.reshape(-1, loc.shape[1])
    srt = np.sort(flat.astype(np.int64), axis=1)
    nV = np.int64(cells.max()) + 1
    key = srt[:, 0]
    for c in range(1, srt.shape[1]):
        key = key * nV + srt[:, c]
    _, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
    return np.ascontiguousarray(flat[cnt[inv.reshape(-1)] == 1]).astype(np.int32)


def rect_mesh(nx: int, ny: int, lx: float = 1.0, ly: float = 1.0, x0: float = 0.0, y0: float = 0.0,
              cut_bottom: bool = False, cut_top: bool = False) -> Mesh:
    """nx x ny cells on [x0, x0+lx] x [y0, y0+ly], each split along the same diagonal into two CCW triangles;
    vertices numbered row-major (x fastest).  cut_bottom / cut_top mark the lower / upper side as an artificial
    partition cut (multi-GPU strips): its facets are left out of the physical boundary "Bord"."""
    i, j = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="xy")
    xyz = np.zeros(((nx + 1) * (ny + 1), 3))
    xyz[:, 0] = x0 + lx * i.reshape(-1) / nx
    xyz[:, 1] = y0 + ly * j.reshape(-1) / ny
    ci, cj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v00 = (cj * (nx + 1) + ci).reshape(-1)
    v10, v01, v11 = v00 + 1, v00 + (nx + 1), v00 + (nx + 2)
    cells = np.empty((2 * nx * ny, 3), np.int32)
    cells[0::2] = np.stack([v00, v10, v11], 1)
    cells[1::2] = np.stack([v00, v11, v01], 1)
    bf = boundary_facets(cells)
    if cut_bottom or cut_top:
        row = bf // (nx + 1)
        keep = np.ones(bf.shape[0], bool)
        if cut_bottom:
            keep &= ~np.all(row == 0, axis=1)
        if cut_top:
            keep &= ~np.all(row == ny, axis=1)
        bf = np.ascontiguousarray(bf[keep])
    return Mesh(2, xyz, cells, bf, 0 if not cut_bottom else None)


def square_mesh(n: int, lx: float = 1.0, ly: float = 1.0, x0: float = 0.0, y0: float = 0.0) -> Mesh:
    """T2D(n): 2 n^2 triangles, (n+1)^2 vertices numbered row-major (x fastest)."""
    return rect_mesh(n, n, lx, ly, x0, y0)


# Kuhn split of the unit cube: one tetrahedron per permutation of the axes, all sharing the main diagonal.
_KUHN = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]


def box_mesh(nx: int, ny: int, nz: int, lz: float = 1.0, z0: float = 0.0, cut_bottom: bool = False,
             cut_top: bool = False) -> Mesh:
    """nx x ny x nz cells on [0,1]^2 x [z0, z0+lz], each split into 6 positively oriented Kuhn tetrahedra; vertices
    numbered x fastest, then y, then z.  cut_bottom / cut_top mark the z = z0 / z = z0+lz side as an artificial
    partition cut (multi-GPU slabs): its facets are left out of the physical boundary "Bord"."""
    k, j, i = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    xyz = np.stack([i.reshape(-1) / float(nx), j.reshape(-1) / float(ny), z0 + lz * k.reshape(-1) / float(nz)], 1)
    ck, cj, ci = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    base = np.stack([ci.reshape(-1), cj.reshape(-1), ck.reshape(-1)], 1)          # (nc, 3)
    stride = np.array([1, nx + 1, (nx + 1) * (ny + 1)])
    cells = np.empty((base.shape[0], 6, 4), np.int64)
    for t, perm in enumerate(_KUHN):
        p = base.copy()
        cells[:, t, 0] = p @ stride
        for s, ax in enumerate(perm):
            p = p.copy()
            p[:, ax] += 1
            cells[:, t, s + 1] = p @ stride
    cells = cells.reshape(-1, 4)
    # make every tetrahedron positively oriented (det > 0) by swapping the last two vertices when needed
    a = xyz[cells[:, 1]] - xyz[cells[:, 0]]
    b = xyz[cells[:, 2]] - xyz[cells[:, 0]]
    d = xyz[cells[:, 3]] - xyz[cells[:, 0]]
    det = np.einsum("ij,ij->i", np.cross(a, b), d)
    neg = det < 0
    cells[neg, 2], cells[neg, 3] = cells[neg, 3].copy(), cells[neg, 2].copy()
    cells = cells.astype(np.int32)
    bf = boundary_facets(cells)
    if cut_bottom or cut_top:
        layer = bf // ((nx + 1) * (ny + 1))
        keep = np.ones(bf.shape[0], bool)
        if cut_bottom:
            keep &= ~np.all(layer == 0, axis=1)
        if cut_top:
            keep &= ~np.all(layer == nz, axis=1)
        bf = np.ascontiguousarray(bf[keep])
    return Mesh(3, np.ascontiguousarray(xyz), cells, bf, 0 if not cut_bottom else None)


def cube_mesh(n: int) -> Mesh:
    """T3D(n): 6 n^3 positively oriented tetrahedra, (n+1)^3 vertices (x fastest, then y, then z)."""
    return box_mesh(n, n, n)


def unstructured_tet_mesh(n: int, seed: int = 0, jitter: float = 0.3) -> Mesh:
    """Unstructured tetrahedral mesh of the unit cube: Delaunay triangulation (scipy) of an (n+1)^3 lattice whose INTERIOR points
    are moved by up to `jitter` cells; slivers are dropped, every tetrahedron is positively oriented.  Vertex and edge valences
    vary from node to node (3 ... 9 tetrahedra around an edge), unlike the Kuhn meshes: the parity tests use it to exercise the
    signature sort and the idle lanes of the row-lane kernels."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    g = np.linspace(0.0, 1.0, n + 1)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    interior = np.all((pts > 1e-12) & (pts < 1 - 1e-12), axis=1)
    pts[interior] += (rng.random((int(interior.sum()), 3)) - 0.5) * 2.0 * jitter / n
    cells = Delaunay(pts).simplices.astype(np.int64)
    P = pts[cells]
    vol = np.einsum("ei,ei->e", np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]), P[:, 3] - P[:, 0]) / 6.0
    keep = np.abs(vol) > 1e-3 / (6.0 * n ** 3)
    cells, vol = cells[keep], vol[keep]
    neg = vol < 0
    cells[neg] = cells[neg][:, [0, 2, 1, 3]]
    cells = np.ascontiguousarray(cells).astype(np.int32)
    # vertices that lost all their tetrahedra would be orphan unknowns: the lattice + jitter never produces any, but check
    assert np.unique(cells).size == pts.shape[0]
    return Mesh(3, np.ascontiguousarray(pts), cells, boundary_facets(cells), point_pressure=0)


def write_msh(mesh: Mesh, path: str) -> None:
    """Gmsh 4.1 ASCII with physical groups Domaine / Bord / PointPression, one geometric entity each."""
    dim, nV, nE, nB = mesh.dim, mesh.n_vertices, mesh.n_cells, mesh.bfacets.shape[0]
    has_pt = mesh.point_pressure is not None
    nm = mesh.names
    lo, hi = mesh.xyz.min(0), mesh.xyz.max(0)
    box = " ".join(repr(float(v)) for v in (*lo, *hi))
    out = ["$MeshFormat", "4.1 0 8", "$EndMeshFormat", "$PhysicalNames", str(2 + int(has_pt))]
    if has_pt:
        out.append(f'0 1 "{nm["point"]}"')
    out.append(f'{dim - 1} 2 "{nm["boundary"]}"')
    out.append(f'{dim} 3 "{nm["domain"]}"')
    out.append("$EndPhysicalNames")
    out.append("$Entities")
    n_pts = 1 if has_pt else 0
    if dim == 2:
        out.append(f"{n_pts} 1 1 0")
    else:
        out.append(f"{n_pts} 0 1 1")
    if has_pt:
        p = [float(v) for v in mesh.xyz[mesh.point_pressure]]
        out.append(f"1 {p[0]!r} {p[1]!r} {p[2]!r} 1 1 ")
    out.append(f"1 {box} 1 2 0 ")          # boundary entity (curve in 2D, surface in 3D)
    out.append(f"1 {box} 1 3 1 1 ")        # domain entity, bounded by the boundary entity
    out.append("$EndEntities")
    out.append("$Nodes")
    out.append(f"1 {nV} 1 {nV}")
    out.append(f"{dim} 1 0 {nV}")
    out.extend(str(t) for t in range(1, nV + 1))
    out.extend(f"{x!r} {y!r} {z!r}" for x, y, z in mesh.xyz.tolist())
    out.append("$EndNodes")
    out.append("$Elements")
    n_blocks = 2 + int(has_pt)
    n_tot = nE + nB + int(has_pt)
    out.append(f"{n_blocks} {n_tot} 1 {n_tot}")
    tag = 1
    if has_pt:
        out.append("0 1 15 1")
        out.append(f"{tag} {mesh.point_pressure + 1}")
        tag += 1
    btype, ctype = (1, 2) if dim == 2 else (2, 4)
    out.append(f"{dim - 1} 1 {btype} {nB}")
    for row in (mesh.bfacets + 1).tolist():
        out.append(f"{tag} " + " ".join(map(str, row)))
        tag += 1
    out.append(f"{dim} 1 {ctype} {nE}")
    for row in (mesh.cells + 1).tolist():
        out.append(f"{tag} " + " ".join(map(str, row)))
        tag += 1
    out.append("$EndElements")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
