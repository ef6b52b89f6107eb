"""Quadrature rules and reference-basis tabulation at the quadrature nodes (host side, numpy).

The engine receives these tables as plain arrays through the C ABI (b200_set_quadrature / b200_add_space); in a
drop-in build the adapter takes them from the host objects (feSpace::_L, _dLdr, ..., src/feSpace.h:140-147).  For
the synthetic benchmarks they are produced here.  Layout is the reference's: table[k][i] = function i at quadrature
node k (src/feSpace.h:381-384).

Bases (Lagrange, reference simplex r,s(,t) >= 0, r+s(+t) <= 1):
  TriP1  src/feSpace_2D.cpp:118-131     TriP2  src/feSpace_2D.cpp:536-544, 727-737
  TetP1/TetP2: barycentric Lagrange, vertex functions then the 6 edge functions in the order
  {0,2},{2,1},{1,0},{1,3},{3,0},{3,2} (src/feTetrahedron.h:31, src/feSpace_3D.cpp:41-).
"""
from __future__ import annotations

import json
import os

import numpy as np

from .mesh import TET_EDGES

_QUAD = None


def quadrature(dim: int, degree: int):
    """(w, pts[nq, dim]) of the symmetric rule the reference selects for this degree."""
    global _QUAD
    if _QUAD is None:
        with open(os.path.join(os.path.dirname(__file__), "data", "quadrature.json")) as f:
            _QUAD = json.load(f)
    q = _QUAD["tri" if dim == 2 else "tet"][str(degree)]
    cols = [q["r"], q["s"]] + ([q["t"]] if dim == 3 else [])
    return np.array(q["w"]), np.ascontiguousarray(np.array(cols).T)


def basis(dim: int, order: int, pts: np.ndarray):
    """L[nq, nF], dL[nq, nF, dim] (reference derivatives) of the scalar Lagrange basis."""
    pts = np.asarray(pts, float)
    nq = pts.shape[0]
    r, s = pts[:, 0], pts[:, 1]
    if dim == 2:
        if order == 1:
            L = np.stack([1.0 - r - s, r, s], 1)
            dL = np.zeros((nq, 3, 2))
            dL[:, 0] = [-1.0, -1.0]
            dL[:, 1] = [1.0, 0.0]
            dL[:, 2] = [0.0, 1.0]
            return L, dL
        if order == 2:
            L = np.stack([(1. - r - s) * (1. - 2. * r - 2. * s), r * (2. * r - 1.), s * (2. * s - 1.),
                          4. * r * (1. - r - s), 4. * r * s, 4. * s * (1. - r - s)], 1)
            z = np.zeros(nq)
            dLdr = np.stack([4. * (r + s) - 3., 4. * r - 1., z, 4. * (1. - 2. * r - s), 4. * s, -4. * s], 1)
            dLds = np.stack([4. * (r + s) - 3., z, 4. * s - 1., -4. * r, 4. * r, 4. * (1. - r - 2. * s)], 1)
            return L, np.stack([dLdr, dLds], 2)
    if dim == 3:
        t = pts[:, 2]
        lam = np.stack([1.0 - r - s - t, r, s, t], 1)                  # (nq, 4)
        dlam = np.array([[-1., -1., -1.], [1., 0., 0.], [0., 1., 0.], [0., 0., 1.]])
        if order == 1:
            return lam, np.broadcast_to(dlam, (nq, 4, 3)).copy()
        if order == 2:
            L = np.zeros((nq, 10))
            dL = np.zeros((nq, 10, 3))
            for i in range(4):
                L[:, i] = lam[:, i] * (2. * lam[:, i] - 1.)
                dL[:, i] = (4. * lam[:, i] - 1.)[:, None] * dlam[i]
            for e, (a, b) in enumerate(TET_EDGES):
                L[:, 4 + e] = 4. * lam[:, a] * lam[:, b]
                dL[:, 4 + e] = 4. * (lam[:, a][:, None] * dlam[b] + lam[:, b][:, None] * dlam[a])
            return L, dL
    raise ValueError(f"unsupported Lagrange space dim={dim} order={order}")
