"""Element colouring with the rule of feCncGeo::colorElements(1) (src/feCncGeo.cpp:752-794).

Each sweep visits the still uncoloured elements in index order and takes every element none of whose vertex neighbours
was taken earlier in the same sweep (the reference marks the neighbours of a taken element -2 and resets them after the
sweep); the elements taken by sweep c get colour c.  The result is what `b200_set_colors` expects for the coloured scatter
(B200_SCATTER_COLORED) and is bit-identical to the reference's `_elmToColor` (tests/test_host_tables.py).

Host set-up code: the sweep is inherently sequential; numba compiles it when available (it is in this image), otherwise the
same loop runs in Python (fine for the parity sizes)."""
from __future__ import annotations

import numpy as np


def _color_loop(cells, n_vertices):
    nE, nv = cells.shape
    color = np.full(nE, -1, np.int32)
    stamp = np.full(n_vertices, -1, np.int32)      # sweep in which a vertex was last taken
    c = 0
    left = nE
    while left > 0:
        for e in range(nE):
            if color[e] >= 0:
                continue
            free = True
            for j in range(nv):
                if stamp[cells[e, j]] == c:
                    free = False
                    break
            if free:
                color[e] = c
                left -= 1
                for j in range(nv):
                    stamp[cells[e, j]] = c
        c += 1
    return color


try:                                               # pragma: no cover - depends on the image
    import numba
    _color_jit = numba.njit(cache=False)(_color_loop)
except Exception:                                  # noqa: BLE001
    _color_jit = None


def color_elements(cells: np.ndarray, n_vertices: int | None = None) -> np.ndarray:
    """colour[nE] (int32, colours 0 .. nColours-1)."""
    cells = np.ascontiguousarray(cells, np.int64)
    nV = int(cells.max()) + 1 if n_vertices is None else int(n_vertices)
    fn = _color_jit if (_color_jit is not None and cells.shape[0] > 2000) else _color_loop
    return fn(cells, nV)
