"""CSR sparsity pattern with the rules of the reference's feEZCompressedRowStorage (host side, numpy).

Rules restated from src/feCompressedRowStorage.cpp:15-133:
  * every row holds its diagonal, even when no form couples it (:33)
  * a (row, col) pair of a matrix form's adrI x adrJ is kept only if both are unknowns, i.e. < nInc (:80)
  * periodic (slave, master) pairs add a column `master` to row `slave` (:98-107)
  * columns ascending and unique per row (:110-116), ia/ja zero-based (:118-133)
The reference builds it by a dry assembly over colours and elements; the order of insertion is irrelevant after the
sort, so this version simply sorts packed (row, col) keys chunk by chunk.
"""
from __future__ import annotations

import numpy as np


def build_pattern(n_inc: int, couplings, periodic=None, chunk: int = 1 << 18):
    """couplings: iterable of (adrI[nE, M], adrJ[nE, N]) for every form that has a matrix.
    Returns (ia int64[n_inc+1], ja int32[nnz])."""
    n = np.int64(n_inc)
    keys = [np.arange(n_inc, dtype=np.int64) * n + np.arange(n_inc, dtype=np.int64)]
    for adrI, adrJ in couplings:
        nE = adrI.shape[0]
        for b in range(0, nE, chunk):
            I = adrI[b:b + chunk].astype(np.int64)
            J = adrJ[b:b + chunk].astype(np.int64)
            k = I[:, :, None] * n + J[:, None, :]
            ok = (I[:, :, None] < n) & (J[:, None, :] < n)
            keys.append(np.unique(k[ok]))
        keys = [np.unique(np.concatenate(keys))]
    if periodic:
        extra = [s * n + m for m, s in periodic if s < n_inc and m < n_inc]
        if extra:
            keys = [np.unique(np.concatenate(keys + [np.array(extra, np.int64)]))]
    key = keys[0]
    rows = key // n
    ja = (key - rows * n).astype(np.int32)
    ia = np.zeros(n_inc + 1, np.int64)
    np.cumsum(np.bincount(rows, minlength=n_inc), out=ia[1:])
    return ia, ja


def slot_map(ia: np.ndarray, ja: np.ndarray, adrI: np.ndarray, adrJ: np.ndarray, chunk: int = 1 << 18):
    """CSR slot of every local entry: slots[e, i, j] = position of (adrI[e,i], adrJ[e,j]) in ja, or -1 when either
    DOF is essential (>= nInc).  This is the precomputed form of the linear row scan the reference performs on
    every scatter (src/feLinearSystemMklPardiso.cpp:619-660)."""
    n_inc = ia.shape[0] - 1
    n = np.int64(n_inc)
    key = np.repeat(np.arange(n_inc, dtype=np.int64), np.diff(ia)) * n + ja.astype(np.int64)
    nE, M = adrI.shape
    N = adrJ.shape[1]
    dt = np.int32 if ja.shape[0] < 2 ** 31 else np.int64
    out = np.empty((nE, M, N), dt)
    for b in range(0, nE, chunk):
        I = adrI[b:b + chunk].astype(np.int64)[:, :, None]
        J = adrJ[b:b + chunk].astype(np.int64)[:, None, :]
        ok = (I < n) & (J < n)
        q = np.where(ok, I * n + J, 0)
        pos = np.searchsorted(key, q)
        out[b:b + chunk] = np.where(ok, pos, -1)
    return out
