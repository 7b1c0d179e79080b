"""Synthetic inputs of the benchmark configurations that are not stencils (host side, numpy only).

C5 (BASELINE.json config 5): a COO matrix with power-law row lengths,
    N = 2^k rows (2^22 in the benchmark), row length min(1e5, floor(8 U^(-1/1.5))) with U uniform (Pareto tail, mean ~ 24,
    i.e. ~1e8 non-zeros at N = 2^22), columns uniform per row, entries sorted by (row, col), values U(-1, 1), and the
    diagonal forced to 1 + sum |row| so that the matrix is strictly diagonally dominant (GMRES has something to solve),
    all from numpy.random.default_rng(12345).
Columns are drawn WITH replacement and duplicates within a row are then dropped (the recipe says "without replacement";
per-row rejection sampling is not vectorisable over 4 M rows -- the difference is ~0.01 % of the entries of the longest
rows), and every row gets its diagonal entry.  The same arrays feed the GPU (lsk_coo_create) and the oracle.
"""
from __future__ import annotations

import numpy as np


def power_law_coo(log2_n: int = 22, seed: int = 12345, mean_scale: float = 8.0, max_row: int = 100_000):
    """Returns (n, entry[f64], row[i64], col[i64]) sorted by (row, col)."""
    n = 1 << log2_n
    rng = np.random.default_rng(seed)
    u = np.maximum(rng.random(n), 1e-12)
    length = np.minimum(max_row, np.floor(mean_scale * u ** (-1.0 / 1.5))).astype(np.int64)
    length = np.minimum(length, n)
    total = int(length.sum())
    row = np.repeat(np.arange(n, dtype=np.int64), length)
    col = rng.integers(0, n, size=total, dtype=np.int64)
    # every row has its diagonal; duplicates within a row are dropped; result sorted by (row, col)
    key = np.concatenate([row * n + col, np.arange(n, dtype=np.int64) * (n + 1)])
    del row, col
    key.sort()  # (np.unique on 1e8 int64 is ~100x slower than sort + compare in numpy 2.3)
    keep = np.empty(key.size, dtype=bool)
    keep[0] = True
    np.not_equal(key[1:], key[:-1], out=keep[1:])
    key = key[keep]
    del keep
    row, col = key >> log2_n, key & (n - 1)
    del key
    entry = rng.uniform(-1.0, 1.0, size=row.size)
    diag = row == col
    off = np.where(diag, 0.0, np.abs(entry))
    rowsum = np.bincount(row, weights=off, minlength=n)
    entry[diag] = 1.0 + rowsum
    return n, entry, row, col


def coo_to_csr_rowptr(n: int, row: np.ndarray):
    """Inclusive (lo, hi) rects of global k per row for entries sorted by row (the CSRMatrix layout)."""
    counts = np.bincount(row, minlength=n).astype(np.int64)
    lo = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
    rect = np.empty(n, dtype=np.dtype([("lo", np.int64), ("hi", np.int64)]))
    rect["lo"], rect["hi"] = lo, lo + counts - 1
    return rect
