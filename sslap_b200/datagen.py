"""Seeded synthetic sparse assignment problems (host-side, numpy) for tests and bench.py.

The recipe mirrors the reference's benchmarking harness (/root/reference/benchmarking.py:29-45: uniform(0,100) float
or integer 1..100 costs, random sparsity mask, every row/column kept feasible) but is generated directly in COO so
that 1M x 1M instances are possible; a planted permutation guarantees a perfect matching exists.  Output is row-major
sorted and duplicate free — the reference's precondition for `loc` (auction_.pyx:33-48).
"""
import numpy as np


def make_problem(n: int, density: float, mode: str = "float", seed: int = 0, m: int = None):
    """Return (loc int32 (K,2), val float64 (K,)) for an n x m (default square) problem."""
    m = n if m is None else m
    assert m >= n
    rng = np.random.default_rng(seed)
    k = int(round(n * m * density))
    r = rng.integers(0, n, k, dtype=np.int64)
    c = rng.integers(0, m, k, dtype=np.int64)
    perm = rng.permutation(m)[:n].astype(np.int64)
    r = np.concatenate([r, np.arange(n, dtype=np.int64)])
    c = np.concatenate([c, perm])
    key = np.unique(r * m + c)
    loc = np.empty((key.size, 2), dtype=np.int32)
    loc[:, 0] = key // m
    loc[:, 1] = key % m
    if mode == "int":
        val = rng.integers(1, 101, key.size).astype(np.float64)
    elif mode == "float":
        val = rng.uniform(0.0, 100.0, key.size)
    else:
        raise ValueError("mode must be 'int' or 'float'")
    return loc, val


def objective(loc, val, sol):
    """float64 objective of an assignment computed from the original values (the reference's meta['obj'] is float32,
    auction_.pyx:489, so parity tests never compare that key alone)."""
    n = int(loc[:, 0].max()) + 1
    m = int(loc[:, 1].max()) + 1
    key = loc[:, 0].astype(np.int64) * m + loc[:, 1]
    want = np.arange(n, dtype=np.int64) * m + np.asarray(sol, dtype=np.int64)
    idx = np.searchsorted(key, want)
    ok = (idx < key.size) & (key[np.minimum(idx, key.size - 1)] == want)
    if not ok.all():
        raise ValueError("assignment uses an entry that is not in the matrix")
    return float(val[idx].sum())
