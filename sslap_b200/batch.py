"""auction_solve_batch — many independent assignment problems in one call (BASELINE.json configs[4]).

The reference has no batch entry point: the equivalent is a Python loop of
``auction_solve(loc=, val=, size=, cardinality_check=False)`` calls (/root/reference/sslap/auction_solve.py:45-46).
Here the whole batch is ONE C-ABI call (sslapb_auction_batch): the problems are laid out block-diagonally in one CSR and
one warp runs each problem's complete eps-scaling auction, so thousands of round chains progress concurrently.
Each result is bit-identical to what the single-problem path (and the reference) returns for that problem.
"""
import ctypes as C

import numpy as np

from . import _native as nat
from .auction_solve import _as_index_array, _as_values, _meta_dict


def pack_problems(problems):
    """Concatenate an iterable of ``(loc, val, size)`` into the packed form ``auction_solve_batch`` marshals to the C ABI:
    dict(loc (K,2), val (K,), nnz_offsets int64[P+1], n_rows int32[P], n_cols int32[P]).  Pack once, solve many times."""
    locs, vals, n_rows, n_cols = [], [], [], []
    for (loc, val, size) in problems:
        loc = _as_index_array(loc)
        val = _as_values(val)
        if loc.shape[0] != val.shape[0]:
            raise ValueError("loc and val must have the same length")
        locs.append(loc)
        vals.append(val)
        n_rows.append(int(size[0]))
        n_cols.append(int(size[1]))
    p = len(locs)
    dtype = np.int64 if any(l.dtype == np.int64 for l in locs) else np.int32
    nnz_off = np.zeros(p + 1, dtype=np.int64)
    if p:
        np.cumsum([l.shape[0] for l in locs], out=nnz_off[1:])
    return dict(loc=np.ascontiguousarray(np.concatenate([l.astype(dtype, copy=False) for l in locs], axis=0)) if p else
                np.zeros((0, 2), dtype=np.int32),
                val=np.ascontiguousarray(np.concatenate(vals)) if p else np.zeros(0),
                nnz_offsets=nnz_off, n_rows=np.asarray(n_rows, dtype=np.int32), n_cols=np.asarray(n_cols, dtype=np.int32))


def auction_solve_batch(problems, problem: str = 'min', eps_start: float = 0., max_iter: int = 1000000,
                        fast: bool = False, packed_result: bool = False, _handle: nat.Handle = None):
    """Solve every problem of ``problems`` — an iterable of ``(loc, val, size)`` with ``loc`` (K x 2, row-sorted, indices
    local to the problem), ``val`` (K float64) and ``size = (rows, cols)``, or the dict returned by ``pack_problems`` —
    and return a list of ``{'sol', 'meta'}`` dicts in the same order, with the reference's keys and rounding
    (``packed_result=True``: one dict with the concatenated ``sol``, ``row_offsets`` and the raw per-problem metas).
    ``problem``, ``eps_start``, ``max_iter`` and ``fast`` apply to every problem and mean what they mean in
    ``auction_solve``; no cardinality check is run (pass feasible problems, or check them with ``hopcroft_solve``)."""
    h = _handle or nat.default_handle()
    L = nat.load()
    pk = problems if isinstance(problems, dict) else pack_problems(problems)
    loc_all, val_all, nnz_off = pk["loc"], pk["val"], pk["nnz_offsets"]
    n_rows, n_cols = pk["n_rows"], pk["n_cols"]
    p = int(n_rows.shape[0])
    if p == 0:
        return []
    row_off = np.zeros(p + 1, dtype=np.int64)
    np.cumsum(n_rows, out=row_off[1:])
    eps = None
    if fast:                                              # auction_.pyx:592,614-615: eps_start = 1/N with `M, N = size`
        eps = (np.float32(1.0) / n_cols.astype(np.float64)).astype(np.float32)
    elif eps_start > 0:
        eps = np.full(p, np.float32(eps_start), dtype=np.float32)
    sol = np.empty(int(row_off[-1]), dtype=np.int32)
    metas = (nat.Meta * p)()
    ib = loc_all.dtype.itemsize
    rc = L.sslapb_auction_batch(h.ptr, p, nnz_off.ctypes.data, n_rows.ctypes.data, n_cols.ctypes.data,
                                loc_all.ctypes.data, loc_all.ctypes.data + ib, ib, 2, val_all.ctypes.data,
                                int(problem != 'min'), eps.ctypes.data if eps is not None else None, int(max_iter),
                                nat.MEM_HOST, sol.ctypes.data, metas)
    if rc == nat.E_FEWER_THAN_N:
        raise ValueError("Matrix is infeasible - a problem of the batch has fewer valid values than rows.")
    if rc == nat.E_EMPTY_ROW:
        raise ValueError("Matrix is infeasible - some row of some problem has no valid value.")
    if rc == nat.E_OUT_OF_RANGE:
        raise ValueError("a problem holds an index outside its own `size`.")
    nat.check(h, rc, "auction_solve_batch")
    if rc != 0:
        raise RuntimeError(f"auction_solve_batch failed with code {rc}: {h.last_error()}")
    if packed_result:
        return dict(sol=sol, row_offsets=row_off, metas=metas)
    return [dict(sol=sol[row_off[k]:row_off[k + 1]].copy(), meta=_meta_dict(metas[k])) for k in range(p)]
