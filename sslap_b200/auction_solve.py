"""auction_solve — same surface as the reference's (/root/reference/sslap/auction_solve.py:6-55), body = marshal
the arrays and make ONE call into the C ABI (sslapb_auction_coo / sslapb_auction_dense); the COO/dense -> CSR build,
the feasibility check and every auction round run on the GPU.
"""
import ctypes as C

import numpy as np

from . import _native as nat

_FORMAT_ERROR = "One of the following formats is expected as input to auction solve: mat OR (loc & val) OR coo_mat."


def _as_index_array(loc):
    """(K,2) C-contiguous int32 or int64 view/copy of `loc` (the reference casts to int32, auction_.pyx:601)."""
    loc = np.asarray(loc)
    if loc.ndim != 2 or loc.shape[1] != 2:
        raise ValueError("loc must have shape (K, 2)")
    if loc.dtype not in (np.int32, np.int64):
        if not np.issubdtype(loc.dtype, np.integer):
            raise ValueError("loc must be an integer array")
        loc = loc.astype(np.int64)
    return np.ascontiguousarray(loc)


def _as_values(val):
    """The reference takes `val` through a float64 typed buffer (auction_.pyx:602): other dtypes raise ValueError."""
    val = np.asarray(val)
    if val.dtype != np.float64:
        raise ValueError(f"Buffer dtype mismatch, expected 'float64' but got '{val.dtype}'")
    if val.ndim != 1:
        raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % val.ndim)
    return np.ascontiguousarray(val)


def _meta_dict(m: nat.Meta) -> dict:
    """AuctionSolver.meta with the reference's keys, types and rounding (auction_.pyx:264,297-304)."""
    return {
        "start_eps": round(float(m.start_eps), 3),
        "eCE": int(m.eCE),
        "its": int(m.its),
        "nreductions": int(m.nreductions),
        "soln_found": int(m.soln_found),
        "n_assigned": int(m.n_assigned),
        "obj": round(float(m.obj), 3),
        "final_eps": round(float(m.final_eps), 3),
        "timer": {"setup": f"{m.setup_ms:.2f}ms", "solve": f"{m.solve_ms:.2f}ms"},
    }


def _raise_infeasible(rc: int, m: nat.Meta, n_rows_msg: int, h: nat.Handle):
    if rc == nat.E_FEWER_THAN_N:                          # auction_.pyx:560,605
        raise ValueError(f"Matrix is infeasible - Fewer than {n_rows_msg} valid values provided for {n_rows_msg} rows.")
    if rc == nat.E_CARDINALITY:                           # auction_.pyx:566,612
        raise ValueError(f"Matrix is infeasible (Maximum matching possible only involves {int(m.cardinality)} "
                         f"out of {n_rows_msg} rows.)")
    if rc == nat.E_EMPTY_ROW:
        raise ValueError("Matrix is infeasible - some row has no valid value (cardinality_check=False would be "
                         "undefined behaviour in the reference).")
    if rc == nat.E_UNSORTED:
        raise ValueError("loc must be sorted by row (precondition of the reference, auction_.pyx:33-48).")
    if rc == nat.E_OUT_OF_RANGE:
        raise ValueError("loc holds an index outside the matrix described by `size`.")
    nat.check(h, rc, "auction_solve")
    if rc != 0:
        raise RuntimeError(f"auction_solve failed with code {rc}: {h.last_error()}")


def auction_solve(mat: np.ndarray = None, loc: np.ndarray = None, val: np.ndarray = None, coo_mat=None,
                  problem: str = 'min', eps_start: float = 0.,
                  max_iter: int = 1000000, fast: bool = False, size=None, cardinality_check=True,
                  _handle: nat.Handle = None, _raw_meta: bool = False, prices_in: np.ndarray = None,
                  return_prices: bool = False) -> dict:
    """Solve an assignment problem i -> j with the eps-scaling auction algorithm on the GPU.

    Inputs, keywords, return value and errors are those of the reference's ``sslap.auction_solve``:
    ``mat`` (N x M float64, -1 / negative = no edge), or ``loc`` (K x 2 row-sorted indices) + ``val`` (K float64), or a
    scipy ``coo_mat``; ``problem`` 'min' minimises, anything else maximises (auction_.pyx:236); ``eps_start``,
    ``max_iter``, ``fast``, ``size`` and ``cardinality_check`` as in auction_solve.py:26-35.
    Returns ``{'sol': int32[N], 'meta': {...}}`` (-1 in ``sol`` only if ``max_iter`` was hit).
    Unlike the reference, the caller's ``val`` is never negated in place.

    Extensions (keyword only, not in the reference): ``prices_in`` — warm start from these object prices (as returned by
    an earlier call with ``return_prices=True``) instead of zeros (auction_.pyx:220), usually together with ``eps_start``;
    ``return_prices`` — add ``'prices'`` (float64, one per column, AuctionSolver.p) to the result.
    """
    h = _handle or nat.default_handle()
    if prices_in is None:
        return _dispatch(h, mat, loc, val, coo_mat, problem, eps_start, max_iter, fast, size, cardinality_check,
                         _raw_meta, return_prices)
    h.set_prices(prices_in)
    try:
        return _dispatch(h, mat, loc, val, coo_mat, problem, eps_start, max_iter, fast, size, cardinality_check,
                         _raw_meta, return_prices)
    finally:
        h.set_prices(None)                                # consumed by a successful solve; cleared after a failed one


def _dispatch(h, mat, loc, val, coo_mat, problem, eps_start, max_iter, fast, size, cardinality_check, _raw_meta,
              return_prices):
    """Format dispatch of auction_solve.py:41-55."""
    L = nat.load()
    maximize = int(problem != 'min')
    m = nat.Meta()
    if mat is not None:                                   # _from_matrix, auction_.pyx:528-571
        mat = np.asarray(mat)
        if mat.dtype != np.float64:
            raise ValueError(f"Buffer dtype mismatch, expected 'double' but got '{mat.dtype}'")
        if mat.ndim != 2:
            raise ValueError("Buffer has wrong number of dimensions (expected 2, got %d)" % mat.ndim)
        mat = np.ascontiguousarray(mat)
        n, mm = mat.shape
        if fast:
            eps_start = float(np.float32(1.0 / n))        # :568-569
        sol = np.empty(n, dtype=np.int32)
        rc = L.sslapb_auction_dense(h.ptr, mat.ctypes.data, n, mm, maximize, float(np.float32(eps_start)),
                                    int(max_iter), int(bool(cardinality_check)), nat.MEM_HOST, sol.ctypes.data,
                                    C.byref(m))
        n_msg = n
    elif loc is not None and val is not None:
        return _solve_sparse(h, L, loc, val, size, maximize, eps_start, max_iter, fast, cardinality_check, _raw_meta,
                             return_prices)
    elif coo_mat is not None:                             # auction_solve.py:47-50
        loc = np.stack([coo_mat.row, coo_mat.col], axis=-1)
        return _solve_sparse(h, L, loc, coo_mat.data, coo_mat.shape, maximize, eps_start, max_iter, fast,
                             cardinality_check, _raw_meta, return_prices)
    else:
        raise ValueError(_FORMAT_ERROR)
    if rc != 0:
        _raise_infeasible(rc, m, n_msg, h)
    out = dict(sol=sol, meta=_meta_dict(m))
    if _raw_meta:
        out["raw"] = m
    if return_prices:
        out["prices"] = h.get_prices(m.n_cols)
    return out


def _solve_sparse(h, L, loc, val, size, maximize, eps_start, max_iter, fast, cardinality_check, raw_meta,
                  return_prices=False):
    """_from_sparse, auction_.pyx:575-617."""
    loc = _as_index_array(loc)
    val = _as_values(val)
    if val.shape[0] != loc.shape[0]:
        raise ValueError("loc and val must have the same length")
    k = loc.shape[0]
    if k == 0:
        raise ValueError("Matrix is infeasible - Fewer than 1 valid values provided for 1 rows.")
    if size is not None:
        n_rows, n_cols = int(size[0]), int(size[1])
        n_chk = int(size[1])                              # the reference unpacks `M, N = size` (:592)
    else:
        n_rows, n_cols = int(loc[:, 0].max()) + 1, int(loc[:, 1].max()) + 1
        n_chk = n_rows - 1                                # :594 takes max() without the +1
    if fast:
        eps_start = float(np.float32(1.0 / n_chk))        # :614-615
    sol = np.empty(n_rows, dtype=np.int32)
    m = nat.Meta()
    ib = loc.dtype.itemsize
    rc = L.sslapb_auction_coo(h.ptr, loc.ctypes.data, loc.ctypes.data + ib, ib, 2, val.ctypes.data, k, n_rows, n_cols,
                              maximize, float(np.float32(eps_start)), int(max_iter), int(bool(cardinality_check)),
                              nat.MEM_HOST, sol.ctypes.data, C.byref(m))
    if rc != 0:
        _raise_infeasible(rc, m, n_rows, h)
    out = dict(sol=sol, meta=_meta_dict(m))
    if raw_meta:
        out["raw"] = m
    if return_prices:
        out["prices"] = h.get_prices(m.n_cols)
    return out
