"""TEST INFRASTRUCTURE ONLY — ctypes front for oracle/sslap_oracle.c (the CPU restatement of sslap v0.2.5's auction
and Hopcroft-Karp).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

`auction_solve` / `hopcroft_solve` mirror the reference signatures (/root/reference/sslap/auction_solve.py:6-55,
check_feasible.py:5-20) closely enough that parity tests read like calls into the reference.
"""
import ctypes as C
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OracleMeta(C.Structure):
    _fields_ = [("start_eps", C.c_float), ("final_eps", C.c_float), ("target_eps", C.c_float),
                ("eCE", C.c_int32), ("soln_found", C.c_int32), ("its", C.c_int64), ("nreductions", C.c_int64),
                ("n_assigned", C.c_int64), ("obj", C.c_float), ("obj64", C.c_double),
                ("entry_visits", C.c_int64), ("total_bids", C.c_int64)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libsslap_oracle.so")
    src = os.path.join(_HERE, "sslap_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libsslap_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.sslap_oracle_auction.restype = C.c_int
        L.sslap_oracle_auction.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int,
                                           C.c_float, C.c_int64, C.c_int, C.c_void_p, C.POINTER(OracleMeta),
                                           C.c_void_p, C.c_void_p]
        L.sslap_oracle_set_ext.restype = None
        L.sslap_oracle_set_ext.argtypes = [C.c_void_p, C.c_int]
        L.sslap_oracle_hopcroft.restype = C.c_int32
        L.sslap_oracle_hopcroft.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.sslap_oracle_bid_sweep.restype = None
        L.sslap_oracle_bid_sweep.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                             C.c_float, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def auction_solve(loc=None, val=None, mat=None, problem="min", eps_start=0.0, max_iter=1000000, fast=False,
                  faithful_scan=False, return_prices=False, frontier_hist=False, prices_in=None, strict=False):
    """CPU oracle of auction_solve (no cardinality check — that is `hopcroft_solve`).  Returns {'sol','meta'} with the
    reference's meta keys and rounding (auction_.pyx:264,297-304) plus unrounded extras under meta['_raw']."""
    if mat is not None:                                   # _from_matrix, auction_.pyx:546-553
        mat = np.asarray(mat, dtype=np.float64)
        r, c = np.nonzero(mat >= 0)
        loc = np.stack([r, c], axis=-1).astype(np.int32)
        val = mat[r, c].astype(np.float64)
    loc = np.ascontiguousarray(loc)
    rows = np.ascontiguousarray(loc[:, 0].astype(np.int32))
    cols = np.ascontiguousarray(loc[:, 1].astype(np.int32))
    val = np.ascontiguousarray(val, dtype=np.float64)
    N = int(rows.max()) + 1                               # auction_.pyx:209-210
    M = int(cols.max()) + 1
    if fast:                                              # auction_.pyx:568-569 / :614-615
        eps_start = float(np.float32(1.0 / N))
    sol = np.empty(N, dtype=np.int32)
    prices = np.empty(M, dtype=np.float64) if return_prices else None
    hist = np.zeros(40, dtype=np.int64) if frontier_hist else None
    meta = OracleMeta()
    if prices_in is not None or strict:                   # extensions beyond the reference: warm start / strict stop rule
        prices_in = None if prices_in is None else np.ascontiguousarray(prices_in, dtype=np.float64)
        assert prices_in is None or prices_in.size == M
        lib().sslap_oracle_set_ext(_ptr(prices_in) if prices_in is not None else None, int(bool(strict)))
    t0 = time.perf_counter()
    rc = lib().sslap_oracle_auction(_ptr(rows), _ptr(cols), _ptr(val), val.size, N, M, int(problem != "min"),
                                    float(np.float32(eps_start)), int(max_iter), int(bool(faithful_scan)), _ptr(sol),
                                    C.byref(meta), _ptr(prices) if prices is not None else None,
                                    _ptr(hist) if hist is not None else None)
    dt = time.perf_counter() - t0
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    out_meta = {
        "start_eps": round(float(meta.start_eps), 3), "eCE": int(meta.eCE), "its": int(meta.its),
        "nreductions": int(meta.nreductions), "soln_found": int(meta.soln_found), "n_assigned": int(meta.n_assigned),
        "obj": round(float(meta.obj), 3), "final_eps": round(float(meta.final_eps), 3),
        "timer": {"setup": "0.00ms", "solve": f"{1000 * dt:.2f}ms"},
        "_raw": {"start_eps": float(meta.start_eps), "final_eps": float(meta.final_eps), "obj64": float(meta.obj64),
                 "target_eps": float(meta.target_eps), "entry_visits": int(meta.entry_visits),
                 "total_bids": int(meta.total_bids), "seconds": dt},
    }
    res = dict(sol=sol, meta=out_meta)
    if return_prices:
        res["prices"] = prices
    if frontier_hist:
        res["frontier_hist"] = hist
    return res


def hopcroft_solve(loc=None, mat=None, lookup=None, N=None, M=None):
    """CPU oracle of hopcroft_solve (feasibility_.pyx:227-283): {'size','left_pairings','right_pairings'}."""
    assert (loc is None) + (mat is None) + (lookup is None) == 2, \
        "Exactly one of the arguments loc, mat, lookup must be provided."
    if mat is not None:
        mat = np.asarray(mat)
        r, c = np.nonzero(mat >= 0)
        loc = np.stack([r, c], axis=-1)
        N, M = mat.shape
    elif lookup is not None:
        loc = np.array([(i, j) for i in lookup for j in lookup[i]], dtype=np.int64).reshape(-1, 2)
        order = np.argsort(loc[:, 0], kind="stable")
        loc = loc[order]
    rows = np.ascontiguousarray(loc[:, 0].astype(np.int32))
    cols = np.ascontiguousarray(loc[:, 1].astype(np.int32))
    N = int(rows.max()) + 1 if N is None else int(N)
    M = int(cols.max()) + 1 if M is None else int(M)
    left = np.empty(N, dtype=np.int32)
    right = np.empty(M, dtype=np.int32)
    size = lib().sslap_oracle_hopcroft(_ptr(rows), _ptr(cols), rows.size, N, M, _ptr(left), _ptr(right))
    return dict(size=int(size), left_pairings=left, right_pairings=right)


def bid_sweep(rowptr, cols, val_folded, prices, bidders, eps):
    """Kernel-level oracle for one bidding sweep (auction_.pyx:339-365)."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    val_folded = np.ascontiguousarray(val_folded, dtype=np.float64)
    prices = np.ascontiguousarray(prices, dtype=np.float64)
    bidders = np.ascontiguousarray(bidders, dtype=np.int32)
    jb = np.empty(bidders.size, dtype=np.int32)
    bd = np.empty(bidders.size, dtype=np.float64)
    lib().sslap_oracle_bid_sweep(_ptr(rowptr), _ptr(cols), _ptr(val_folded), _ptr(prices), _ptr(bidders),
                                 bidders.size, float(np.float32(eps)), _ptr(jb), _ptr(bd))
    return jb, bd
