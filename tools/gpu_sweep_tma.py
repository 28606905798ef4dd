"""Streamed (TMA) full-frontier sweep vs per-row kernel vs oracle: python tools/gpu_sweep_tma.py [quick]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
import sslap_b200
from sslap_b200 import _native as nat
from sslap_b200.datagen import make_problem
from oracle import oracle
h = nat.default_handle(); L = nat.load()

def sweep(n, prices, merge, iters=1, flush=0, eps=0.37):
    jb = np.empty(n, dtype=np.int32); bd = np.empty(n, dtype=np.float64); ms = C.c_float(0)
    rc = L.sslapb_bid_sweep(h.ptr, prices.ctypes.data if prices is not None else None, None, n, eps, merge, iters, flush,
                            jb.ctypes.data, bd.ctypes.data, C.byref(ms))
    assert rc == 0, (rc, h.last_error())
    return jb, bd, ms.value

ok = True
cases = [(1000, 0.01, "int", 3, None), (4000, 0.02, "float", 4, None), (257, 0.5, "int", 5, None), (3000, 0.4, "float", 6, None),
         (20000, 0.0002, "float", 7, None), (5000, 0.01, "float", 8, 7000), (64, 1.0, "float", 9, None), (1500, 0.9, "int", 10, None)]
for (n, d, mode, seed, m) in cases:
    loc, val = make_problem(n, d, mode, seed=seed, m=m)
    M = m or n
    sslap_b200.auction_solve(loc=loc, val=val, size=(n, M), cardinality_check=False, max_iter=1)
    rng = np.random.default_rng(seed)
    for pk in range(3):
        if pk == 0: prices = rng.integers(0, 40, M).astype(np.float64) if mode == "int" else rng.uniform(0, 50, M)
        elif pk == 1: prices = np.zeros(M)
        else:
            prices = rng.uniform(0, 5, M); prices[rng.integers(0, M, max(1, M // 50))] = np.inf
        rowptr = np.searchsorted(loc[:, 0], np.arange(n + 1)).astype(np.int64)
        oj, ob = oracle.bid_sweep(rowptr, loc[:, 1], -val, prices, np.arange(n, dtype=np.int32), 0.37)
        for merge in (0, 2, 4):
            jb, bd, _ = sweep(n, prices, merge)
            good = np.array_equal(jb, oj) and np.array_equal(bd, ob)
            ok &= good
            if not good:
                bad = np.flatnonzero((jb != oj) | (bd != ob))
                print(f"MISMATCH n={n} d={d} {mode} prices={pk} merge={merge}: {bad.size} rows, first {bad[:5]}, got {jb[bad[:3]]} {bd[bad[:3]]} want {oj[bad[:3]]} {ob[bad[:3]]}", flush=True)
    print(f"case n={n} d={d} {mode} done", flush=True)
print("ALL SWEEP OK" if ok else "SOME SWEEP MISMATCH", flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "quick": sys.exit(0)
n = 100000
loc, val = make_problem(n, 0.001, "float", seed=0)
r = sslap_b200.auction_solve(loc=loc, val=val, size=(n, n), cardinality_check=False)
by = 12 * val.size + 44 * n
j0, b0, _ = sweep(n, None, 0, eps=1e-5)
for merge in (0, 4, 2, 6, 1, 5):
    jb, bd, ms = sweep(n, None, merge, iters=20, flush=1, eps=1e-5)
    print(f"C3 final prices merge={merge} ({'streamed' if merge & 4 else 'per-row'}{', no pruning' if merge & 2 else ''}{', atomics' if merge & 1 else ''}): "
          f"{ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s  same={np.array_equal(jb, j0) and np.array_equal(bd, b0)}", flush=True)
prices = np.zeros(n)
for merge in (0, 4):
    jb, bd, ms = sweep(n, prices, merge, iters=20, flush=1, eps=1e-5)
    print(f"C3 zero prices merge={merge}: {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s", flush=True)
