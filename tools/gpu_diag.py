"""First-contact GPU diagnostics: runs the CUDA path next to the oracle and prints where they diverge.
Usage (on a GPU box): python tools/gpu_diag.py [quick|full]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle
import sslap_b200
from sslap_b200 import _native as nat
from sslap_b200.datagen import make_problem, objective

mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
h = nat.default_handle()
L = nat.load()
import ctypes as C


def gpu_prices(m):
    p = np.empty(m, dtype=np.float64)
    assert L.sslapb_get_prices(h.ptr, p.ctypes.data) == 0
    return p


def compare(name, loc, val, problem="min", t_small=32, max_iter=1000000, **kw):
    h.set_option("t_small", t_small)
    t = time.perf_counter()
    try:
        g = sslap_b200.auction_solve(loc=loc, val=val, size=(int(loc[:, 0].max()) + 1, int(loc[:, 1].max()) + 1),
                                     problem=problem, cardinality_check=False, max_iter=max_iter, _raw_meta=True, **kw)
    except Exception as e:
        print(f"[{name}] t_small={t_small} EXCEPTION {type(e).__name__}: {e}")
        return False
    tg = time.perf_counter() - t
    o = oracle.auction_solve(loc=loc, val=val, problem=problem, max_iter=max_iter, return_prices=True, **kw)
    pg = gpu_prices(int(loc[:, 1].max()) + 1)
    same_sol = np.array_equal(g["sol"], o["sol"])
    same_p = np.array_equal(pg, o["prices"])
    keys = ("start_eps", "eCE", "its", "nreductions", "soln_found", "n_assigned", "obj", "final_eps")
    diffs = {k: (g["meta"][k], o["meta"][k]) for k in keys if g["meta"][k] != o["meta"][k]}
    r = g["raw"]
    ok = same_sol and same_p and not diffs
    print(f"[{name}] t_small={t_small} {'OK ' if ok else 'MISMATCH'} sol={same_sol} prices={same_p} diffs={diffs} "
          f"its={g['meta']['its']} rounds(g/w/s)={r.rounds_grid}/{r.rounds_warp}/{r.rounds_solo} "
          f"solve={r.solve_ms:.3f}ms setup={r.setup_ms:.3f}ms h2d={r.h2d_ms:.3f}ms wall={tg*1e3:.1f}ms "
          f"oracle={o['meta']['_raw']['seconds']*1e3:.1f}ms prof={[round(x,2) for x in r.prof_ms]}", flush=True)
    return ok


def first_divergence(name, loc, val, problem, t_small, hi):
    lo = 0
    while lo + 1 < hi:
        mid = (lo + hi) // 2
        h.set_option("t_small", t_small)
        g = sslap_b200.auction_solve(loc=loc, val=val, size=(int(loc[:, 0].max()) + 1, int(loc[:, 1].max()) + 1),
                                     problem=problem, cardinality_check=False, max_iter=mid)
        o = oracle.auction_solve(loc=loc, val=val, problem=problem, max_iter=mid, return_prices=True)
        pg = gpu_prices(int(loc[:, 1].max()) + 1)
        if np.array_equal(g["sol"], o["sol"]) and np.array_equal(pg, o["prices"]):
            lo = mid
        else:
            hi = mid
    print(f"[{name}] first diverging round (1-based) = {hi}")
    g = sslap_b200.auction_solve(loc=loc, val=val, size=(int(loc[:, 0].max()) + 1, int(loc[:, 1].max()) + 1),
                                 problem=problem, cardinality_check=False, max_iter=hi, _raw_meta=True)
    o = oracle.auction_solve(loc=loc, val=val, problem=problem, max_iter=hi, return_prices=True)
    pg = gpu_prices(int(loc[:, 1].max()) + 1)
    bad = np.nonzero(pg != o["prices"])[0][:10]
    print("   price diffs at", bad, pg[bad], o["prices"][bad])
    bs = np.nonzero(g["sol"] != o["sol"])[0][:10]
    print("   sol diffs at", bs, g["sol"][bs], o["sol"][bs], "n_assigned", g["meta"]["n_assigned"], o["meta"]["n_assigned"],
          "rounds g/w/s", g["raw"].rounds_grid, g["raw"].rounds_warp, g["raw"].rounds_solo)


# ---- 1. kernel-level bid sweep parity
for (n, d, md) in [(1000, 0.01, "int"), (3000, 0.02, "float")]:
    loc, val = make_problem(n, d, md, seed=3)
    h.set_option("t_small", 32)
    sslap_b200.auction_solve(loc=loc, val=val, size=(n, n), cardinality_check=False, max_iter=1)   # makes the CSR resident
    rng = np.random.default_rng(5)
    prices = rng.uniform(0, 50, n)
    bidders = rng.permutation(n).astype(np.int32)[: n // 2]
    jb = np.empty(bidders.size, dtype=np.int32); bd = np.empty(bidders.size, dtype=np.float64); ms = C.c_float(0)
    rc = L.sslapb_bid_sweep(h.ptr, prices.ctypes.data, bidders.ctypes.data, bidders.size, 0.37, 1, 1, 0,
                            jb.ctypes.data, bd.ctypes.data, C.byref(ms))
    rows = loc[:, 0]
    rowptr = np.searchsorted(rows, np.arange(n + 1)).astype(np.int64)
    oj, ob = oracle.bid_sweep(rowptr, loc[:, 1], -val, prices, bidders, 0.37)
    print(f"[bid_sweep n={n} {md}] rc={rc} j_equal={np.array_equal(jb, oj)} bid_equal={np.array_equal(bd, ob)} ms={ms.value:.4f}",
          flush=True)

# ---- 2. full-solve parity, smallest first
np.random.seed(1); mat = np.random.uniform(0, 10, (5, 5))
g = sslap_b200.auction_solve(mat, problem="min"); print("known answer 1:", g)
np.random.seed(2); mat[np.random.rand(5, 5) > 0.5] = -1
g = sslap_b200.auction_solve(mat=mat, problem="max"); print("known answer 2:", g)
print("known HK:", sslap_b200.hopcroft_solve(lookup={0: [0, 1], 1: [1, 2], 2: [1, 4], 3: [2], 4: [3]}))

cases = [("n20f", 20, 0.3, "float", "min"), ("n20i", 20, 0.3, "int", "max"), ("n64i", 64, 0.2, "int", "min"),
         ("n200f", 200, 0.05, "float", "min"), ("n1000i", 1000, 0.01, "int", "min"), ("n1000f", 1000, 0.01, "float", "max"),
         ("n3000i", 3000, 0.01, "int", "min")]
allok = True
for (name, n, d, md, pb) in cases:
    loc, val = make_problem(n, d, md, seed=1)
    for ts in (32, 4, 0):
        ok = compare(name, loc, val, pb, ts)
        allok &= ok
        if not ok:
            o = oracle.auction_solve(loc=loc, val=val, problem=pb)
            try:
                first_divergence(name, loc, val, pb, ts, o["meta"]["its"] + 1)
            except Exception as e:
                print("   divergence search failed:", e)
print("ALL SMALL OK" if allok else "SOME SMALL FAILED", flush=True)

# ---- 3. Hopcroft-Karp
for (n, d) in [(500, 0.004), (2000, 0.001), (10000, 0.0003)]:
    rng = np.random.default_rng(n)
    k = int(n * n * d)
    key = np.unique(rng.integers(0, n, k).astype(np.int64) * n + rng.integers(0, n, k))
    loc = np.stack([key // n, key % n], -1).astype(np.int32)
    t = time.perf_counter(); g = sslap_b200.hopcroft_solve(loc=loc); tg = time.perf_counter() - t
    o = oracle.hopcroft_solve(loc=loc)
    lp, rp = g["left_pairings"], g["right_pairings"]
    valid = all(rp[lp[u]] == u for u in range(len(lp)) if lp[u] >= 0) and (lp >= 0).sum() == g["size"]
    print(f"[hk n={n}] gpu={g['size']} oracle={o['size']} valid={valid} t={tg*1e3:.1f}ms", flush=True)

if mode == "full":
    loc, val = make_problem(10000, 0.01, "float", seed=0)
    for ts in (32, 0):
        compare("C2", loc, val, "min", ts)
    compare("C2-repeat", loc, val, "min", 32)
    t = time.perf_counter(); loc, val = make_problem(100000, 0.001, "float", seed=0); print("gen C3", time.perf_counter() - t)
    compare("C3", loc, val, "min", 32)
    compare("C3-repeat", loc, val, "min", 32)
    n = 100000
    h.set_option("t_small", 32)
    for flush in (0, 1):
        ms = C.c_float(0)
        rc = L.sslapb_bid_sweep(h.ptr, None, None, n, 0.5, 1, 10, flush, None, None, C.byref(ms))
        by = 12 * val.size + 36 * n
        print(f"[C3 full sweep flush={flush}] rc={rc} {ms.value*1e3:.1f} us  {by/ms.value/1e6:.1f} GB/s", flush=True)
