import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
import sslap_b200
from sslap_b200 import _native as nat
from sslap_b200.datagen import make_problem
from oracle import oracle
h = nat.default_handle(); L = nat.load()
def prices(m):
    p = np.empty(m); L.sslapb_get_prices(h.ptr, p.ctypes.data); return p
rng = np.random.default_rng(2024)
bad = 0
for case in range(120):
    n = int(rng.integers(2, 400))
    m = n + int(rng.integers(0, 30)) if rng.random() < 0.3 else n
    density = float(rng.choice([0.02, 0.05, 0.1, 0.3, 0.7, 1.0]))
    mode = "int" if rng.random() < 0.5 else "float"
    loc, val = make_problem(n, density, mode, seed=1000 + case, m=m)
    if rng.random() < 0.2:
        val = np.round(val / 10.0)
    problem = "min" if rng.random() < 0.5 else "max"
    kw = {}
    r = rng.random()
    if r < 0.15: kw["eps_start"] = float(rng.choice([0.5, 3.0, 40.0]))
    elif r < 0.25: kw["max_iter"] = int(rng.integers(1, 200))
    t_small = int(rng.choice([32, 32, 16, 4, 1, 0]))
    def run(ts, mi=None):
        k2 = dict(kw)
        if mi is not None: k2["max_iter"] = mi
        h.set_option("t_small", ts)
        g = sslap_b200.auction_solve(loc=loc, val=val, size=(n, m), problem=problem, cardinality_check=False, _raw_meta=True, **k2)
        o = oracle.auction_solve(loc=loc, val=val, problem=problem, return_prices=True, **k2)
        return g, o, np.array_equal(g["sol"], o["sol"]) and np.array_equal(prices(m), o["prices"])
    g, o, ok = run(t_small)
    if not ok:
        bad += 1
        print(f"case {case}: n={n} m={m} dens={density} {mode} {problem} kw={kw} t_small={t_small} its gpu/oracle={g['meta']['its']}/{o['meta']['its']} deg-1 rows={(np.bincount(loc[:,0], minlength=n)==1).sum()}")
        for ts in (32, 16, 4, 1, 0):
            print("   t_small", ts, "ok" if run(ts)[2] else "MISMATCH")
        lo, hi = 0, min(o["meta"]["its"], g["meta"]["its"]) + 1
        while lo + 1 < hi:
            mid = (lo + hi) // 2
            if run(t_small, mid)[2]: lo = mid
            else: hi = mid
        g, o, _ = run(t_small, hi)
        pg = prices(m)
        d = np.nonzero(pg != o["prices"])[0][:6]
        print(f"   first bad round {hi}: rounds g/w/s={g['raw'].rounds_grid}/{g['raw'].rounds_warp}/{g['raw'].rounds_solo} price diffs at {d} gpu={pg[d]} oracle={o['prices'][d]}  sol diffs {np.nonzero(g['sol'] != o['sol'])[0][:6]}")
        dd = np.nonzero(g['sol'] != o['sol'])[0]
        print("   persons", dd, "gpu sol", g['sol'][dd], "oracle sol", o['sol'][dd])
        for i_ in dd:
            js = loc[loc[:, 0] == i_, 1]; vs = val[loc[:, 0] == i_]
            print("     person", i_, "cols", js, "vals", vs, "prices", o['prices'][js])
        gp, op, _ = run(t_small, hi - 1)
        print("   previous round sol equal:", np.array_equal(gp['sol'], op['sol']), " unassigned before:", np.nonzero(op['sol'] < 0)[0][:20])
        if bad >= 3: break
print("bad cases:", bad)
