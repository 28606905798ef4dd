"""BASELINE.json configs[3]: 1M x 1M at 0.01 % (~101 M nnz), HK check on.  python tools/gpu_c4.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sslap_b200
from sslap_b200.datagen import make_problem, objective
n = 1000000
t = time.perf_counter()
loc, val = make_problem(n, 0.0001, "float", seed=0)
print(f"generated nnz={len(val)} in {time.perf_counter()-t:.1f}s", flush=True)
t = time.perf_counter()
hk = sslap_b200.hopcroft_solve(loc=loc)
print(f"hopcroft_solve: size={hk['size']} in {time.perf_counter()-t:.2f}s", flush=True)
for cc in (True, False):
    t = time.perf_counter()
    r = sslap_b200.auction_solve(loc=loc, val=val, size=(n, n), problem="min", cardinality_check=cc, max_iter=50000000, _raw_meta=True)
    dt = time.perf_counter() - t
    m = r["raw"]
    print(f"auction_solve(cardinality_check={cc}): wall {dt:.2f}s solve {m.solve_ms/1e3:.2f}s hk {m.hk_ms/1e3:.2f}s h2d {m.h2d_ms:.0f}ms "
          f"its={m.its} rounds g/w/s={m.rounds_grid}/{m.rounds_warp}/{m.rounds_solo} meta={ {k: v for k, v in r['meta'].items() if k != 'timer'} }", flush=True)
sol = r["sol"]
print("perfect matching:", bool(np.array_equal(np.sort(sol), np.arange(n))), "objective", objective(loc, val, sol), "(reference: 1630435.704360, SURVEY 6.2)")
