import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, sslap_b200
from sslap_b200.datagen import make_problem
for (n, d) in [(10000, 0.01), (100000, 0.001)]:
    loc, val = make_problem(n, d, "float", seed=0)
    for rep in range(3):
        t = time.perf_counter()
        r = sslap_b200.auction_solve(loc=loc, val=val, size=(n, n), cardinality_check=True, _raw_meta=True)
        w = time.perf_counter() - t
        m = r["raw"]
        print(f"n={n}: wall {w*1e3:.1f} ms  hk {m.hk_ms:.2f} ms  solve {m.solve_ms:.1f} ms  h2d {m.h2d_ms:.1f} setup {m.setup_ms:.2f}", flush=True)
