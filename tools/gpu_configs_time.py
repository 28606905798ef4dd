"""Timings of BASELINE.json's configs through the drop-in call (best of 3, host arrays; C4 once): python tools/gpu_configs_time.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sslap_b200
from sslap_b200 import _native as nat
from sslap_b200.datagen import make_problem, objective
h = nat.default_handle()
for tag, n, d, mode, hk, reps in [("C1", 1000, 0.01, "int", True, 5), ("C2", 10000, 0.01, "float", True, 3), ("C3", 100000, 0.001, "float", False, 3),
                                  ("C3+HK", 100000, 0.001, "float", True, 2), ("C4", 1000000, 1e-4, "float", True, 2)]:
    if tag == "C4" and os.environ.get("SKIP_C4"):
        continue
    loc, val = make_problem(n, d, mode, seed=0)
    for hot in ((1, 0) if tag in ("C1", "C2", "C4") else (1,)):
        h.set_option("hot", hot)
        best, m = 1e9, None
        for _ in range(reps):
            t = time.perf_counter()
            r = sslap_b200.auction_solve(loc=loc, val=val, size=(n, n), problem="min", cardinality_check=hk, max_iter=50000000, _raw_meta=True)
            dt = time.perf_counter() - t
            if dt < best: best, m = dt, r["raw"]
        print(f"{tag:6s} hot={hot} wall {best*1e3:9.2f} ms  solve {m.solve_ms:9.2f} hk {m.hk_ms:7.2f} setup {m.setup_ms:6.2f} h2d {m.h2d_ms:7.2f}  its {m.its} "
              f"rounds g/m/w/s {m.rounds_grid}/{m.rounds_mid}/{m.rounds_warp}/{m.rounds_solo} nohole {m.rounds_nohole} hot grid {m.hot_grid_bids}/{m.hot_grid_fallbacks} tail {m.hot_tail_rounds}/{m.hot_tail_fallbacks} "
              f"obj {objective(loc, val, r['sol']):.6f}", flush=True)
        print("        sections ms: " + " ".join(f"{x:.1f}" for x in m.prof_ms), flush=True)
    h.set_option("hot", 1)
