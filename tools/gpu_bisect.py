"""Bisect helper: sparse-ish dense-input problems (the reference's density sweep shape) against the oracle, for the library
named by SSLAP_B200_LIB, with the hot lists on and off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import sslap_b200
from sslap_b200 import _native as nat
from oracle import oracle
from benchmarking import make_matrix
h = nat.default_handle()
print("lib", nat.LIB_PATH)
for hot in (1, 0):
    h.set_option("hot", hot)
    for (n, d) in [(500, 0.01), (500, 0.05), (300, 0.02), (1000, 0.01), (200, 0.03), (500, 0.2)]:
        mat = make_matrix(n, d, "float")
        want = oracle.auction_solve(mat=mat, problem="max")
        try:
            got = sslap_b200.auction_solve(mat, problem="max", _raw_meta=True)
            ok = np.array_equal(got["sol"], want["sol"]) and got["meta"]["its"] == want["meta"]["its"]
            m = got["raw"]
            print(f"hot={hot} n={n} d={d}: {'OK ' if ok else 'BAD'} its {got['meta']['its']} vs {want['meta']['its']} unique {np.unique(got['sol']).size} "
                  f"rounds g/w/s {m.rounds_grid}/{m.rounds_warp}/{m.rounds_solo} hot tail {m.hot_tail_rounds} fell {m.hot_tail_fallbacks} grid {m.hot_grid_bids}/{m.hot_grid_fallbacks}", flush=True)
        except Exception as e:
            print(f"hot={hot} n={n} d={d}: EXC {e!r}", flush=True)
