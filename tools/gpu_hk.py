"""Hopcroft-Karp timings: device-resident phase loop (default) against round 1's host-driven loop (option hk_host_loop).
python tools/gpu_hk.py [c4]   -> wall ms per call (pinned-free, pageable inputs), cardinality checked against scipy"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import maximum_bipartite_matching
import sslap_b200
from sslap_b200 import _native as nat
from sslap_b200.datagen import make_problem
h = nat.default_handle()
rng = np.random.default_rng(9)


def graph(n, m, e):
    key = np.unique(rng.integers(0, n, e).astype(np.int64) * m + rng.integers(0, m, e))
    return np.stack([key // m, key % m], -1).astype(np.int32)


cases = [("deficient 100k x 100k, 300k edges", graph(100000, 100000, 300000)),
         ("deficient 1M x 1M, 2M edges", graph(1000000, 1000000, 2000000)),
         ("C3 graph (100k, 10.1M edges)", make_problem(100000, 0.001, "float", seed=0)[0])]
if len(sys.argv) > 1 and sys.argv[1] == "c4":
    cases.append(("C4 graph (1M, 101M edges)", make_problem(1000000, 1e-4, "float", seed=0)[0]))
for name, loc in cases:
    n, m = int(loc[:, 0].max()) + 1, int(loc[:, 1].max()) + 1
    g = sp.csr_matrix((np.ones(len(loc), dtype=np.int8), (loc[:, 0], loc[:, 1])), shape=(n, m))
    t = time.perf_counter(); card = int((maximum_bipartite_matching(g, perm_type="column") >= 0).sum()); t_sc = time.perf_counter() - t
    for mode in (0, 1):
        h.set_option("hk_host_loop", mode)
        best = 1e9
        for rep in range(3):
            t = time.perf_counter(); r = sslap_b200.hopcroft_solve(loc=loc); best = min(best, time.perf_counter() - t)
        assert r["size"] == card, (name, r["size"], card)
        print(f"{name:36s} {'host loop (r1)' if mode else 'device loop  '}: {best*1e3:8.2f} ms wall  card={r['size']}  (scipy {t_sc*1e3:.0f} ms)", flush=True)
h.set_option("hk_host_loop", 0)
