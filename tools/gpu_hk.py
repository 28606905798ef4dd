"""Hopcroft-Karp timing/variance on several graphs: python tools/gpu_hk.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sslap_b200
from sslap_b200.datagen import make_problem
from oracle import oracle
def rnd_graph(n, e, seed):
    rng = np.random.default_rng(seed)
    key = np.unique(rng.integers(0, n, e).astype(np.int64) * n + rng.integers(0, n, e))
    return np.stack([key // n, key % n], -1).astype(np.int32)
cases = [("C2 graph", make_problem(10000, 0.01, "float", 0)[0]), ("N=100k deg3", rnd_graph(100000, 300000, 1)),
         ("N=100k deg1.2", rnd_graph(100000, 120000, 2)), ("C3 graph", make_problem(100000, 0.001, "float", 0)[0]),
         ("N=1M deg2", rnd_graph(1000000, 2000000, 3))]
for name, loc in cases:
    ts = []
    for rep in range(4):
        t = time.perf_counter(); r = sslap_b200.hopcroft_solve(loc=loc); ts.append(time.perf_counter() - t)
    t = time.perf_counter(); o = oracle.hopcroft_solve(loc=loc); to = time.perf_counter() - t
    print(f"{name:16s} edges={len(loc):9d} gpu size={r['size']} oracle size={o['size']} gpu times(s)={[round(x,3) for x in ts]} oracle C {to:.3f}s", flush=True)
