"""Full-frontier bid sweep at C4 size (1M x 1M, ~101 M nnz, final prices): python tools/gpu_sweep_c4.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
import sslap_b200
from sslap_b200 import _native as nat
from sslap_b200.datagen import make_problem
h = nat.default_handle(); L = nat.load()
n = 1000000
loc, val = make_problem(n, 0.0001, "float", seed=0)
r = sslap_b200.auction_solve(loc=loc, val=val, size=(n, n), cardinality_check=False, max_iter=50000000)
print("solved, its", r["meta"]["its"], flush=True)
by = 12 * val.size + 44 * n
for merge in (1, 0, 3):
    ms = C.c_float(0)
    rc = L.sslapb_bid_sweep(h.ptr, None, None, n, 1e-6, merge, 10, 1, None, None, C.byref(ms))
    assert rc == 0
    print(f"C4 full sweep merge={merge}: {ms.value*1e3:.1f} us  {by/ms.value/1e6:.0f} GB/s  frac={by/ms.value/1e6/6544:.3f}", flush=True)
