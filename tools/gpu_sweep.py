"""Stand-alone sweep experiments on C3 with final prices: python tools/gpu_sweep.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
import sslap_b200
from sslap_b200 import _native as nat
from sslap_b200.datagen import make_problem
h = nat.default_handle(); L = nat.load()
n = 100000
loc, val = make_problem(n, 0.001, "float", seed=0)
sslap_b200.auction_solve(loc=loc, val=val, size=(n, n), cardinality_check=False)
by = lambda nb: 12 * val.size * nb / n + 36 * nb
for (nb, merge, flush) in [(n, 1, 1), (n, 0, 1), (n, 3, 1), (n // 2, 1, 1), (n // 4, 1, 1), (n, 1, 0)]:
    ms = C.c_float(0)
    rc = L.sslapb_bid_sweep(h.ptr, None, None, nb, 1e-5, merge, 10, flush, None, None, C.byref(ms))
    print(f"nb={nb} merge={merge} flush={flush}: {ms.value*1e3:.1f} us  {by(nb)/ms.value/1e6:.0f} GB/s", flush=True)
