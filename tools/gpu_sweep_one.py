"""C3, final prices, N launches of one sweep variant (ncu target): python tools/gpu_sweep_one.py [merge] [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
import sslap_b200
from sslap_b200 import _native as nat
from sslap_b200.datagen import make_problem
h = nat.default_handle(); L = nat.load()
merge = int(sys.argv[1]) if len(sys.argv) > 1 else 0
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n = 100000
loc, val = make_problem(n, 0.001, "float", seed=0)
sslap_b200.auction_solve(loc=loc, val=val, size=(n, n), cardinality_check=False)
ms = C.c_float(0)
rc = L.sslapb_bid_sweep(h.ptr, None, None, n, 1e-5, merge, iters, 1, None, None, C.byref(ms))
assert rc == 0
print(f"merge={merge}: {ms.value*1e3:.1f} us  {(12*val.size+44*n)/ms.value/1e6:.0f} GB/s")
