"""A/B of the full-frontier bidding sweep on C3 (final prices, L2 flushed before every launch): the software-pipelined kernel
(default) in its CTA sizes, with the trimmed / untrimmed stage C, against the round-1 per-row kernel and the TMA ring.
python tools/gpu_sweep.py [iters]     -> one line per variant: us per launch, GB/s, fraction of the measured copy peak"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
import sslap_b200
from sslap_b200 import _native as nat
from sslap_b200.datagen import make_problem
h = nat.default_handle(); L = nat.load()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
peak = 6455.6
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
n = 100000
loc, val = make_problem(n, 0.001, "float", seed=0)
sslap_b200.auction_solve(loc=loc, val=val, size=(n, n), cardinality_check=False)
by = 12 * val.size + 36 * n
variants = [("4 rows/warp, 8 lanes/row", 1 | 128), ("4 rows/warp, pruning off", 3 | 128),
            ("per-row kernel (default)", 1), ("per-row kernel, pruning off", 3),
            ("pipelined per-row 768 thr", 1 | 8), ("pipelined per-row 1024 thr", 1 | 8 | 16), ("TMA ring", 1 | 4),
            ("hot form (hot lists + exact fallback)", 1 | 256),
            ("lean full-row sweep + redo list", 1 | 512), ("lean full-row sweep, pruning off", 3 | 512)]
for rep in range(2):
    for name, merge in variants:
        ms = C.c_float(0)
        out = []
        for fl in (1, 2):                                   # 1: L2 flushed by a 256 MB memset; 2: memset + 256 MB read (clean lines)
            rc = L.sslapb_bid_sweep(h.ptr, None, None, n, 1e-5, merge, iters, fl, None, None, C.byref(ms))
            out.append(f"flush{fl}: {ms.value*1e3:6.1f} us {by/ms.value/1e6:6.0f} GB/s frac={by/ms.value/1e6/peak:.3f}")
        print(f"rep{rep} {name:42s} rc={rc}  " + "   ".join(out), flush=True)
