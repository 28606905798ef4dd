"""Where does the time of a SMALL dense problem go (the regime of the reference's published figures, benchmarking.py)?
python tools/gpu_small.py  -> per size: wall of the drop-in call (best of 30), device timers, and the same with the
persistent kernel's grid capped (option max_ctas)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sslap_b200
from sslap_b200 import _native as nat
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from benchmarking import make_matrix
h = nat.default_handle()
ref = None
try:
    from oracle import ref_loader
    if ref_loader.available():
        ref = ref_loader.load()
except Exception:
    pass
for n in (10, 20, 32, 48, 64, 100, 128, 200, 316):
    mat = make_matrix(n, 1.0, "float")
    tr = float("nan")
    if ref is not None:
        tr = min(_t for _t in [(lambda t0: (ref.auction_solve(mat.copy(), problem="max"), time.perf_counter() - t0)[1])(time.perf_counter()) for _ in range(10)])
    h.set_option("small_max_n", 256)
    for small in (1, 0):
        h.set_option("small_path", small)
        for cc in (True, False):
            best, m = 1e9, None
            for _ in range(30):
                t = time.perf_counter()
                r = sslap_b200.auction_solve(mat, problem="max", cardinality_check=cc, _raw_meta=True)
                dt = time.perf_counter() - t
                if dt < best:
                    best, m = dt, r["raw"]
            print(f"N={n:4d} small_path={small} (taken {m.small_path}) hk={int(cc)} wall {best*1e3:7.3f} ms (reference {tr*1e3:6.3f})  solve {m.solve_ms:6.3f} setup {m.setup_ms:6.3f} "
                  f"hk {m.hk_ms:6.3f} h2d {m.h2d_ms:6.3f}  its {m.its}", flush=True)
    h.set_option("small_path", 1); h.set_option("small_max_n", 128)
