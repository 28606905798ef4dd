"""Dense N = 3162 (the reference's published benchmark shape; long-row kernel instance) with the mid regime off / on:
python tools/gpu_dense_mid.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sslap_b200
from sslap_b200 import _native as nat
h = nat.default_handle()
for mode in ("float", "int"):
    np.random.seed(1)
    mat = np.random.uniform(0, 100, (3162, 3162)) if mode == "float" else np.random.randint(1, 100, (3162, 3162)).astype(np.float64)
    ref = None
    for t_mid in (0, 128, 0, 128):
        h.set_option("t_mid", t_mid)
        t = time.perf_counter()
        r = sslap_b200.auction_solve(mat=mat, problem="max", _raw_meta=True)
        w = time.perf_counter() - t
        m = r["raw"]
        ref = r["sol"] if ref is None else ref
        print(f"dense 3162 {mode:5s} t_mid={t_mid:3d} wall {w*1e3:7.1f} ms solve {m.solve_ms:7.2f} its {m.its} rounds g/m/w/s {m.rounds_grid}/{m.rounds_mid}/{m.rounds_warp}/{m.rounds_solo} "
              f"mid {m.prof_ms[6]:.2f} ms same_sol {bool(np.array_equal(ref, r['sol']))}", flush=True)
h.set_option("t_mid", 128)
