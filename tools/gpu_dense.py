import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, sslap_b200
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from benchmarking import make_matrix
names = ["grid_bid", "grid_assign", "grid_compact", "warp", "solo", "ece_phase", "-", "grid_barriers"]
for n in (1000, 3162):
    mat = make_matrix(n, 1.0, 'float')
    for rep in range(2):
        t = time.perf_counter(); r = sslap_b200.auction_solve(mat, problem='max', _raw_meta=True); w = time.perf_counter() - t
        m = r["raw"]
    print(f"N={n} dense: wall {w*1e3:.1f} ms solve {m.solve_ms:.1f} setup {m.setup_ms:.2f} h2d {m.h2d_ms:.2f} hk {m.hk_ms:.2f} its={m.its} rounds g/w/s={m.rounds_grid}/{m.rounds_warp}/{m.rounds_solo}")
    print("    " + "  ".join(f"{k}={v:.2f}ms" for k, v in zip(names, m.prof_ms)))
    per = lambda t, c: (1e3 * t / c) if c else 0.0
    print(f"    per-round us: grid={per(m.prof_ms[0]+m.prof_ms[1]+m.prof_ms[2]+m.prof_ms[7], m.rounds_grid):.2f} warp={per(m.prof_ms[3], m.rounds_warp):.2f} solo={per(m.prof_ms[4], m.rounds_solo):.2f}", flush=True)
