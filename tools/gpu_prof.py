"""Per-regime device-time profile of the persistent kernel on C2 / C3 (python tools/gpu_prof.py [c2|c3|both])."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sslap_b200
from sslap_b200 import _native as nat
from sslap_b200.datagen import make_problem
import ctypes as C
h = nat.default_handle(); L = nat.load()
which = sys.argv[1] if len(sys.argv) > 1 else "both"
names = ["grid_bid", "grid_assign", "grid_compact", "warp", "solo", "ece_phase", "mid", "grid_barriers"]
_problems = {}
def run(tag, n, d, reps=2, **kw):
    if (n, d) not in _problems:
        f = f"/tmp/sslap_prof_{n}_{d}.npz"                   # several library builds are profiled back to back (gpu_mid_ab.sh)
        if os.path.exists(f):
            z = np.load(f); _problems[(n, d)] = (z["loc"], z["val"])
        else:
            _problems[(n, d)] = make_problem(n, d, "float", seed=0)
            np.savez(f, loc=_problems[(n, d)][0], val=_problems[(n, d)][1])
    loc, val = _problems[(n, d)]
    for r in range(reps):
        t = time.perf_counter()
        g = sslap_b200.auction_solve(loc=loc, val=val, size=(n, n), cardinality_check=False, _raw_meta=True, **kw)
        w = time.perf_counter() - t
        m = g["raw"]
        mid = int(getattr(m, "rounds_mid", m.its - m.rounds_grid - m.rounds_warp - m.rounds_solo))
        print(f"[{tag}] its={m.its} rounds g/w/s={m.rounds_grid}/{m.rounds_warp}/{m.rounds_solo} solve={m.solve_ms:.2f}ms "
              f"setup={m.setup_ms:.2f} h2d={m.h2d_ms:.2f} wall={w*1e3:.1f}ms 2nd_pass_rows={m.prune_second_pass}", flush=True)
        print("    " + "  ".join(f"{k}={v:.2f}ms" for k, v in zip(names, m.prof_ms)))
        print(f"    hot lists: grid bids {m.hot_grid_bids} (+{m.hot_grid_fallbacks} handed to the full row)  tail rounds {m.hot_tail_rounds} (bids handed on: {m.hot_tail_fallbacks})  no-hole grid rounds {m.rounds_nohole}")
        per = lambda t, c: (1e3 * t / c) if c else 0.0
        print(f"    per-round us: grid={per(m.prof_ms[0]+m.prof_ms[1]+m.prof_ms[2]+m.prof_ms[7], m.rounds_grid):.2f} "
              f"(barriers {per(m.prof_ms[7], m.rounds_grid):.2f}) warp={per(m.prof_ms[3], m.rounds_warp):.2f} solo={per(m.prof_ms[4], m.rounds_solo):.2f} "
              f"mid={per(m.prof_ms[6], mid):.2f} ({mid} rounds)  sol crc {int(np.bitwise_xor.reduce(g['sol'].astype(np.int64) * np.arange(1, n + 1)))}", flush=True)
    return loc, val
if which in ("c2", "both"):
    run("C2", 10000, 0.01)
    h.set_option("hot", 0); run("C2 hot off", 10000, 0.01, reps=1); h.set_option("hot", 1)
if which == "midsweep":                                      # the mid regime's threshold (0 = off) on C2 and C3
    for t_mid in (0, 64, 128, 256):
        try:
            h.set_option("t_mid", t_mid)
        except Exception as e:
            if "unknown option" in str(e):
                print("t_mid not settable:", e); run("C3", 100000, 0.001, reps=2); break
            continue                                         # a build with a smaller SSLAPB_MID
        run(f"C3 t_mid={t_mid}", 100000, 0.001, reps=2)
        if os.environ.get("PROF_C2"): run(f"C2 t_mid={t_mid}", 10000, 0.01, reps=2)
if which == "c3mid0":                                        # C3 with the mid regime off (code-placement A/B of the other loops)
    try:
        h.set_option("t_mid", 0)
    except Exception:
        pass
    run("C3 t_mid=0", 100000, 0.001, reps=2)
if which == "c3only":
    run("C3", 100000, 0.001, reps=2)
if which in ("c3", "both"):
    loc, val = run("C3", 100000, 0.001)
    h.set_option("hot", 0); run("C3 hot off", 100000, 0.001, reps=1); h.set_option("hot", 1)
    h.set_option("l2_persist", 0); run("C3 hot on, no L2 window", 100000, 0.001, reps=2); h.set_option("l2_persist", 1)
    n = 100000
    for merge in (1, 3):
        for flush in (0, 1):
            ms = C.c_float(0)
            rc = L.sslapb_bid_sweep(h.ptr, None, None, n, 1e-5, merge, 20, flush, None, None, C.byref(ms))
            by = 12 * val.size + 36 * n
            print(f"[C3 full sweep prune={'off' if merge & 2 else 'on'} flush={flush}] rc={rc} {ms.value*1e3:.1f} us  {by/ms.value/1e6:.1f} GB/s  frac={by/ms.value/1e6/6544:.3f}", flush=True)
