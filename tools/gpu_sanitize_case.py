"""A few small solves for compute-sanitizer runs: compute-sanitizer --tool memcheck|racecheck python tools/gpu_sanitize_case.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sslap_b200
from sslap_b200 import _native as nat
from sslap_b200.datagen import make_problem
from oracle import oracle
h = nat.default_handle()
h.set_option("watchdog_ms", 100000)
ok = True
for (n, d, mode, seed, ts) in [(120, 0.1, "float", 1, 32), (200, 0.05, "int", 2, 32), (90, 0.3, "float", 3, 4), (300, 0.5, "float", 4, 32)]:
    loc, val = make_problem(n, d, mode, seed=seed)
    h.set_option("t_small", ts)
    got = sslap_b200.auction_solve(loc=loc, val=val, size=(n, n), problem="max", cardinality_check=True, max_iter=20000)
    want = oracle.auction_solve(loc=loc, val=val, problem="max", max_iter=20000)
    good = np.array_equal(got["sol"], want["sol"]) and got["meta"]["its"] == want["meta"]["its"]
    ok &= good
    print(n, d, mode, ts, "ok" if good else "MISMATCH", got["meta"]["its"], flush=True)
print("SANITIZE CASES OK" if ok else "SANITIZE CASES FAILED", flush=True)
