"""C5 throughput: 4096 independent 512x512 problems at 5 % — the sub-warp batch kernel (default) against round 1's
(option batch_v1), device time and wall time, results checked against the oracle.  python tools/gpu_batch.py [P]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sslap_b200
from sslap_b200 import _native as nat
from sslap_b200.datagen import make_problem
from oracle import oracle
h = nat.default_handle()
P = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
t = time.perf_counter()
base = [make_problem(512, 0.05, "float", seed=s) for s in range(64)]
probs = [(base[k % 64][0], base[k % 64][1], (512, 512)) for k in range(P)]
print(f"generated {P} problems ({time.perf_counter()-t:.1f}s), nnz/problem ~{len(base[0][1])}", flush=True)
packed = sslap_b200.pack_problems(probs)
for v1 in (0, 1, 0):
    h.set_option("batch_v1", v1)
    for rep in range(2):
        t = time.perf_counter()
        r = sslap_b200.auction_solve_batch(packed, packed_result=True)
        dt = time.perf_counter() - t
        print(f"{'round-1 kernel' if v1 else 'sub-warp kernel'} rep {rep}: wall {dt*1e3:.1f} ms  kernel {r['metas'][0].solve_ms:.2f} ms  "
              f"h2d {r['metas'][0].h2d_ms:.1f} ms  {P/dt:.0f} problems/s", flush=True)
h.set_option("batch_v1", 0)
res = sslap_b200.auction_solve_batch(probs[:64])
ok = True
t = time.perf_counter()
for k in range(64):
    o = oracle.auction_solve(loc=base[k][0], val=base[k][1])
    ok &= np.array_equal(o["sol"], res[k]["sol"]) and o["meta"]["its"] == res[k]["meta"]["its"]
dt = time.perf_counter() - t
print(f"64 distinct problems identical to the oracle: {ok}; oracle (C port, 1 core): {dt/64*1e3:.2f} ms/problem -> {dt/64*P:.1f} s for the batch")
