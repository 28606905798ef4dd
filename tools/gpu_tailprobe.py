"""Debug build only (-DSSLAPB_TAIL_PROBE, abtest/lib_tailprobe.so): where do the cycles of a chain round go?
SSLAP_B200_LIB=abtest/lib_tailprobe.so python tools/gpu_tailprobe.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sslap_b200
from sslap_b200 import _native as nat
from sslap_b200.datagen import make_problem
L = nat.load()
n = 100000
loc, val = make_problem(n, 0.001, "float", seed=0)
out = (C.c_longlong * 8)()
for rep in range(2):
    g = sslap_b200.auction_solve(loc=loc, val=val, size=(n, n), cardinality_check=False, _raw_meta=True)
    L.sslapb_tail_probe_read(out)
    m = g["raw"]
    k = max(1, out[4])
    print(f"solve {m.solve_ms:.1f} ms, chain rounds probed {out[4]} (of {m.rounds_solo}), chain {1e3 * m.prof_ms[4] / m.rounds_solo:.3f} us/round")
    print(f"  cycles per round: wait for the bidder's hot row {out[0]/k:.0f} | hot row -> winner known (record gather + first reduction) {out[1]/k:.0f} | "
          f"-> bid known (second-best reductions) {out[2]/k:.0f} | -> commit issued {out[3]/k:.0f} | sum {sum(out[:4])/k:.0f}")
