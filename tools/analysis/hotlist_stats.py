"""Runs tools/analysis/hotlist_stats.c (the oracle with a per-bid statistics hook) on a BASELINE config.
usage: python tools/analysis/hotlist_stats.py [n] [density] [mode]"""
import ctypes as C, os, subprocess, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from sslap_b200.datagen import make_problem
from oracle.oracle import OracleMeta
so = os.path.join(HERE, "_hotlist_stats.so")
subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-o", so, os.path.join(HERE, "hotlist_stats.c"), "-lm"])
L = C.CDLL(so)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
d = float(sys.argv[2]) if len(sys.argv) > 2 else 0.001
mode = sys.argv[3] if len(sys.argv) > 3 else "float"
loc, val = make_problem(n, d, mode, seed=0)
rows = np.ascontiguousarray(loc[:, 0], dtype=np.int32); cols = np.ascontiguousarray(loc[:, 1], dtype=np.int32)
sol = np.empty(n, dtype=np.int32); meta = OracleMeta()
L.sslap_oracle_auction.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int, C.c_float, C.c_int64, C.c_int,
                                   C.c_void_p, C.POINTER(OracleMeta), C.c_void_p, C.c_void_p]
L.sslap_oracle_auction(rows.ctypes.data, cols.ctypes.data, val.ctypes.data, val.size, n, n, 0, 0.0, 50000000, 0, sol.ctypes.data, C.byref(meta), None, None)
print("its", meta.its, "nnz", val.size)
L.stats_report()
