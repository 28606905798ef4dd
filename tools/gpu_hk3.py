import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, sslap_b200
from sslap_b200.datagen import make_problem
loc, val = make_problem(100000, 0.001, "float", seed=0)
for rep in range(2):
    t = time.perf_counter(); r = sslap_b200.hopcroft_solve(loc=loc); print("hk", r["size"], time.perf_counter() - t, flush=True)
