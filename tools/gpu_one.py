"""One solve of a named config (for ncu): python tools/gpu_one.py c2|c1|c3"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sslap_b200
from sslap_b200.datagen import make_problem
cfg = {"c1": (1000, 0.01, "int"), "c2": (10000, 0.01, "float"), "c3": (100000, 0.001, "float")}[sys.argv[1]]
loc, val = make_problem(cfg[0], cfg[1], cfg[2], seed=0)
for _ in range(2):
    r = sslap_b200.auction_solve(loc=loc, val=val, size=(cfg[0], cfg[0]), cardinality_check=False)
print(r["meta"])
