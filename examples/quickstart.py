"""sslap_b200 in five calls (needs a B200: there is no CPU fallback).  python examples/quickstart.py

The calls and the returned dictionaries are those of the reference package (`from sslap import auction_solve,
hopcroft_solve`); `auction_solve_batch` is the one addition."""
import os
import sys

import numpy as np
from scipy.sparse import coo_matrix

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sslap_b200 import auction_solve, auction_solve_batch, hopcroft_solve  # noqa: E402

rng = np.random.default_rng(0)

# 1. dense matrix: entries >= 0 are valid, negative entries mean "no edge"
cost = rng.uniform(0, 10, (6, 6))
cost[rng.random((6, 6)) < 0.3] = -1
cost[np.arange(6), rng.permutation(6)] = rng.uniform(0, 10, 6)          # keep it feasible
res = auction_solve(mat=cost, problem="min")
print("dense     sol", res["sol"], "obj", res["meta"]["obj"], "its", res["meta"]["its"])

# 2. the same problem as (row, column) pairs + values: the super-sparse form, rows sorted
r, c = np.nonzero(cost >= 0)
loc = np.stack([r, c], axis=-1).astype(np.int32)
val = cost[r, c]
res2 = auction_solve(loc=loc, val=val, problem="min", size=cost.shape)
assert np.array_equal(res["sol"], res2["sol"])

# 3. scipy COO matrix (zeros are "no edge" there)
res3 = auction_solve(coo_mat=coo_matrix((val + 1.0, (r, c)), shape=cost.shape), problem="min")
print("coo       sol", res3["sol"])

# 4. feasibility only: maximum-cardinality matching (Hopcroft-Karp)
hk = hopcroft_solve(loc=loc)
print("hopcroft  size", hk["size"], "left", hk["left_pairings"])

# 5. many independent small problems in one call (one warp per problem on the device)
problems = []
for k in range(64):
    m = rng.uniform(0, 100, (32, 32))
    rr, cc = np.nonzero(m > 40)
    keep = np.stack([rr, cc], axis=-1).astype(np.int32)
    diag = np.stack([np.arange(32), rng.permutation(32)], axis=-1).astype(np.int32)     # planted perfect matching
    allp = np.unique(np.concatenate([keep, diag]), axis=0)
    problems.append((allp, m[allp[:, 0], allp[:, 1]], (32, 32)))
batch = auction_solve_batch(problems, problem="max")
print("batch     ", len(batch), "problems, first sol", batch[0]["sol"][:8], "...")
