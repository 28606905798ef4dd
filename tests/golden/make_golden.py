"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref, built by oracle/build_ref.sh from
/root/reference) on seeded inputs.  Run in the build container only:  python tests/golden/make_golden.py

The reference ships no tests of its own; these fixtures (plus the three known answers printed by its examples/) are
what pins the oracle and the CUDA path.  Inputs are stored next to the outputs so nothing has to be regenerated on
the GPU box (which has no /root/reference).
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader                      # noqa: E402
from sslap_b200.datagen import make_problem        # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ref = ref_loader.load()
META_KEYS = ("start_eps", "eCE", "its", "nreductions", "soln_found", "n_assigned", "obj", "final_eps")


def meta_arr(meta):
    return np.array([float(meta[k]) for k in META_KEYS], dtype=np.float64)


def save(name, **kw):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **kw)
    print("wrote", name, {k: getattr(v, "shape", v) for k, v in kw.items()})


def sparse_case(name, n, density, mode, problem, seed, store_inputs=True, **kw):
    loc, val = make_problem(n, density, mode, seed=seed)
    r = ref.auction_solve(loc=loc, val=val.copy(), size=(n, n), problem=problem, cardinality_check=False, **kw)
    extra = dict(loc=loc, val=val) if store_inputs else dict(
        loc_crc=np.array([zlib.crc32(loc.tobytes())], dtype=np.int64), val_crc=np.array([zlib.crc32(val.tobytes())], dtype=np.int64))
    save(name, n=n, density=density, mode=mode, problem=problem, seed=seed, sol=r["sol"], meta=meta_arr(r["meta"]),
         kwargs=repr(kw), **extra)


# --- the reference's own example scripts (examples/test_auction.py:7-46, examples/test_feasibility.py:16-32)
np.random.seed(1)
mat = np.random.uniform(0, 10, (5, 5)).astype(np.float64)
r = ref.auction_solve(mat.copy(), problem="min")
save("example_dense", mat=mat, problem="min", sol=r["sol"], meta=meta_arr(r["meta"]))
np.random.seed(2)
mat2 = mat.copy()
mat2[np.random.rand(5, 5) > 0.5] = -1
r = ref.auction_solve(mat=mat2.copy(), problem="max")
save("example_sparse", mat=mat2, problem="max", sol=r["sol"], meta=meta_arr(r["meta"]))
lookup = {0: [0, 1], 1: [1, 2], 2: [1, 4], 3: [2], 4: [3]}
r = ref.hopcroft_solve(lookup=lookup)
save("example_hopcroft", loc=np.array([(i, j) for i in lookup for j in lookup[i]], dtype=np.int32), size=r["size"],
     left=r["left_pairings"], right=r["right_pairings"])

# --- BASELINE.json configs[0] (C1) and relatives, small enough to store the inputs
sparse_case("c1_int_min", 1000, 0.01, "int", "min", 0)
sparse_case("c1_float_max", 1000, 0.01, "float", "max", 0)
sparse_case("n300_int_max", 300, 0.05, "int", "max", 7)
sparse_case("n64_int_min_ties", 64, 0.25, "int", "min", 3)
sparse_case("n500_float_fast", 500, 0.02, "float", "min", 4, fast=True)
sparse_case("n500_float_eps1", 500, 0.02, "float", "min", 5, eps_start=1.0)
sparse_case("n400_int_maxiter", 400, 0.03, "int", "min", 6, max_iter=150)
# BASELINE.json configs[1] (C2): 1M nnz — inputs are regenerated from the seed and verified by CRC
sparse_case("c2_float_min", 10000, 0.01, "float", "min", 0, store_inputs=False)
# configs[4] (C5) shape: one 512x512 problem at 5 %
sparse_case("c5_one", 512, 0.05, "float", "min", 11)

# --- dense input with a -1 mask and a rectangular N < M problem (auction_.pyx:528-571)
rng = np.random.default_rng(21)
matd = rng.uniform(0, 100, (40, 40))
matd[rng.random((40, 40)) > 0.3] = -1
matd[np.arange(40), rng.permutation(40)] = rng.uniform(0, 100, 40)
r = ref.auction_solve(mat=matd.copy(), problem="max")
save("dense_masked_40", mat=matd, problem="max", sol=r["sol"], meta=meta_arr(r["meta"]))
matr = rng.integers(1, 50, (12, 20)).astype(np.float64)
r = ref.auction_solve(mat=matr.copy(), problem="min")
save("dense_rect_12x20", mat=matr, problem="min", sol=r["sol"], meta=meta_arr(r["meta"]))

# --- Hopcroft-Karp: random graphs with deficient matchings (feasibility_.pyx:199-221)
for i, (n, m, e) in enumerate([(50, 50, 90), (300, 280, 700), (2000, 2000, 5000)]):
    rng = np.random.default_rng(100 + i)
    key = np.unique(rng.integers(0, n, e).astype(np.int64) * m + rng.integers(0, m, e))
    loc = np.stack([key // m, key % m], -1).astype(np.int32)
    r = ref.hopcroft_solve(loc=loc)
    save(f"hopcroft_{n}x{m}", loc=loc, size=r["size"], left=r["left_pairings"], right=r["right_pairings"])
