"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the oracle and the reference's golden
vectors.  Bar: bit-exact `sol`, `its`, `nreductions`, prices and meta (integer AND float costs — the device kernel
reproduces the reference's Jacobi trajectory including its tie rules); Hopcroft-Karp cardinality bit-exact."""
import ctypes as C
import zlib

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import (assert_meta_equal, assert_valid_matching, dense_golden_names, hopcroft_golden_names, load_golden,
                      sparse_golden_names)
from sslap_b200.datagen import make_problem, objective

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import sslap_b200
    from sslap_b200 import _native as nat
    h = nat.default_handle()
    return sslap_b200, nat, h


def prices_of(nat, h, m):
    p = np.empty(m, dtype=np.float64)
    assert nat.load().sslapb_get_prices(h.ptr, p.ctypes.data) == 0
    return p


@pytest.mark.parametrize("t_small", [32, 4, 0])
@pytest.mark.parametrize("name", sparse_golden_names())
def test_sparse_golden_bit_exact(gpu, oracle_mod, name, t_small):
    """Every regime split (warp-list / grid only) must reproduce the reference's sol + meta exactly."""
    sslap_b200, nat, h = gpu
    g = load_golden(name)
    n = int(g["n"])
    h.set_option("t_small", t_small)
    try:
        r = sslap_b200.auction_solve(loc=g["loc"], val=g["val"], size=(n, n), problem=g["problem"],
                                     cardinality_check=False, **g["kwargs"])
    finally:
        h.set_option("t_small", 32)
    assert np.array_equal(r["sol"], g["sol"])
    assert_meta_equal(r["meta"], g["meta"])
    kw = {k: v for k, v in g["kwargs"].items() if k != "fast"}
    if not g["kwargs"].get("fast"):
        o = oracle_mod.auction_solve(loc=g["loc"], val=g["val"], problem=g["problem"], return_prices=True, **kw)
        assert np.array_equal(prices_of(nat, h, n), o["prices"])       # float64 prices identical bit for bit


@pytest.mark.parametrize("name", dense_golden_names())
def test_dense_golden_bit_exact(gpu, name):
    sslap_b200, nat, h = gpu
    g = load_golden(name)
    r = sslap_b200.auction_solve(mat=g["mat"], problem=g["problem"])
    assert np.array_equal(r["sol"], g["sol"])
    assert_meta_equal(r["meta"], g["meta"])


def test_coo_matrix_input_and_no_mutation(gpu):
    sslap_b200, nat, h = gpu
    g = load_golden("example_sparse")
    mat = g["mat"].copy()
    mat[mat < 0] = 0                                      # examples/test_auction.py:33-46 (0 = missing for scipy)
    coo = sp.coo_matrix(mat)
    data_before = coo.data.copy()
    r = sslap_b200.auction_solve(coo_mat=coo, problem="max")
    assert r["sol"].tolist() == [0, 3, 4, 2, 1] and r["meta"]["obj"] == 23.681
    assert np.array_equal(coo.data, data_before)          # the reference negates in place for 'min'; we never do
    loc, val = make_problem(200, 0.05, "float", seed=2)
    v0 = val.copy()
    sslap_b200.auction_solve(loc=loc.astype(np.int64), val=val, size=(200, 200), problem="min")
    assert np.array_equal(val, v0)


def test_c2_matches_reference_and_oracle(gpu, oracle_mod):
    """BASELINE.json configs[1]: 10k x 10k, 1 %, float — sol/its identical to the reference run recorded in golden."""
    sslap_b200, nat, h = gpu
    g = load_golden("c2_float_min")
    loc, val = make_problem(10000, 0.01, "float", seed=0)
    assert zlib.crc32(loc.tobytes()) == int(g["loc_crc"][0])
    r = sslap_b200.auction_solve(loc=loc, val=val, size=(10000, 10000), problem="min")     # HK check on
    assert np.array_equal(r["sol"], g["sol"])
    assert_meta_equal(r["meta"], g["meta"])
    assert abs(objective(loc, val, r["sol"]) - 16336.192344) < 1e-5


def test_bid_sweep_kernel_bit_exact(gpu, oracle_mod):
    """Kernel-level parity of the bidding sweep (auction_.pyx:339-365) on a scattered frontier with random prices."""
    sslap_b200, nat, h = gpu
    L = nat.load()
    for (n, d, mode, seed) in [(1000, 0.01, "int", 3), (4000, 0.02, "float", 4), (257, 0.5, "int", 5)]:
        loc, val = make_problem(n, d, mode, seed=seed)
        h.set_option("small_path", 0)                     # the sweep needs the CSR resident on the handle
        try:
            sslap_b200.auction_solve(loc=loc, val=val, size=(n, n), cardinality_check=False, max_iter=1)
        finally:
            h.set_option("small_path", 1)
        rng = np.random.default_rng(seed)
        prices = rng.integers(0, 40, n).astype(np.float64) if mode == "int" else rng.uniform(0, 50, n)
        bidders = rng.permutation(n).astype(np.int32)[: max(1, n // 2)]
        jb = np.empty(bidders.size, dtype=np.int32)
        bd = np.empty(bidders.size, dtype=np.float64)
        ms = C.c_float(0)
        rc = L.sslapb_bid_sweep(h.ptr, prices.ctypes.data, bidders.ctypes.data, bidders.size, 0.37, 1, 1, 0,
                                jb.ctypes.data, bd.ctypes.data, C.byref(ms))
        assert rc == 0
        rowptr = np.searchsorted(loc[:, 0], np.arange(n + 1)).astype(np.int64)
        oj, ob = oracle_mod.bid_sweep(rowptr, loc[:, 1], -val, prices, bidders, 0.37)
        assert np.array_equal(jb, oj) and np.array_equal(bd, ob)


def test_bid_sweep_streamed_bit_exact(gpu, oracle_mod):
    """Full-frontier sweep through every variant of the kernel — per-row (default), TMA ring (merge bit 2), software-pipelined
    (bit 3, with its CTA sizes) and four rows per warp (bit 7): all bit-exact against the oracle, with and without bound
    pruning, incl. long rows, a rectangular problem and +inf prices."""
    sslap_b200, nat, h = gpu
    L = nat.load()
    cases = [(1000, 0.01, "int", 3, None), (3000, 0.4, "float", 6, None), (20000, 0.0002, "float", 7, None),
             (5000, 0.01, "float", 8, 7000), (64, 1.0, "float", 9, None)]
    for (n, d, mode, seed, m) in cases:
        loc, val = make_problem(n, d, mode, seed=seed, m=m)
        M = m or n
        h.set_option("small_path", 0)                     # the sweep needs the CSR resident on the handle
        try:
            sslap_b200.auction_solve(loc=loc, val=val, size=(n, M), cardinality_check=False, max_iter=1)
        finally:
            h.set_option("small_path", 1)
        rng = np.random.default_rng(seed)
        rowptr = np.searchsorted(loc[:, 0], np.arange(n + 1)).astype(np.int64)
        for kind in range(3):
            if kind == 0:
                prices = rng.integers(0, 40, M).astype(np.float64) if mode == "int" else rng.uniform(0, 50, M)
            elif kind == 1:
                prices = np.zeros(M)
            else:
                prices = rng.uniform(0, 5, M)
                prices[rng.integers(0, M, max(1, M // 50))] = np.inf
            oj, ob = oracle_mod.bid_sweep(rowptr, loc[:, 1], -val, prices, np.arange(n, dtype=np.int32), 0.37)
            for merge in (0, 2, 4, 6, 128, 130, 8, 10, 8 | 16, 8 | 32 | 2, 8 | 48, 8 | 64, 256, 258, 512, 514):   # per-row (default), TMA ring, 4 rows/warp, pipelined, hot form, lean
                jb = np.empty(n, dtype=np.int32)
                bd = np.empty(n, dtype=np.float64)
                ms = C.c_float(0)
                rc = L.sslapb_bid_sweep(h.ptr, prices.ctypes.data, None, n, 0.37, merge, 1, 0, jb.ctypes.data,
                                        bd.ctypes.data, C.byref(ms))
                assert rc == 0, h.last_error()
                assert np.array_equal(jb, oj) and np.array_equal(bd, ob), (n, d, mode, kind, merge)


@pytest.mark.parametrize("name", hopcroft_golden_names())
def test_hopcroft_cardinality_bit_exact(gpu, name):
    sslap_b200, nat, h = gpu
    g = load_golden(name)
    loc = g["loc"]
    if name == "example_hopcroft":
        lookup = {}
        for i, j in loc.tolist():
            lookup.setdefault(i, []).append(j)
        r = sslap_b200.hopcroft_solve(lookup=lookup)
    else:
        r = sslap_b200.hopcroft_solve(loc=loc)
    assert r["size"] == int(g["size"])
    assert r["left_pairings"].dtype == np.int32 and r["left_pairings"].shape == g["left"].shape
    assert r["right_pairings"].shape == g["right"].shape
    assert_valid_matching(r, loc)


def test_hopcroft_dense_and_int64_inputs(gpu, oracle_mod):
    sslap_b200, nat, h = gpu
    mat = -np.ones((5, 5))
    mat[[0, 0, 1, 1, 2, 2, 3, 4], [0, 1, 1, 2, 1, 4, 2, 3]] = 1          # examples/test_feasibility.py:22-26
    r = sslap_b200.hopcroft_solve(mat=mat)
    assert r["size"] == 5
    loc = np.array([[0, 0], [0, 1], [1, 1], [1, 2], [2, 1], [2, 4], [3, 2], [4, 3]])   # int64: the reference raises here
    r = sslap_b200.hopcroft_solve(loc=loc)
    assert r["size"] == 5
    assert_valid_matching(r, loc)
    rng = np.random.default_rng(9)
    n = 20000
    key = np.unique(rng.integers(0, n, 3 * n).astype(np.int64) * n + rng.integers(0, n, 3 * n))
    loc = np.stack([key // n, key % n], -1).astype(np.int32)
    r = sslap_b200.hopcroft_solve(loc=loc)
    assert r["size"] == oracle_mod.hopcroft_solve(loc=loc)["size"]
    assert_valid_matching(r, loc)


def test_infeasible_inputs_raise_like_the_reference(gpu):
    sslap_b200, nat, h = gpu
    # fewer entries than rows (auction_.pyx:559-560 / :604-605)
    loc = np.array([[0, 0], [1, 1]], dtype=np.int32)
    with pytest.raises(ValueError, match="Fewer than 3 valid values provided for 3 rows"):
        sslap_b200.auction_solve(loc=loc, val=np.ones(2), size=(3, 3))
    # cardinality < N (auction_.pyx:565-566): rows 0 and 1 both only reach column 0
    mat = -np.ones((3, 3))
    mat[0, 0] = 1; mat[1, 0] = 2; mat[2, 1] = 3; mat[2, 2] = 1
    with pytest.raises(ValueError, match=r"Maximum matching possible only involves 2 out of 3 rows"):
        sslap_b200.auction_solve(mat=mat)
    # 6 x 4 is infeasible, 4 x 6 is fine
    rng = np.random.default_rng(1)
    with pytest.raises(ValueError, match="infeasible"):
        sslap_b200.auction_solve(mat=rng.uniform(1, 9, (6, 4)))
    r = sslap_b200.auction_solve(mat=rng.uniform(1, 9, (4, 6)))
    assert len(set(r["sol"].tolist())) == 4 and r["meta"]["soln_found"] == 1


def test_unsorted_input_is_sorted_on_the_device(gpu, oracle_mod):
    """The reference needs row-sorted `loc` (auction_.pyx:33-48; unsorted input silently yields garbage there).  The CUDA
    path detects an unsorted stream and sorts it on the GPU, STABLY: shuffling whole rows around (keeping the order
    inside each row) must give exactly the result of the sorted stream, ties included."""
    sslap_b200, nat, h = gpu
    rng = np.random.default_rng(12)
    for (n, d, mode) in [(300, 0.05, "int"), (2000, 0.01, "float"), (70000, 0.0005, "float")]:
        loc, val = make_problem(n, d, mode, seed=13)
        want = oracle_mod.auction_solve(loc=loc, val=val, problem="min")
        # shuffled stream: rows dealt into 50 random blocks, original order kept inside every row (what a stable sort undoes)
        block = rng.integers(0, 50, n)[loc[:, 0]]
        shuffled = np.lexsort((np.arange(len(val)), block))
        loc_u, val_u = loc[shuffled], val[shuffled]
        assert not np.all(np.diff(loc_u[:, 0]) >= 0)
        for dtype in (np.int32, np.int64):
            got = sslap_b200.auction_solve(loc=loc_u.astype(dtype), val=val_u, size=(n, n), problem="min",
                                           cardinality_check=(n <= 2000))
            assert np.array_equal(got["sol"], want["sol"])
            assert_meta_equal(got["meta"], want["meta"])
    hk = sslap_b200.hopcroft_solve(loc=loc_u)
    assert hk["size"] == n


def test_max_iter_returns_partial_assignment(gpu, oracle_mod):
    sslap_b200, nat, h = gpu
    loc, val = make_problem(300, 0.05, "int", seed=8)
    for mi in (1, 7, 60):
        r = sslap_b200.auction_solve(loc=loc, val=val, size=(300, 300), max_iter=mi, cardinality_check=False)
        o = oracle_mod.auction_solve(loc=loc, val=val, max_iter=mi)
        assert np.array_equal(r["sol"], o["sol"]) and (r["sol"] == -1).any()
        assert_meta_equal(r["meta"], o["meta"])
        assert r["meta"]["soln_found"] == 0


def test_c3_full_size_properties_and_oracle(gpu, oracle_mod):
    """BASELINE.json configs[2] at full size (N=100k, 0.1 %, ~10M nnz, float): valid perfect matching, eps-CS at
    target eps, and bit-exact agreement with the oracle's trajectory."""
    sslap_b200, nat, h = gpu
    n = 100000
    loc, val = make_problem(n, 0.001, "float", seed=0)
    r = sslap_b200.auction_solve(loc=loc, val=val, size=(n, n), problem="min", cardinality_check=False)
    sol = r["sol"]
    assert np.array_equal(np.sort(sol), np.arange(n))                      # perfect matching
    obj = objective(loc, val, sol)                                          # every (i, sol[i]) is an entry
    assert r["meta"]["eCE"] == 1 and r["meta"]["soln_found"] == 1
    # eps-complementary slackness re-checked on the host from the device prices (auction_.pyx:443-485)
    p = prices_of(nat, h, n)
    rows = loc[:, 0].astype(np.int64)
    v = -val - p[loc[:, 1]]
    best = np.full(n, -np.inf)
    np.maximum.at(best, rows, v)
    key = rows * n + loc[:, 1]
    chosen = np.searchsorted(key, np.arange(n, dtype=np.int64) * n + sol)
    assert (v[chosen] + 1e-7 >= best - np.float32(1.0 / n)).all()
    o = oracle_mod.auction_solve(loc=loc, val=val, problem="min")
    assert np.array_equal(sol, o["sol"])
    assert_meta_equal(r["meta"], o["meta"])
    assert abs(obj - 162903.850803) < 1e-4                                  # SURVEY.md §6.2 (reference run)


def test_batch_matches_single_problem_calls(gpu, oracle_mod):
    """BASELINE.json configs[4] shape (512 x 512 at 5 %): every problem of a batch must come back bit-identical to the
    oracle's (= the reference's) answer for that problem, ties and max_iter included."""
    sslap_b200, nat, h = gpu
    probs, want = [], []
    for k in range(40):
        n = 512 if k < 8 else int(20 + 13 * k)
        mode = "int" if k % 3 == 0 else "float"
        loc, val = make_problem(n, 0.05 if n == 512 else 0.2, mode, seed=100 + k)
        probs.append((loc, val, (n, n)))
        want.append(oracle_mod.auction_solve(loc=loc, val=val, problem="max"))
    for v1 in (0, 1):                                     # the sub-warp kernel (default) and round 1's kernel
        h.set_option("batch_v1", v1)
        try:
            got = sslap_b200.auction_solve_batch(probs, problem="max")
        finally:
            h.set_option("batch_v1", 0)
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert np.array_equal(g["sol"], w["sol"])
            assert_meta_equal(g["meta"], w["meta"])
    # golden: the reference's own result for one 512 x 512 problem, run as a batch of copies
    gold = load_golden("c5_one")
    res = sslap_b200.auction_solve_batch([(gold["loc"], gold["val"], (512, 512))] * 5, problem=gold["problem"])
    for r in res:
        assert np.array_equal(r["sol"], gold["sol"])
        assert_meta_equal(r["meta"], gold["meta"])
    # max_iter truncation and eps_start / fast
    loc, val = make_problem(200, 0.1, "int", seed=5)
    for kw in (dict(max_iter=37), dict(eps_start=1.0), dict(fast=True)):
        r = sslap_b200.auction_solve_batch([(loc, val, (200, 200))] * 3, **kw)[1]
        o = oracle_mod.auction_solve(loc=loc, val=val, **{k: v for k, v in kw.items() if k != "fast"},
                                     **({"eps_start": float(np.float32(1.0 / 200))} if kw.get("fast") else {}))
        assert np.array_equal(r["sol"], o["sol"])
        assert_meta_equal(r["meta"], o["meta"])


def _single_entry_rows_problem(n, seed):
    """Feasible problem in which a few persons have exactly ONE admissible object: they bid +inf (w_i = -inf,
    auction_.pyx:344,360), prices become +inf and other persons see -inf values for those objects."""
    rng = np.random.default_rng(seed)
    loc, val = make_problem(n, 0.15, "float", seed=seed)
    perm_rows = rng.permutation(n)[:4]
    keep = np.ones(len(val), dtype=bool)
    for r in perm_rows:
        idx = np.nonzero(loc[:, 0] == r)[0]
        keep[idx[1:]] = False                              # row r keeps only its first entry
    loc, val = loc[keep], val[keep]
    # drop rows that now collide on the same single object (keeps the instance feasible in practice)
    return loc, val


def test_single_entry_rows_and_infinite_prices(gpu, oracle_mod):
    sslap_b200, nat, h = gpu
    ok = 0
    for seed in range(12):
        loc, val = _single_entry_rows_problem(60, seed)
        if oracle_mod.hopcroft_solve(loc=loc)["size"] < 60:
            continue                                       # the surgery made it infeasible: skip
        for t_small in (32, 0):
            h.set_option("t_small", t_small)
            try:
                g = sslap_b200.auction_solve(loc=loc, val=val, size=(60, 60), problem="max", max_iter=20000)
            finally:
                h.set_option("t_small", 32)
            o = oracle_mod.auction_solve(loc=loc, val=val, problem="max", max_iter=20000)
            assert np.array_equal(g["sol"], o["sol"])
            assert_meta_equal(g["meta"], o["meta"], keys=("eCE", "its", "nreductions", "soln_found", "n_assigned"))
        ok += 1
    assert ok >= 4


def test_rectangular_duplicates_and_odd_values(gpu, oracle_mod):
    sslap_b200, nat, h = gpu
    rng = np.random.default_rng(4)
    # N < M through loc/val, values include 0, negatives and repeats; problem 'max' and 'min'
    loc, val = make_problem(40, 0.3, "int", seed=9, m=70)
    val = val - 50.0
    val[::7] = 0.0
    for problem in ("min", "max"):
        g = sslap_b200.auction_solve(loc=loc, val=val, size=(40, 70), problem=problem)
        o = oracle_mod.auction_solve(loc=loc, val=val, problem=problem)
        assert np.array_equal(g["sol"], o["sol"])
        assert_meta_equal(g["meta"], o["meta"])
    # duplicate (i, j) entries: the row sweep treats them as separate candidates, get_obj adds every match (auction_.pyx:514-521)
    loc2, val2 = make_problem(30, 0.3, "float", seed=10)
    dup = rng.integers(0, len(val2), 25)
    order = np.argsort(np.concatenate([np.arange(len(val2)), dup + 0.5]), kind="stable")
    loc_d = np.concatenate([loc2, loc2[dup]])[order]
    val_d = np.concatenate([val2, rng.uniform(0, 100, 25)])[order]
    assert np.all(np.diff(loc_d[:, 0]) >= 0)
    g = sslap_b200.auction_solve(loc=loc_d, val=val_d, size=(30, 30), problem="min")
    o = oracle_mod.auction_solve(loc=loc_d, val=val_d, problem="min")
    assert np.array_equal(g["sol"], o["sol"])
    assert_meta_equal(g["meta"], o["meta"], keys=("eCE", "its", "nreductions", "soln_found", "n_assigned"))
    assert abs(g["meta"]["obj"] - o["meta"]["obj"]) <= 1e-3 * max(1.0, abs(o["meta"]["obj"]))


def test_long_rows_take_the_generic_sweep(gpu, oracle_mod):
    """Rows longer than one warp pass (> 125 entries) use the multi-trip sweep in every regime; dense 300 x 300 input."""
    sslap_b200, nat, h = gpu
    rng = np.random.default_rng(21)
    mat = rng.integers(1, 1000, (300, 300)).astype(np.float64)
    mat[rng.random((300, 300)) < 0.1] = -1
    want = oracle_mod.auction_solve(mat=mat, problem="min")
    for t_small in (32, 4, 0):
        h.set_option("t_small", t_small)
        try:
            got = sslap_b200.auction_solve(mat=mat, problem="min")
        finally:
            h.set_option("t_small", 32)
        assert np.array_equal(got["sol"], want["sol"])
        assert_meta_equal(got["meta"], want["meta"])


def test_very_long_rows_use_the_cooperative_kernel_instance(gpu, oracle_mod):
    """Rows of more than 1021 entries select the second instance of the persistent kernel (auction_long.cu), where the
    whole CTA sweeps such rows in the multi-bidder regime: dense inputs, a rectangular one, and a sparse problem with a
    few dense rows (pending and ordinary bidders in the same round)."""
    sslap_b200, nat, h = gpu
    rng = np.random.default_rng(33)
    cases = []
    cases.append(("dense-int", rng.integers(1, 200, (1100, 1100)).astype(np.float64), "min"))
    cases.append(("dense-float", rng.uniform(0, 100, (1300, 1300)), "max"))
    rect = rng.uniform(0, 50, (1050, 1500))
    rect[rng.random(rect.shape) < 0.05] = -1
    cases.append(("rect", rect, "max"))
    mixed = -np.ones((2500, 2500))
    mask = rng.random(mixed.shape) < 0.01
    mixed[mask] = rng.uniform(0, 100, int(mask.sum()))
    mixed[np.arange(2500), rng.permutation(2500)] = rng.uniform(0, 100, 2500)
    for r in (3, 700, 701, 1500, 2499):
        mixed[r] = rng.uniform(0, 100, 2500)
    cases.append(("mixed", mixed, "max"))
    for name, mat, problem in cases:
        want = oracle_mod.auction_solve(mat=mat, problem=problem, return_prices=True)
        for t_small in (32, 4):
            h.set_option("t_small", t_small)
            try:
                got = sslap_b200.auction_solve(mat=mat.copy(), problem=problem, cardinality_check=False)
            finally:
                h.set_option("t_small", 32)
            assert np.array_equal(got["sol"], want["sol"]), name
            assert_meta_equal(got["meta"], want["meta"])
            assert np.array_equal(prices_of(nat, h, mat.shape[1]), want["prices"]), name


def test_two_handles_in_two_threads_and_option_validation(gpu, oracle_mod):
    """One handle = one stream + its own scratch; distinct handles may be driven from different threads (the persistent
    cooperative kernels of the two solves simply take turns on the device).  Options reject bad values."""
    import threading
    sslap_b200, nat, h = gpu
    for name, value in (("t_small", 33), ("t_small", -1), ("t_cluster", 64), ("t_shard", -5), ("watchdog_ms", 0), ("no_such_option", 1)):
        with pytest.raises(ValueError):
            h.set_option(name, value)
    problems = [make_problem(n, d, "float", seed=s) for (n, d, s) in ((1500, 0.01, 1), (900, 0.03, 2), (2500, 0.004, 3))]
    wants = [oracle_mod.auction_solve(loc=l, val=v, problem="max") for (l, v) in problems]
    results, errors = {}, []

    def work(tid):
        try:
            hh = nat.Handle(0)
            for rep in range(4):
                for k, (l, v) in enumerate(problems):
                    n = int(l[:, 0].max()) + 1
                    results[(tid, rep, k)] = sslap_b200.auction_solve(loc=l, val=v, size=(n, n), problem="max",
                                                                      cardinality_check=bool(rep % 2), _handle=hh)
        except Exception as e:                                   # pragma: no cover
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert len(results) == 2 * 4 * len(problems)
    for (tid, rep, k), got in results.items():
        assert np.array_equal(got["sol"], wants[k]["sol"])
        assert_meta_equal(got["meta"], wants[k]["meta"])


def test_randomized_differential_against_the_oracle(gpu, oracle_mod):
    """120 seeded random instances across shapes, densities, cost kinds, objectives, eps options, iteration caps and
    regime splits: `sol` and the integer meta keys must equal the oracle's bit for bit; so must the float64 prices."""
    sslap_b200, nat, h = gpu
    rng = np.random.default_rng(2024)
    checked = 0
    for case in range(120):
        n = int(rng.integers(2, 400))
        m = n + int(rng.integers(0, 30)) if rng.random() < 0.3 else n
        density = float(rng.choice([0.02, 0.05, 0.1, 0.3, 0.7, 1.0]))
        mode = "int" if rng.random() < 0.5 else "float"
        loc, val = make_problem(n, density, mode, seed=1000 + case, m=m)
        if rng.random() < 0.2:
            val = np.round(val / 10.0)                      # heavy ties, zeros included
        problem = "min" if rng.random() < 0.5 else "max"
        kw = {}
        r = rng.random()
        if r < 0.15:
            kw["eps_start"] = float(rng.choice([0.5, 3.0, 40.0]))
        elif r < 0.25:
            kw["max_iter"] = int(rng.integers(1, 200))
        t_small = int(rng.choice([32, 32, 16, 4, 1, 0]))
        h.set_option("t_small", t_small)
        try:
            g = sslap_b200.auction_solve(loc=loc if case % 2 else loc.astype(np.int64), val=val, size=(n, m),
                                         problem=problem, cardinality_check=bool(case % 3), **kw)
        finally:
            h.set_option("t_small", 32)
        o = oracle_mod.auction_solve(loc=loc, val=val, problem=problem, return_prices=True, **kw)
        assert np.array_equal(g["sol"], o["sol"]), (case, n, m, density, mode, problem, kw, t_small)
        assert_meta_equal(g["meta"], o["meta"])
        assert np.array_equal(prices_of(nat, h, m), o["prices"]), (case, "prices")
        checked += 1
    assert checked == 120


def test_randomized_dense_batch_and_hopcroft(gpu, oracle_mod):
    """Seeded random instances through the other entry points: dense `mat` input (device dense -> CSR build), the batch
    call, and Hopcroft-Karp on random (mostly deficient) graphs."""
    sslap_b200, nat, h = gpu
    rng = np.random.default_rng(77)
    # dense matrices with -1 holes, rectangular allowed
    for case in range(30):
        n = int(rng.integers(2, 150)); m = n + int(rng.integers(0, 20))
        mat = rng.integers(0, 50, (n, m)).astype(np.float64) if case % 2 else rng.uniform(0, 100, (n, m))
        mat[rng.random((n, m)) < float(rng.choice([0.0, 0.3, 0.8]))] = -1
        mat[np.arange(n), rng.permutation(m)[:n]] = rng.integers(0, 50, n)      # keep it feasible
        problem = "min" if case % 3 else "max"
        g = sslap_b200.auction_solve(mat=mat, problem=problem)
        o = oracle_mod.auction_solve(mat=mat, problem=problem)
        assert np.array_equal(g["sol"], o["sol"]), ("dense", case)
        assert_meta_equal(g["meta"], o["meta"])
    # one batch of 60 heterogeneous problems
    probs, want = [], []
    for k in range(60):
        n = int(rng.integers(2, 200)); m = n + int(rng.integers(0, 10))
        loc, val = make_problem(n, float(rng.choice([0.05, 0.2, 0.6])), "int" if k % 2 else "float", seed=500 + k, m=m)
        probs.append((loc, val, (n, m)))
        want.append(oracle_mod.auction_solve(loc=loc, val=val, problem="min"))
    got = sslap_b200.auction_solve_batch(probs, problem="min")
    for k, (g, w) in enumerate(zip(got, want)):
        assert np.array_equal(g["sol"], w["sol"]), ("batch", k)
        assert_meta_equal(g["meta"], w["meta"])
    # Hopcroft-Karp on random graphs
    for case in range(40):
        n = int(rng.integers(1, 3000)); m = int(rng.integers(1, 3000)); e = int(rng.integers(1, 4 * max(n, m)))
        key = np.unique(rng.integers(0, n, e).astype(np.int64) * m + rng.integers(0, m, e))
        loc = np.stack([key // m, key % m], -1).astype(np.int32 if case % 2 else np.int64)
        g = sslap_b200.hopcroft_solve(loc=loc)
        o = oracle_mod.hopcroft_solve(loc=loc)
        assert g["size"] == o["size"], ("hk", case)
        assert_valid_matching(g, loc)
