"""GPU tests (-m gpu) of the BASELINE.json configurations that round 1 left to tool logs (C4 at full size, the full
4096-problem C5 batch, C3 against the LIVE reference), of the boundary's thread safety, of the warm-start / strict
extensions, and of the row-sharded solve run with K virtual ranks on ONE GPU (same kernels, K concurrent
persistent kernels, bit-identical prices / sol / its to K = 1)."""
import ctypes as C
import threading

import numpy as np
import pytest

from conftest import assert_meta_equal
from sslap_b200.datagen import make_problem, objective

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import sslap_b200
    from sslap_b200 import _native as nat
    return sslap_b200, nat, nat.default_handle()


def test_c3_against_the_live_reference(gpu):
    """BASELINE.json configs[2] against the UNMODIFIED reference (oracle/_ref, built from /root/reference by
    oracle/build_ref.sh and shipped to the GPU box): sol, its, nreductions and every meta key identical."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref (the compiled reference) is not present")
    sslap_b200, nat, h = gpu
    ref = ref_loader.load()
    n = 100000
    loc, val = make_problem(n, 0.001, "float", seed=0)
    want = ref.auction_solve(loc=loc, val=val.copy(), size=(n, n), problem="min", cardinality_check=False)
    got = sslap_b200.auction_solve(loc=loc, val=val, size=(n, n), problem="min", cardinality_check=False)
    assert np.array_equal(got["sol"], want["sol"])
    assert_meta_equal(got["meta"], want["meta"])
    assert got["meta"]["its"] == 202985


def test_c4_full_size_with_hopcroft_karp(gpu):
    """BASELINE.json configs[3]: N = 1M, 0.01 % (~101M nnz), float costs, Hopcroft-Karp check ON.  The reference cannot run
    this configuration as asked (its HK mallocs an N^2 queue and segfaults, feasibility_.pyx:132; README.md:40), so the
    pins are the reference's own auction run recorded in BASELINE.md / SURVEY.md §6.2 (max_iter = 5e7,
    cardinality_check=False: its = 1,935,709, objective 1630435.704360) and scipy's maximum matching for the check."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import maximum_bipartite_matching
    sslap_b200, nat, h = gpu
    n = 1000000
    loc, val = make_problem(n, 1e-4, "float", seed=0)
    assert val.size == 100994837
    r = sslap_b200.auction_solve(loc=loc, val=val, size=(n, n), problem="min", cardinality_check=True,
                                 max_iter=50000000, _raw_meta=True)
    sol = r["sol"]
    assert np.array_equal(np.sort(sol), np.arange(n))                       # perfect matching
    assert r["meta"]["soln_found"] == 1 and r["meta"]["eCE"] == 1
    assert r["meta"]["its"] == 1935709
    assert abs(objective(loc, val, sol) - 1630435.704360) <= 1e-3
    graph = sp.csr_matrix((np.ones(val.size, dtype=np.int8), (loc[:, 0], loc[:, 1])), shape=(n, n))
    card = int((maximum_bipartite_matching(graph, perm_type="column") >= 0).sum())
    assert r["raw"].cardinality == card == n
    hk = sslap_b200.hopcroft_solve(loc=loc)                                 # stand-alone entry on the same graph
    assert hk["size"] == card
    lp, rp = hk["left_pairings"], hk["right_pairings"]
    assert (lp >= 0).all() and np.array_equal(rp[lp], np.arange(n))
    key = loc[:, 0].astype(np.int64) * n + loc[:, 1]
    pos = np.searchsorted(key, np.arange(n, dtype=np.int64) * n + lp)
    assert np.array_equal(key[pos], np.arange(n, dtype=np.int64) * n + lp)  # every matched pair is an edge


def test_c5_full_batch_of_4096(gpu, oracle_mod):
    """BASELINE.json configs[4] at full size: 4096 independent 512 x 512 problems at 5 % (seeds 0..4095) in ONE batch
    call; every problem's sol / its / meta equal to the oracle's (= the reference's) answer for that problem."""
    sslap_b200, nat, h = gpu
    probs = []
    for k in range(4096):
        loc, val = make_problem(512, 0.05, "float", seed=k)
        probs.append((loc, val, (512, 512)))
    got = sslap_b200.auction_solve_batch(probs, problem="min")
    assert len(got) == 4096
    for k, (loc, val, _) in enumerate(probs):
        want = oracle_mod.auction_solve(loc=loc, val=val, problem="min")
        assert np.array_equal(got[k]["sol"], want["sol"]), k
        assert_meta_equal(got[k]["meta"], want["meta"])


def test_two_threads_share_the_default_handle(gpu, oracle_mod):
    """The reference is re-entrant under the GIL (auction_solve.py:6-55); ctypes releases the GIL, so the library
    serialises concurrent calls on one handle.  Two threads hammer `auction_solve` WITHOUT `_handle=`."""
    sslap_b200, nat, h = gpu
    problems = [make_problem(n, d, mode, seed=s) for (n, d, mode, s) in
                ((1200, 0.01, "float", 11), (300, 0.2, "int", 12), (4000, 0.003, "float", 13), (64, 1.0, "int", 14))]
    wants = [oracle_mod.auction_solve(loc=l, val=v, problem="min") for (l, v) in problems]
    errors, results = [], {}

    def work(tid):
        try:
            for rep in range(6):
                for k, (l, v) in enumerate(problems):
                    n = int(l[:, 0].max()) + 1
                    results[(tid, rep, k)] = sslap_b200.auction_solve(loc=l, val=v, size=(n, n), problem="min",
                                                                      cardinality_check=bool((rep + tid) % 2))
                    hk = sslap_b200.hopcroft_solve(loc=l)
                    assert hk["size"] == n
        except Exception as e:                                               # pragma: no cover
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert len(results) == 2 * 6 * len(problems)
    for (tid, rep, k), got in results.items():
        assert np.array_equal(got["sol"], wants[k]["sol"])
        assert_meta_equal(got["meta"], wants[k]["meta"])


def test_warm_start_and_strict_mode(gpu, oracle_mod):
    """Extensions of SURVEY.md §8(f) rank 4.  Warm start: a solve started from given prices follows exactly the trajectory
    the oracle follows from the same prices, and re-solving a perturbed problem from the previous prices with a small
    eps_start reaches the cold solve's optimal objective in far fewer rounds.  Strict mode: eps = 1/(N+1), zero
    tolerance — for integer costs the objective equals scipy's optimum."""
    from scipy.optimize import linear_sum_assignment
    sslap_b200, nat, h = gpu
    rng = np.random.default_rng(5)
    for (n, d, mode, seed) in ((400, 0.05, "float", 1), (250, 0.2, "int", 2), (3000, 0.004, "float", 3)):
        loc, val = make_problem(n, d, mode, seed=seed)
        cold = sslap_b200.auction_solve(loc=loc, val=val, size=(n, n), problem="min", return_prices=True)
        o = oracle_mod.auction_solve(loc=loc, val=val, problem="min", return_prices=True)
        assert np.array_equal(cold["prices"], o["prices"])
        # (a) arbitrary start prices: bit-exact against the oracle started from the same prices
        p0 = rng.uniform(0, 30, n) if mode == "float" else rng.integers(0, 30, n).astype(np.float64)
        g = sslap_b200.auction_solve(loc=loc, val=val, size=(n, n), problem="min", prices_in=p0, return_prices=True,
                                     _raw_meta=True)
        w = oracle_mod.auction_solve(loc=loc, val=val, problem="min", prices_in=p0, return_prices=True)
        assert g["raw"].warm_start == 1
        assert np.array_equal(g["sol"], w["sol"])
        assert_meta_equal(g["meta"], w["meta"])
        assert np.array_equal(g["prices"], w["prices"])
        # the warm start is consumed by ONE call
        again = sslap_b200.auction_solve(loc=loc, val=val, size=(n, n), problem="min", _raw_meta=True)
        assert again["raw"].warm_start == 0 and np.array_equal(again["sol"], cold["sol"])
        # (b) the intended use: perturb a few costs, restart from the old prices with a small eps
        val2 = val.copy()
        idx = rng.integers(0, val.size, 5)
        val2[idx] = val2[idx] + (rng.integers(1, 4, 5) if mode == "int" else rng.uniform(0.5, 3.0, 5))
        cold2 = sslap_b200.auction_solve(loc=loc, val=val2, size=(n, n), problem="min")
        warm2 = sslap_b200.auction_solve(loc=loc, val=val2, size=(n, n), problem="min", prices_in=cold["prices"],
                                         eps_start=float(np.float32(1.0 / n)))
        assert warm2["meta"]["soln_found"] == 1
        o_warm, o_cold = objective(loc, val2, warm2["sol"]), objective(loc, val2, cold2["sol"])
        if mode == "float":
            assert abs(o_warm - o_cold) <= 1.0 + 1e-6                       # both within N * eps_final of the optimum
            assert warm2["meta"]["its"] < cold2["meta"]["its"]              # and the warm start skips the coarse phases
        else:
            assert o_warm == o_cold
    with pytest.raises(RuntimeError, match="n_cols differs"):
        loc, val = make_problem(50, 0.3, "float", seed=4)
        sslap_b200.auction_solve(loc=loc, val=val, size=(50, 50), prices_in=np.zeros(49))
    # strict mode on integer costs with heavy ties: optimum = scipy's, trajectory = the oracle's strict run
    h.set_option("strict", 1)
    try:
        for seed in range(6):
            n = 60 + 20 * seed
            mat = rng.integers(1, 12, (n, n)).astype(np.float64)
            g = sslap_b200.auction_solve(mat=mat, problem="min", _raw_meta=True)
            assert g["raw"].strict == 1
            r, c = linear_sum_assignment(mat)
            assert float(mat[np.arange(n), g["sol"]].sum()) == float(mat[r, c].sum())
            w = oracle_mod.auction_solve(mat=mat, problem="min", strict=True)
            assert np.array_equal(g["sol"], w["sol"])
            assert_meta_equal(g["meta"], w["meta"])
    finally:
        h.set_option("strict", 0)


def test_reference_benchmarking_sweeps_at_reduced_size(gpu):
    """The reference's own benchmark (benchmarking.py:29-147: seeded dense matrices, size sweep at 100 % density in float
    and int costs, density sweep at fixed size, problem='max', dense `mat` input with the Hopcroft-Karp check on) at
    reduced size, with the checks that script computes but never asserts (tools/benchmarking.py): complete assignment,
    only admissible entries, objective = scipy's optimum (exactly for int costs, within N * eps for float), `sol` equal to
    the live reference's when oracle/_ref is present — and, where the margin is wide (dense N = 1000: 18 ms against the
    reference's 63 ms), faster than the reference on the box's host."""
    import os
    import sys
    import time
    sslap_b200, nat, h = gpu
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import benchmarking as bm
    from oracle import ref_loader
    ref = ref_loader.load() if ref_loader.available() else None

    def run(mat, mode):
        n = mat.shape[0]
        res = sslap_b200.auction_solve(mat, problem="max")
        obj = bm.check(mat, res["sol"], mode)
        best = bm.scipy_optimum(mat)
        if mode == "int":
            assert obj == best, (n, mode, obj, best)
        else:
            assert abs(obj - best) <= n * (1.0 / n + 1e-7) + 1e-6, (n, mode, obj, best)
        if ref is not None:
            rr = ref.auction_solve(mat.copy(), problem="max")
            assert np.array_equal(rr["sol"], res["sol"]), (n, mode)
            assert rr["meta"]["its"] == res["meta"]["its"]

    for n in (10, 32, 100, 316, 562):
        for mode in ("float", "int"):
            run(bm.make_matrix(n, 1.0, mode), mode)
    for dens in (0.05, 0.2, 0.5):
        run(bm.make_matrix(500, dens, "float"), "float")
    run(bm.make_matrix(1000, 0.01, "float"), "float")
    # the same generator at N = 500, 1 % (rows of ~5 entries, several with one): single-choice bidders price objects at +inf
    # and the auction runs into max_iter — in the reference too.  The GPU must stop on the very same state.
    from oracle import oracle
    mat = bm.make_matrix(500, 0.01, "float")
    want = oracle.auction_solve(mat=mat, problem="max")
    got = sslap_b200.auction_solve(mat, problem="max")
    assert want["meta"]["its"] == 1000000 and want["meta"]["soln_found"] == 0
    assert np.array_equal(got["sol"], want["sol"])
    for k in ("its", "nreductions", "soln_found", "n_assigned", "eCE"):
        assert got["meta"][k] == want["meta"][k], k
    if ref is not None:
        mat = bm.make_matrix(1000, 1.0, "float")
        sslap_b200.auction_solve(mat, problem="max")
        t = time.perf_counter()
        sslap_b200.auction_solve(mat, problem="max")
        t_gpu = time.perf_counter() - t
        t = time.perf_counter()
        ref.auction_solve(mat.copy(), problem="max")
        t_ref = time.perf_counter() - t
        assert t_gpu < t_ref, (t_gpu, t_ref)


def test_hot_lists_decide_bids_and_change_nothing(gpu, oracle_mod):
    """Hot lists (csrc/hot.cu): the 32 largest entries of a row decide a bid only when that is provably exact.  The solve
    must use them (from the third eps-phase on most bids are decided there, some are handed on to the full-row sweep),
    must skip compaction in rounds without a hole, and sol / meta / float64 prices must equal the oracle's and the
    hot-less run's, bit for bit — sparse float and int costs, wide rows (more than one warp pass) and a dense matrix."""
    sslap_b200, nat, h = gpu
    cases = [(4000, 0.02, "float", 11), (3000, 0.03, "int", 12), (1500, 0.2, "float", 13), (20000, 0.003, "float", 14)]
    for (n, d, mode, seed) in cases:
        loc, val = make_problem(n, d, mode, seed=seed)
        want = oracle_mod.auction_solve(loc=loc, val=val, problem="min", return_prices=True)
        got = sslap_b200.auction_solve(loc=loc, val=val, size=(n, n), problem="min", cardinality_check=False, _raw_meta=True,
                                       return_prices=True)
        raw = got["raw"]
        assert raw.hot_grid_bids > 0 and raw.hot_tail_rounds > 0, (n, raw.hot_grid_bids, raw.hot_tail_rounds)
        assert raw.hot_grid_fallbacks > 0                                    # the first phase has no bounds yet: all handed on
        assert 0 <= raw.rounds_nohole <= raw.rounds_grid                     # (small frontiers: the mid regime takes them)
        h.set_option("hot", 0)
        try:
            off = sslap_b200.auction_solve(loc=loc, val=val, size=(n, n), problem="min", cardinality_check=False,
                                           _raw_meta=True, return_prices=True)
        finally:
            h.set_option("hot", 1)
        assert off["raw"].hot_grid_bids == 0 and off["raw"].hot_tail_rounds == 0 and off["raw"].rounds_mid == 0
        assert 0 < off["raw"].rounds_nohole <= off["raw"].rounds_grid
        for r in (got, off):
            assert np.array_equal(r["sol"], want["sol"]), (n, mode)
            assert_meta_equal(r["meta"], want["meta"])
            assert np.array_equal(r["prices"], want["prices"]), (n, mode)
    # exactness guard: where the smallest eps comes within rounding distance of the prices (|a| * N > 1e14) neither the hot
    # lists nor the bound pruning are used — and the trajectory is still the oracle's
    loc, val = make_problem(600, 0.05, "float", seed=21)
    val = val * 1e12
    want = oracle_mod.auction_solve(loc=loc, val=val, problem="min", return_prices=True, max_iter=200000)
    got = sslap_b200.auction_solve(loc=loc, val=val, size=(600, 600), problem="min", cardinality_check=False, _raw_meta=True,
                                   return_prices=True, max_iter=200000)
    assert got["raw"].hot_grid_bids == 0 and got["raw"].hot_tail_rounds == 0 and got["raw"].prune_second_pass == 0
    assert np.array_equal(got["sol"], want["sol"]) and np.array_equal(got["prices"], want["prices"])
    assert_meta_equal(got["meta"], want["meta"])
    rng = np.random.default_rng(5)
    mat = rng.uniform(0, 100, (1300, 1300))                                  # dense: rows of 1300 entries, the long-row instance
    want = oracle_mod.auction_solve(mat=mat, problem="max", return_prices=True)
    got = sslap_b200.auction_solve(mat=mat, problem="max", _raw_meta=True, return_prices=True)
    assert got["raw"].hot_tail_rounds > 0
    assert np.array_equal(got["sol"], want["sol"]) and np.array_equal(got["prices"], want["prices"])
    assert_meta_equal(got["meta"], want["meta"])


def test_mid_regime_runs_and_changes_nothing(gpu, oracle_mod):
    """Mid regime (auction.cu mid_regime): in eps-phases whose bids the hot lists decide, CTA 0 runs the rounds of
    33..t_mid bidders alone.  It must run at the default threshold, and sol / meta / float64 prices must equal the oracle's
    bit for bit at every threshold (off, just above the warp regimes, default, maximum) — float costs, integer costs and
    heavily tied costs (equal bids: the earliest bidder in list order wins, auction_.pyx:379), a rectangular instance
    (holes and push_all_left, :137-162), and iteration caps that end the solve inside a mid-regime round."""
    sslap_b200, nat, h = gpu
    cases = [(4000, 4000, 0.02, "float", 31, False), (3000, 3000, 0.03, "int", 32, False), (2500, 2500, 0.04, "int", 33, True),
             (2000, 2300, 0.03, "float", 34, False), (6000, 6000, 0.01, "float", 35, True)]
    for (n, m, d, mode, seed, tied) in cases:
        loc, val = make_problem(n, d, mode, seed=seed, m=m)
        if tied:
            val = np.round(val / 10.0)
        want = oracle_mod.auction_solve(loc=loc, val=val, problem="max", return_prices=True)
        mids = {}
        for t_mid in (0, 33, 128, 256):
            h.set_option("t_mid", t_mid)
            try:
                got = sslap_b200.auction_solve(loc=loc, val=val, size=(n, m), problem="max", cardinality_check=False,
                                               _raw_meta=True, return_prices=True)
            finally:
                h.set_option("t_mid", 128)
            raw = got["raw"]
            mids[t_mid] = int(raw.rounds_mid)
            assert raw.rounds_grid + raw.rounds_mid + raw.rounds_warp + raw.rounds_solo == raw.its, (n, t_mid)
            assert np.array_equal(got["sol"], want["sol"]), (n, mode, tied, t_mid)
            assert_meta_equal(got["meta"], want["meta"])
            assert np.array_equal(got["prices"], want["prices"]), (n, mode, tied, t_mid)
        assert mids[0] == 0, (n, mids)
        assert mids[128] > 0 or tied or n != m, (n, mids)        # (tied costs rarely let a hot list prove its answer)
    # iteration caps spread over the whole solve: some end inside a mid-regime round (the list goes back to global memory)
    loc, val = make_problem(3000, 0.03, "float", seed=36)
    full = oracle_mod.auction_solve(loc=loc, val=val, problem="min")
    its = int(full["meta"]["its"])
    hit = 0
    for cap in sorted({max(1, its * k // 40) for k in range(1, 40)}):
        want = oracle_mod.auction_solve(loc=loc, val=val, problem="min", max_iter=cap)
        got = sslap_b200.auction_solve(loc=loc, val=val, size=(3000, 3000), problem="min", cardinality_check=False,
                                       max_iter=cap, _raw_meta=True)
        assert np.array_equal(got["sol"], want["sol"]), cap
        assert_meta_equal(got["meta"], want["meta"])
        hit += int(got["raw"].rounds_mid > 0)
    assert hit > 0
    for bad in (1, 32, 257, -3):
        with pytest.raises(ValueError):
            h.set_option("t_mid", bad)


@pytest.mark.parametrize("k_ranks", [2, 4, 8])
def test_row_sharded_solve_with_virtual_ranks_on_one_gpu(gpu, oracle_mod, k_ranks):
    """SURVEY.md §8(e) with K virtual ranks on ONE GPU: K handles, K concurrent persistent kernels (grids of sms/K CTAs so
    that they are co-resident), the same in-kernel exchange (peer stores + system-scope flags, here inside one device).
    Every rank must reproduce the single-GPU trajectory bit for bit: sol, meta, float64 prices; the device's nnz-balanced
    row split must equal the numpy statement."""
    import torch
    sslap_b200, nat, h = gpu
    from sslap_b200 import parallel
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    handles = [nat.Handle(0) for _ in range(k_ranks)]
    try:
        for hh in handles:
            hh.set_option("max_ctas", max(1, (sms - 4) // k_ranks))           # one CTA per SM: all K grids must be co-resident
            hh.set_option("coop", 0)                                         # the driver runs ONE cooperative kernel at a time
            hh.set_option("t_shard", 48)                                     # shard (almost) every grid round
            hh.set_option("watchdog_ms", 20000)
        parallel.connect_local(handles, 8192)
        cases = [(3000, 0.004, "float", 1, 3000), (1000, 0.01, "int", 0, 1000), (700, 0.05, "int", 7, 760),
                 (8000, 0.002, "float", 2, 8000)]
        for (n, d, mode, seed, m) in cases:
            loc, val = make_problem(n, d, mode, seed=seed, m=m)
            if mode == "int" and seed == 7:
                val = np.round(val / 10.0)                                   # heavy ties: the position tie-break pass runs
            want = oracle_mod.auction_solve(loc=loc, val=val, problem="max", return_prices=True, max_iter=200000)
            out, errors = [None] * k_ranks, []

            def work(r):
                try:
                    out[r] = sslap_b200.auction_solve(loc=loc, val=val, size=(n, m), problem="max", max_iter=200000,
                                                      cardinality_check=False, _handle=handles[r], _raw_meta=True,
                                                      return_prices=True)
                except Exception as e:                                       # pragma: no cover
                    errors.append((r, repr(e)))

            threads = [threading.Thread(target=work, args=(r,)) for r in range(k_ranks)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
            assert not errors, "\n".join(f"rank {r}: {e}" for r, e in errors)
            indptr = np.searchsorted(loc[:, 0], np.arange(n + 1))
            split = parallel.balanced_row_split(indptr, k_ranks)
            for r in range(k_ranks):
                g = out[r]
                assert g["raw"].n_ranks == k_ranks and g["raw"].rank == r and g["raw"].rounds_sharded > 0
                assert (g["raw"].row_lo, g["raw"].row_hi) == (int(split[r]), int(split[r + 1]))
                assert np.array_equal(g["sol"], want["sol"]), (n, r)
                assert_meta_equal(g["meta"], want["meta"])
                assert np.array_equal(g["prices"][:want["prices"].size], want["prices"]), (n, r)
    finally:
        for hh in handles:
            hh.close()


def test_hopcroft_device_loop_matches_host_loop_and_scipy(gpu, oracle_mod):
    """The device-resident Hopcroft-Karp phase loop (one cooperative launch, frontier queues) against round 1's host-driven
    loop, the oracle and scipy: cardinality identical on deficient graphs that need many phases and long paths."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import maximum_bipartite_matching
    from conftest import assert_valid_matching
    sslap_b200, nat, h = gpu
    rng = np.random.default_rng(31)
    cases = [(100000, 100000, 300000), (5000, 7000, 9000), (7000, 5000, 30000), (300, 300, 310), (1, 1, 1), (50000, 50000, 60000)]
    # a path graph: ONE augmenting path of maximal length after the greedy pass
    n = 2000
    chain = np.array([[i, i] for i in range(n)] + [[i, i + 1] for i in range(n - 1)], dtype=np.int32)
    chain = chain[np.lexsort((chain[:, 1], chain[:, 0]))]
    graphs = [chain[1:]]                                                    # drop (0,0): forces the long alternating path
    for (n, m, e) in cases:
        key = np.unique(rng.integers(0, n, e).astype(np.int64) * m + rng.integers(0, m, e))
        graphs.append(np.stack([key // m, key % m], -1).astype(np.int32))
    for loc in graphs:
        n, m = int(loc[:, 0].max()) + 1, int(loc[:, 1].max()) + 1
        g = sp.csr_matrix((np.ones(len(loc), dtype=np.int8), (loc[:, 0], loc[:, 1])), shape=(n, m))
        card = int((maximum_bipartite_matching(g, perm_type="column") >= 0).sum())
        sizes = []
        for mode in (0, 1):
            h.set_option("hk_host_loop", mode)
            try:
                r = sslap_b200.hopcroft_solve(loc=loc)
            finally:
                h.set_option("hk_host_loop", 0)
            assert_valid_matching(r, loc) if len(loc) < 50000 else None
            lp, rp = r["left_pairings"], r["right_pairings"]
            assert int((lp >= 0).sum()) == r["size"] == int((rp >= 0).sum())
            sizes.append(r["size"])
        assert sizes == [card, card], (n, m, sizes, card)
        if n <= 100000:
            assert oracle_mod.hopcroft_solve(loc=loc)["size"] == card


def test_small_problems_take_the_single_launch_path(gpu, oracle_mod):
    """csrc/small.cu: N, M <= 256 and <= 12288 entries are built, checked for feasibility and solved by ONE kernel launch
    with everything in shared memory — and must give the oracle's sol / meta / float64 prices bit for bit, for dense and
    COO input, int32 and int64 indices (int64 goes to the general path), rectangular shapes, heavy ties, eps_start, max_iter,
    infeasible inputs (the reference's two ValueErrors) and unsorted input (handed to the general path, which sorts)."""
    sslap_b200, nat, h = gpu
    rng = np.random.default_rng(2025)
    n_small = 0
    h.set_option("small_max_n", 256)                                         # default 128 (where it pays); the kernel holds 256
    for case in range(160):
        n = int(rng.integers(2, 200))
        m = n + int(rng.integers(0, 40)) if rng.random() < 0.3 else n
        density = float(rng.choice([0.05, 0.1, 0.3, 0.7, 1.0]))
        mode = "int" if rng.random() < 0.5 else "float"
        loc, val = make_problem(n, density, mode, seed=5000 + case, m=m)
        if rng.random() < 0.3:
            val = np.round(val / 10.0)
        problem = "min" if rng.random() < 0.5 else "max"
        kw = {}
        r = rng.random()
        if r < 0.15:
            kw["eps_start"] = float(rng.choice([0.5, 3.0, 40.0]))
        elif r < 0.3:
            kw["max_iter"] = int(rng.integers(1, 300))
        want = oracle_mod.auction_solve(loc=loc, val=val, problem=problem, return_prices=True, **kw)
        use_dense = rng.random() < 0.4 and (val >= 0).all()
        if use_dense:
            mat = -np.ones((n, m))
            mat[loc[:, 0], loc[:, 1]] = val
            got = sslap_b200.auction_solve(mat=mat, problem=problem, cardinality_check=bool(case % 2), _raw_meta=True,
                                           return_prices=True, **kw)
        else:
            got = sslap_b200.auction_solve(loc=loc if case % 3 else loc.astype(np.int64), val=val, size=(n, m), problem=problem,
                                           cardinality_check=bool(case % 2), _raw_meta=True, return_prices=True, **kw)
        n_small += int(got["raw"].small_path)
        if len(val) <= 12288 and (use_dense or case % 3):
            assert got["raw"].small_path == 1, (case, n, m, len(val))
        assert np.array_equal(got["sol"], want["sol"]), (case, n, m, density, mode, problem, kw)
        assert_meta_equal(got["meta"], want["meta"])
        assert np.array_equal(got["prices"][:want["prices"].size], want["prices"]), (case, "prices")
    assert n_small >= 80
    h.set_option("small_max_n", 128)
    # the reference's two infeasibility errors, raised from the single-launch path
    mat = -np.ones((3, 3))
    mat[0, 0] = 1; mat[1, 0] = 2; mat[2, 1] = 3; mat[2, 2] = 1
    with pytest.raises(ValueError, match=r"Maximum matching possible only involves 2 out of 3 rows"):
        sslap_b200.auction_solve(mat=mat)
    with pytest.raises(ValueError, match="Fewer than 3 valid values provided for 3 rows"):
        sslap_b200.auction_solve(loc=np.array([[0, 0], [1, 1]], dtype=np.int32), val=np.ones(2), size=(3, 3))
    # unsorted COO: the kernel hands it on, the general path sorts it on the device
    loc, val = make_problem(90, 0.2, "float", seed=9)
    want = oracle_mod.auction_solve(loc=loc, val=val, problem="min")
    order = np.lexsort((np.arange(len(val)), (loc[:, 0] * 7) % 13))            # rows regrouped, order inside a row kept
    got = sslap_b200.auction_solve(loc=loc[order], val=val[order], size=(90, 90), problem="min", _raw_meta=True)
    assert got["raw"].small_path == 0 and np.array_equal(got["sol"], want["sol"])
    # the general path on the same small problems stays reachable
    h.set_option("small_path", 0)
    try:
        got = sslap_b200.auction_solve(loc=loc, val=val, size=(90, 90), problem="min", _raw_meta=True)
    finally:
        h.set_option("small_path", 1)
    assert got["raw"].small_path == 0 and np.array_equal(got["sol"], want["sol"])
