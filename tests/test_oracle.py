"""CPU tests: the C restatement (oracle/) against the reference's golden vectors, known answers and scipy."""
import zlib

import numpy as np
import pytest
import scipy.sparse as sp
from scipy.sparse.csgraph import maximum_bipartite_matching, min_weight_full_bipartite_matching

from conftest import (assert_meta_equal, assert_valid_matching, dense_golden_names, hopcroft_golden_names, load_golden,
                      sparse_golden_names)
from sslap_b200.datagen import make_problem, objective


def test_known_answers_from_reference_examples(oracle_mod):
    # /root/reference/examples/test_auction.py:7-30 and test_feasibility.py:12-13 (values recorded in SURVEY.md §4)
    g = load_golden("example_dense")
    r = oracle_mod.auction_solve(mat=g["mat"], problem="min")
    assert r["sol"].tolist() == [0, 1, 4, 3, 2]
    assert (r["meta"]["obj"], r["meta"]["start_eps"], r["meta"]["its"], r["meta"]["nreductions"], r["meta"]["final_eps"]) == \
        (10.845, 4.841, 10, 2, 0.109)
    g = load_golden("example_sparse")
    r = oracle_mod.auction_solve(mat=g["mat"], problem="max")
    assert r["sol"].tolist() == [0, 3, 4, 2, 1] and r["meta"]["obj"] == 23.681 and r["meta"]["its"] == 15
    r = oracle_mod.hopcroft_solve(lookup={0: [0, 1], 1: [1, 2], 2: [1, 4], 3: [2], 4: [3]})
    assert r["size"] == 5 and r["left_pairings"].tolist() == [0, 1, 4, 2, 3] and r["right_pairings"].tolist() == [0, 1, 3, 4, 2]


@pytest.mark.parametrize("name", sparse_golden_names())
@pytest.mark.parametrize("faithful", [False, True])
def test_oracle_matches_reference_sparse(oracle_mod, name, faithful):
    g = load_golden(name)
    r = oracle_mod.auction_solve(loc=g["loc"], val=g["val"], problem=g["problem"], faithful_scan=faithful, **g["kwargs"])
    assert np.array_equal(r["sol"], g["sol"])          # bit-exact, ties included
    assert_meta_equal(r["meta"], g["meta"])


@pytest.mark.parametrize("name", dense_golden_names())
def test_oracle_matches_reference_dense(oracle_mod, name):
    g = load_golden(name)
    r = oracle_mod.auction_solve(mat=g["mat"], problem=g["problem"])
    assert np.array_equal(r["sol"], g["sol"])
    assert_meta_equal(r["meta"], g["meta"])


def test_oracle_matches_reference_c2(oracle_mod):
    g = load_golden("c2_float_min")
    loc, val = make_problem(int(g["n"]), float(g["density"]), g["mode"], seed=int(g["seed"]))
    assert zlib.crc32(loc.tobytes()) == int(g["loc_crc"][0]) and zlib.crc32(val.tobytes()) == int(g["val_crc"][0])
    r = oracle_mod.auction_solve(loc=loc, val=val, problem=g["problem"])
    assert np.array_equal(r["sol"], g["sol"])
    assert_meta_equal(r["meta"], g["meta"])
    assert r["meta"]["its"] == 21113


@pytest.mark.parametrize("name", hopcroft_golden_names())
def test_oracle_hopcroft_matches_reference(oracle_mod, name):
    g = load_golden(name)
    loc = g["loc"]
    if name == "example_hopcroft":
        r = oracle_mod.hopcroft_solve(loc=loc)
    else:
        r = oracle_mod.hopcroft_solve(loc=loc)
    assert r["size"] == int(g["size"])
    assert np.array_equal(r["left_pairings"], g["left"]) and np.array_equal(r["right_pairings"], g["right"])
    assert_valid_matching(r, loc)


@pytest.mark.parametrize("n,density,mode,problem", [(200, 0.05, "int", "min"), (300, 0.03, "float", "max"),
                                                    (1000, 0.01, "int", "min")])
def test_oracle_reaches_scipy_optimum(oracle_mod, n, density, mode, problem):
    loc, val = make_problem(n, density, mode, seed=17)
    r = oracle_mod.auction_solve(loc=loc, val=val, problem=problem)
    assert r["meta"]["soln_found"] == 1 and sorted(r["sol"].tolist()) == list(range(n))
    w = sp.csr_matrix((val if problem == "min" else -val + 1000.0, (loc[:, 0], loc[:, 1])), shape=(n, n))
    rows, cols = min_weight_full_bipartite_matching(w)
    best = float(val[np.searchsorted(loc[:, 0].astype(np.int64) * n + loc[:, 1], rows.astype(np.int64) * n + cols)].sum())
    got = objective(loc, val, r["sol"])
    if mode == "int":
        assert got == best
    else:
        assert abs(got - best) <= n * r["meta"]["_raw"]["target_eps"] + 1e-6


def test_oracle_hopcroft_cardinality_vs_scipy(oracle_mod):
    rng = np.random.default_rng(5)
    for n, e in [(100, 150), (1000, 1800), (5000, 9000)]:
        key = np.unique(rng.integers(0, n, e).astype(np.int64) * n + rng.integers(0, n, e))
        loc = np.stack([key // n, key % n], -1).astype(np.int32)
        r = oracle_mod.hopcroft_solve(loc=loc, N=n, M=n)
        g = sp.csr_matrix((np.ones(len(loc)), (loc[:, 0], loc[:, 1])), shape=(n, n))
        assert r["size"] == int((maximum_bipartite_matching(g, perm_type="column") >= 0).sum())


def test_oracle_live_reference_random(oracle_mod):
    """When the reference build (oracle/_ref) is present, compare on fresh random inputs, ties and all."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref not built")
    ref = ref_loader.load()
    for seed in range(6):
        n = 50 + 37 * seed
        loc, val = make_problem(n, 0.1, "int" if seed % 2 else "float", seed=seed)
        problem = "min" if seed % 3 else "max"
        want = ref.auction_solve(loc=loc, val=val.copy(), size=(n, n), problem=problem, cardinality_check=False)
        got = oracle_mod.auction_solve(loc=loc, val=val, problem=problem)
        assert np.array_equal(want["sol"], got["sol"])
        assert_meta_equal(got["meta"], want["meta"])
        assert ref.hopcroft_solve(loc=loc)["size"] == oracle_mod.hopcroft_solve(loc=loc)["size"]
