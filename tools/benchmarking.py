"""Size / density sweeps in the style of the reference's benchmarking.py (/root/reference/benchmarking.py:29-147), with
the validity checks that script computes but never asserts: complete assignment, only admissible entries used, and the
objective compared with scipy's optimum.  No plots (matplotlib is not in the image): prints one table per sweep.

    python tools/benchmarking.py [--max-size 10000] [--reference]

Inputs follow the reference: seeded dense matrix (np.random.seed(1), uniform(0,100) or randint(1,100)), sparsified with
np.random.seed(2) at the requested density, every row/column kept feasible, invalid entries = -1, problem='max',
dense `mat` input (so the timing includes the device-side dense -> CSR build and the Hopcroft-Karp check).
"""
import argparse
import os
import sys
import time

import numpy as np
from scipy.optimize import linear_sum_assignment

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

seed = 1
problem = 'max'


def _improve_feasibility(mask):
    """benchmarking.py:17-26 — every row and column keeps at least one admissible entry."""
    R, C = mask.shape
    for r in np.nonzero((~mask).sum(axis=1) == 0)[0]:
        mask[r, np.random.randint(C)] = False
    for c in np.nonzero((~mask).sum(axis=0) == 0)[0]:
        mask[np.random.randint(R), c] = False
    return mask


def make_matrix(size, density=1.0, mode='float'):
    """benchmarking.py:29-45."""
    np.random.seed(seed)
    if mode == 'int':
        mat = np.random.randint(1, 100, (size, size)).astype(np.float64)
    else:
        mat = np.random.uniform(0., 100., size=(size, size)).astype(np.float64)
    np.random.seed(seed + 1)
    mask = _improve_feasibility(np.random.random(mat.shape) > density)
    mat[mask] = -1
    return mat


def check(mat, sol, mode):
    """benchmarking.py:56-64, asserted; returns the objective."""
    n = mat.shape[0]
    assert np.unique(sol).size == n and (sol >= 0).all() and (sol < n).all(), "incomplete assignment"
    sel = mat[np.arange(n), sol]
    assert (sel >= 0).all(), "assignment uses an inadmissible entry"
    return float(sel.sum())


def scipy_optimum(mat):
    big = -1e6
    w = np.where(mat >= 0, mat, big)
    r, c = linear_sum_assignment(w, maximize=True)
    assert (mat[r, c] >= 0).all(), "scipy had to use an inadmissible entry (instance infeasible)"
    return float(mat[r, c].sum())


def timeit(fn, reps=3):
    best = float("inf")
    out = None
    for _ in range(reps):
        t = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t)
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-size", type=int, default=10000)
    ap.add_argument("--reference", action="store_true", help="also time the unmodified reference (oracle/_ref) when present")
    args = ap.parse_args()
    import sslap_b200
    ref = None
    if args.reference:
        from oracle import ref_loader
        if ref_loader.available():
            ref = ref_loader.load()

    def run(mat, mode):
        n = mat.shape[0]
        tg, res = timeit(lambda: sslap_b200.auction_solve(mat, problem=problem))
        obj = check(mat, res['sol'], mode)
        ts, best = timeit(lambda: scipy_optimum(mat), reps=1)
        if mode == 'int':
            assert obj == best, (obj, best)
        else:
            assert abs(obj - best) <= n * (1.0 / n + 1e-7) + 1e-6, (obj, best)
        tr = None
        if ref is not None:
            tr, rr = timeit(lambda: ref.auction_solve(mat.copy(), problem=problem), reps=1)
            assert np.array_equal(rr['sol'], res['sol']), "sol differs from the reference"
        return tg, ts, tr, obj, best, res['meta']['its']

    print("size sweep, density 100 % (figs/size_benchmarking.png)")
    print(f"{'N':>6} {'mode':>5} {'gpu ms':>9} {'scipy ms':>9} {'ref ms':>9} {'rounds':>8}  objective (= scipy optimum)")
    sizes = [int(round(10 ** e)) for e in np.arange(1.0, np.log10(args.max_size) + 1e-9, 0.5)]
    for n in sizes:
        for mode in ('float', 'int'):
            mat = make_matrix(n, 1.0, mode)
            tg, ts, tr, obj, best, its = run(mat, mode)
            print(f"{n:6d} {mode:>5} {tg*1e3:9.2f} {ts*1e3:9.2f} {(tr*1e3 if tr else float('nan')):9.2f} {its:8d}  {obj:.6f}", flush=True)
    print("density sweep, N = 1000 (figs/density_benchmarking.png)")
    for dens in (0.01, 0.02, 0.05, 0.1, 0.2, 0.5, 1.0):
        mat = make_matrix(1000, dens, 'float')
        tg, ts, tr, obj, best, its = run(mat, 'float')
        print(f"{dens:6.2f} float {tg*1e3:9.2f} {ts*1e3:9.2f} {(tr*1e3 if tr else float('nan')):9.2f} {its:8d}  {obj:.6f}", flush=True)
    print("all completeness / validity / objective checks passed")


if __name__ == "__main__":
    main()
