import ast
import glob
import os
import sys

import numpy as np
import pytest

# The virtual-rank test runs up to 8 persistent kernels side by side on ONE device, each on its own stream; with the default
# of 8 hardware work queues two of those streams can share a queue, and the second kernel then waits for the first to
# finish — which it never does.  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
META_KEYS = ("start_eps", "eCE", "its", "nreductions", "soln_found", "n_assigned", "obj", "final_eps")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    for k in ("problem", "mode", "kwargs"):
        if k in d:
            d[k] = str(d[k])
    if "kwargs" in d:
        d["kwargs"] = ast.literal_eval(d["kwargs"])
    if "meta" in d:
        d["meta"] = dict(zip(META_KEYS, d["meta"].tolist()))
    return d


def sparse_golden_names():
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))):
        z = np.load(p)
        if "loc" in z.files and "val" in z.files and "sol" in z.files:
            out.append(os.path.basename(p)[:-4])
    return out


def dense_golden_names():
    return [os.path.basename(p)[:-4] for p in sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))) if "mat" in np.load(p).files]


def hopcroft_golden_names():
    return [os.path.basename(p)[:-4] for p in sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))
            if os.path.basename(p).startswith(("hopcroft_", "example_hopcroft"))]


def assert_meta_equal(got, want, keys=META_KEYS):
    for k in keys:
        assert float(got[k]) == float(want[k]), f"meta[{k}]: {got[k]} != {want[k]}"


def assert_valid_matching(res, loc, n_left=None, n_right=None):
    """left/right pairings consistent, every matched pair is an edge, size = number of matched left vertices."""
    lp, rp = res["left_pairings"], res["right_pairings"]
    edges = set(map(tuple, np.asarray(loc).tolist()))
    m = 0
    for u, v in enumerate(lp.tolist()):
        if v >= 0:
            assert rp[v] == u
            assert (u, v) in edges
            m += 1
    assert m == res["size"]
    assert int((rp >= 0).sum()) == m


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.lib()
    return oracle
