"""hopcroft_solve — same surface as the reference's (/root/reference/sslap/check_feasible.py:5-20,
feasibility_.pyx:227-283); the matching itself runs on the GPU (csrc/hopcroft.cu) through the C ABI.
"""
import ctypes as C

import numpy as np

from . import _native as nat


def hopcroft_solve(loc: np.ndarray = None, mat: np.ndarray = None, lookup: dict = None,
                   _handle: nat.Handle = None) -> dict:
    """Maximum matching of a bipartite graph given as ``loc`` (E x 2 integer edges), ``mat`` (2-D, entry >= 0 = edge)
    or ``lookup`` ({i: [j, ...]}).  Returns ``{'size', 'left_pairings' int32[|I|], 'right_pairings' int32[|J|]}`` with
    -1 for unmatched vertices.  ``size`` equals the reference's; the pairings are a valid maximum matching (maximum
    matchings are not unique, so they need not be the reference's)."""
    n_none = (loc is None) + (mat is None) + (lookup is None)
    assert n_none == 2, "Exactly one of the arguments loc, mat, lookup must be provided."   # feasibility_.pyx:232-233
    h = _handle or nat.default_handle()
    L = nat.load()
    size = C.c_int32(0)
    if mat is not None:                                   # feasibility_.pyx:249-266
        mat = np.ascontiguousarray(np.asarray(mat, dtype=np.float64))
        if mat.ndim != 2:
            raise ValueError("mat must be 2-D")
        n, m = mat.shape
        left = np.empty(n, dtype=np.int32)
        right = np.empty(m, dtype=np.int32)
        rc = L.sslapb_hopcroft_dense(h.ptr, mat.ctypes.data, n, m, nat.MEM_HOST, left.ctypes.data, right.ctypes.data,
                                     C.byref(size))
    else:
        if lookup is not None:                            # feasibility_.pyx:268-279
            edges = [(i, j) for i in lookup for j in lookup[i]]
            loc = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
            loc = loc[np.argsort(loc[:, 0], kind="stable")]
        loc = np.asarray(loc)
        if loc.ndim != 2 or loc.shape[1] != 2 or not np.issubdtype(loc.dtype, np.integer):
            raise ValueError("loc must be an integer array of shape (E, 2)")
        if loc.dtype not in (np.int32, np.int64):
            loc = loc.astype(np.int64)
        loc = np.ascontiguousarray(loc)
        if loc.shape[0] == 0:
            raise ValueError("empty graph")
        n, m = int(loc[:, 0].max()) + 1, int(loc[:, 1].max()) + 1   # feasibility_.pyx:245-246
        left = np.empty(n, dtype=np.int32)
        right = np.empty(m, dtype=np.int32)
        ib = loc.dtype.itemsize
        rc = L.sslapb_hopcroft_coo(h.ptr, loc.ctypes.data, loc.ctypes.data + ib, ib, 2, loc.shape[0], n, m,
                                   nat.MEM_HOST, left.ctypes.data, right.ctypes.data, C.byref(size))
    if rc == nat.E_UNSORTED:
        raise ValueError("loc must be sorted by row (precondition of the reference, feasibility_.pyx:22-46).")
    if rc == nat.E_OUT_OF_RANGE:
        raise ValueError("loc holds a negative index.")
    nat.check(h, rc, "hopcroft_solve")
    if rc != 0:
        raise RuntimeError(f"hopcroft_solve failed with code {rc}: {h.last_error()}")
    return dict(size=int(size.value), left_pairings=left, right_pairings=right)
