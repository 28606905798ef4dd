"""ctypes binding of the C ABI in include/sslap_b200.h (csrc/libsslap_b200.so, sm_100a CUDA).

There is no CPU fallback: importing the solvers without the built library, or calling them without a CUDA device,
raises.  Build with ``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C sslap_b200/csrc``.
"""
import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SSLAP_B200_LIB", os.path.join(_HERE, "csrc", "libsslap_b200.so"))

OK, E_FEWER_THAN_N, E_CARDINALITY, E_UNSORTED, E_BAD_ARG, E_OUT_OF_RANGE, E_EMPTY_ROW, E_ABORTED = range(8)
MEM_HOST, MEM_DEVICE_IN, MEM_DEVICE_OUT = 0, 1, 2

ABI_VERSION = 4
COMM_EXPORT_BYTES, COMM_MAX_RANKS = 128, 8

# every symbol include/sslap_b200.h declares (tests/test_abi.py checks this list against the header)
EXPORTS = ("sslapb_create", "sslapb_destroy", "sslapb_last_error", "sslapb_meta_size", "sslapb_abi_version",
           "sslapb_set_option", "sslapb_host_alloc", "sslapb_host_free", "sslapb_auction_coo", "sslapb_auction_dense",
           "sslapb_hopcroft_coo", "sslapb_hopcroft_dense", "sslapb_auction_batch", "sslapb_get_prices",
           "sslapb_set_prices", "sslapb_comm_init", "sslapb_comm_connect", "sslapb_comm_destroy", "sslapb_bid_sweep")


class Meta(C.Structure):
    """struct sslapb_meta (include/sslap_b200.h)."""
    _fields_ = [("start_eps", C.c_float), ("final_eps", C.c_float), ("target_eps", C.c_float),
                ("eCE", C.c_int32), ("soln_found", C.c_int32), ("its", C.c_int64), ("nreductions", C.c_int64),
                ("n_assigned", C.c_int64), ("obj", C.c_float), ("obj64", C.c_double),
                ("setup_ms", C.c_float), ("solve_ms", C.c_float), ("hk_ms", C.c_float), ("h2d_ms", C.c_float),
                ("cardinality", C.c_int32), ("n_rows", C.c_int32), ("n_cols", C.c_int32), ("nnz", C.c_int64),
                ("rounds_grid", C.c_int64), ("rounds_warp", C.c_int64), ("rounds_solo", C.c_int64),
                ("prof_ms", C.c_float * 8), ("prune_second_pass", C.c_int64),
                ("stop_reason", C.c_int32), ("small_path", C.c_int32),
                ("n_ranks", C.c_int32), ("rank", C.c_int32), ("row_lo", C.c_int32), ("row_hi", C.c_int32),
                ("rounds_sharded", C.c_int64), ("xchg_ms", C.c_float), ("sharded_ms", C.c_float),
                ("sweep_insitu_us", C.c_float), ("sweep_insitu_n", C.c_int32), ("warm_start", C.c_int32),
                ("strict", C.c_int32),
                ("hot_grid_bids", C.c_int64), ("hot_grid_fallbacks", C.c_int64),
                ("hot_tail_rounds", C.c_int64), ("hot_tail_fallbacks", C.c_int64), ("rounds_nohole", C.c_int64),
                ("rounds_mid", C.c_int64)]


_lib = None
_lock = threading.Lock()
_handles = {}


def load():
    """Load libsslap_b200.so and declare the prototypes.  Raises ImportError when the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: sslap_b200 has no CPU fallback — build the CUDA library first "
                          f"(`make -C sslap_b200/csrc` or `__graft_entry__.build()`)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    L.sslapb_create.restype = C.c_int
    L.sslapb_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.sslapb_destroy.restype = None
    L.sslapb_destroy.argtypes = [vp]
    L.sslapb_last_error.restype = C.c_char_p
    L.sslapb_last_error.argtypes = [vp]
    L.sslapb_set_option.restype = C.c_int
    L.sslapb_set_option.argtypes = [vp, C.c_char_p, i64]
    L.sslapb_host_alloc.restype = vp
    L.sslapb_host_alloc.argtypes = [C.c_size_t]
    L.sslapb_host_free.restype = None
    L.sslapb_host_free.argtypes = [vp]
    L.sslapb_auction_coo.restype = C.c_int
    L.sslapb_auction_coo.argtypes = [vp, vp, vp, C.c_int, i64, vp, i64, i32, i32, C.c_int, f32, i64, C.c_int, C.c_int,
                                     vp, C.POINTER(Meta)]
    L.sslapb_auction_dense.restype = C.c_int
    L.sslapb_auction_dense.argtypes = [vp, vp, i32, i32, C.c_int, f32, i64, C.c_int, C.c_int, vp, C.POINTER(Meta)]
    L.sslapb_hopcroft_coo.restype = C.c_int
    L.sslapb_hopcroft_coo.argtypes = [vp, vp, vp, C.c_int, i64, i64, i32, i32, C.c_int, vp, vp, C.POINTER(i32)]
    L.sslapb_hopcroft_dense.restype = C.c_int
    L.sslapb_hopcroft_dense.argtypes = [vp, vp, i32, i32, C.c_int, vp, vp, C.POINTER(i32)]
    L.sslapb_auction_batch.restype = C.c_int
    L.sslapb_auction_batch.argtypes = [vp, i32, vp, vp, vp, vp, vp, C.c_int, i64, vp, C.c_int, vp, i64, C.c_int, vp, vp]
    L.sslapb_get_prices.restype = C.c_int
    L.sslapb_get_prices.argtypes = [vp, vp]
    L.sslapb_bid_sweep.restype = C.c_int
    L.sslapb_bid_sweep.argtypes = [vp, vp, vp, i32, f32, C.c_int, C.c_int, C.c_int, vp, vp, C.POINTER(f32)]
    L.sslapb_meta_size.restype = C.c_size_t
    L.sslapb_meta_size.argtypes = []
    L.sslapb_abi_version.restype = C.c_int
    L.sslapb_abi_version.argtypes = []
    L.sslapb_set_prices.restype = C.c_int
    L.sslapb_set_prices.argtypes = [vp, vp, i32]
    L.sslapb_comm_init.restype = C.c_int
    L.sslapb_comm_init.argtypes = [vp, C.c_int, C.c_int, i64, vp]
    L.sslapb_comm_connect.restype = C.c_int
    L.sslapb_comm_connect.argtypes = [vp, vp]
    L.sslapb_comm_destroy.restype = C.c_int
    L.sslapb_comm_destroy.argtypes = [vp]
    # layout guard: a binding written against another header would be overrun by the library's memset of the meta block
    if L.sslapb_abi_version() != ABI_VERSION or L.sslapb_meta_size() != C.sizeof(Meta):
        raise ImportError(f"{LIB_PATH}: ABI {L.sslapb_abi_version()} / sizeof(sslapb_meta) {L.sslapb_meta_size()} does not "
                          f"match this binding (ABI {ABI_VERSION}, {C.sizeof(Meta)} bytes): rebuild the library")
    _lib = L
    return L


class Handle:
    """One device + stream + grow-only HBM scratch (sslapb_handle)."""

    def __init__(self, device: int = 0):
        L = load()
        self._h = C.c_void_p()
        rc = L.sslapb_create(int(device), C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"sslapb_create(device={device}) failed with code {rc}: a CUDA device with cooperative "
                               f"launch support is required (there is no CPU fallback)")
        self.device = device

    @property
    def ptr(self):
        return self._h

    def last_error(self) -> str:
        return load().sslapb_last_error(self._h).decode()

    def set_option(self, name: str, value: int):
        rc = load().sslapb_set_option(self._h, name.encode(), int(value))
        if rc != 0:
            raise ValueError(f"sslapb_set_option({name}, {value}) -> {rc}: {self.last_error()}")

    def set_prices(self, prices):
        """Warm start: the next auction call on this handle starts from `prices` (float64, one per column; None clears)."""
        import numpy as np
        if prices is None:
            rc = load().sslapb_set_prices(self._h, None, 0)
        else:
            prices = np.ascontiguousarray(prices, dtype=np.float64)
            rc = load().sslapb_set_prices(self._h, prices.ctypes.data, int(prices.shape[0]))
        if rc != 0:
            raise RuntimeError(f"sslapb_set_prices -> {rc}: {self.last_error()}")

    def get_prices(self, n_cols: int):
        import numpy as np
        p = np.empty(int(n_cols), dtype=np.float64)
        rc = load().sslapb_get_prices(self._h, p.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"sslapb_get_prices -> {rc}: {self.last_error()}")
        return p

    def comm_init(self, n_ranks: int, rank: int, capacity_rows: int) -> bytes:
        """Create this rank's exchange buffer; returns the opaque export blob to hand to every rank's comm_connect."""
        buf = C.create_string_buffer(COMM_EXPORT_BYTES)
        rc = load().sslapb_comm_init(self._h, int(n_ranks), int(rank), int(capacity_rows), buf)
        if rc != 0:
            raise RuntimeError(f"sslapb_comm_init -> {rc}: {self.last_error()}")
        return buf.raw

    def comm_connect(self, exports):
        """`exports`: the blobs of ALL ranks in rank order."""
        blob = b"".join(exports)
        rc = load().sslapb_comm_connect(self._h, blob)
        if rc != 0:
            raise RuntimeError(f"sslapb_comm_connect -> {rc}: {self.last_error()}")

    def comm_destroy(self):
        load().sslapb_comm_destroy(self._h)

    def close(self):
        if self._h:
            load().sslapb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_handle(device: int = None) -> Handle:
    """Process-wide handle per device (LOCAL_RANK picks the device under torchrun).  Shared by all Python threads: the
    library serialises concurrent calls on one handle with a per-handle mutex (ctypes releases the GIL), so threads are
    safe but take turns; give each thread its own `Handle` to overlap their solves."""
    if device is None:
        device = int(os.environ.get("SSLAP_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    with _lock:
        h = _handles.get(device)
        if h is None:
            h = Handle(device)
            _handles[device] = h
        return h


def check(h: Handle, rc: int, what: str):
    """Map a negative (CUDA) or unexpected return code to an exception."""
    if rc < 0:
        raise RuntimeError(f"{what}: CUDA error {-rc}: {h.last_error()}")
    if rc in (E_BAD_ARG, E_ABORTED):
        raise RuntimeError(f"{what}: {h.last_error()}")
