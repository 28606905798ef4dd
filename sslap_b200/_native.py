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

EXPORTS = ("sslapb_create", "sslapb_destroy", "sslapb_last_error", "sslapb_set_option", "sslapb_host_alloc",
           "sslapb_host_free", "sslapb_auction_coo", "sslapb_auction_dense", "sslapb_hopcroft_coo",
           "sslapb_hopcroft_dense", "sslapb_get_prices", "sslapb_bid_sweep", "sslapb_auction_batch")


class Meta(C.Structure):
    """struct sslapb_meta (include/sslap_b200.h)."""
    _fields_ = [("start_eps", C.c_float), ("final_eps", C.c_float), ("target_eps", C.c_float),
                ("eCE", C.c_int32), ("soln_found", C.c_int32), ("its", C.c_int64), ("nreductions", C.c_int64),
                ("n_assigned", C.c_int64), ("obj", C.c_float), ("obj64", C.c_double),
                ("setup_ms", C.c_float), ("solve_ms", C.c_float), ("hk_ms", C.c_float), ("h2d_ms", C.c_float),
                ("cardinality", C.c_int32), ("n_rows", C.c_int32), ("n_cols", C.c_int32), ("nnz", C.c_int64),
                ("rounds_grid", C.c_int64), ("rounds_warp", C.c_int64), ("rounds_solo", C.c_int64),
                ("prof_ms", C.c_float * 8), ("prune_second_pass", C.c_int64),
                ("stop_reason", C.c_int32), ("rounds_cluster", C.c_int32)]


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
    """Process-wide handle per device (LOCAL_RANK picks the device under torchrun)."""
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
