"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/sslap_b200.h declares,
and the Python front mirrors the reference's argument handling.  No compute calls (no GPU here)."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest

import sslap_b200
from sslap_b200 import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sslap_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sslapb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(nat.LIB_PATH), "build the CUDA library first: make -C sslap_b200/csrc"
    lib = ctypes.CDLL(nat.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/sslap_b200.h but not exported"
    assert set(nat.EXPORTS) == set(syms)


def test_meta_struct_layout_matches_header():
    # field order/types of struct sslapb_meta as declared in the header
    text = open(os.path.join(ROOT, "include", "sslap_b200.h")).read()
    body = re.search(r"typedef struct sslapb_meta \{(.*?)\} sslapb_meta;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            names += [re.sub(r"\[\d+\]", "", n.strip()) for n in decl.split(None, 1)[1].split(",")]
    assert names == [f[0] for f in nat.Meta._fields_]
    assert ctypes.sizeof(nat.Meta) % 8 == 0


def test_integration_stub_matches_the_binding():
    """The ctypes stub INTEGRATION.md shows a maintainer of the reference must describe the SAME struct as the header and
    the package's own binding (round 1 shipped a stub that was one field short: the library would have overrun it)."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"class Meta\(C\.Structure\):.*?_fields_ = \[(.*?)\]\n", text, flags=re.S).group(1)
    fields = re.findall(r'\("([A-Za-z0-9_]+)",\s*C\.(c_[a-z0-9]+)(?:\s*\*\s*(\d+))?\)', block)
    got = [(name, getattr(ctypes, ct) * int(n) if n else getattr(ctypes, ct)) for name, ct, n in fields]
    assert [g[0] for g in got] == [f[0] for f in nat.Meta._fields_]
    for (name, ct), (_, want) in zip(got, nat.Meta._fields_):
        assert ctypes.sizeof(ct) == ctypes.sizeof(want) and getattr(ct, "_type_", ct) == getattr(want, "_type_", want), name
    assert "sslapb_meta_size" in text and "sslapb_abi_version" in text
    lib = ctypes.CDLL(nat.LIB_PATH)
    lib.sslapb_meta_size.restype = ctypes.c_size_t
    assert lib.sslapb_meta_size() == ctypes.sizeof(nat.Meta)
    assert lib.sslapb_abi_version() == nat.ABI_VERSION
    n_syms = int(re.search(r"\((\d+) `extern \"C\"` symbols", text).group(1))
    assert n_syms == len(declared_symbols()) == len(nat.EXPORTS)
    design = open(os.path.join(ROOT, "DESIGN.md")).read()
    assert f"{n_syms} `extern \"C\"` symbols" in design


def test_every_option_is_documented():
    """Every name sslapb_set_option accepts (csrc/api.cu) is described in the header and in INTEGRATION.md."""
    api = open(os.path.join(ROOT, "sslap_b200", "csrc", "api.cu")).read()
    body = api[api.index('extern "C" int sslapb_set_option'):]
    body = body[:body.index("\n}\n")]
    names = re.findall(r'!strcmp\(name, "([a-z0-9_]+)"\)', body)
    assert len(names) >= 10 and "t_mid" in names and "t_small" in names
    header = open(os.path.join(ROOT, "include", "sslap_b200.h")).read()
    integration = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for n in names:
        assert f'"{n}"' in header, f"option {n} is not described in include/sslap_b200.h"
        assert f"`{n}`" in integration, f"option {n} is not described in INTEGRATION.md"


def test_signatures_mirror_the_reference():
    # /root/reference/sslap/auction_solve.py:6-8 and check_feasible.py:5
    p = inspect.signature(sslap_b200.auction_solve).parameters
    assert list(p)[:10] == ["mat", "loc", "val", "coo_mat", "problem", "eps_start", "max_iter", "fast", "size",
                            "cardinality_check"]
    assert (p["problem"].default, p["eps_start"].default, p["max_iter"].default, p["fast"].default,
            p["size"].default, p["cardinality_check"].default) == ("min", 0.0, 1000000, False, None, True)
    q = inspect.signature(sslap_b200.hopcroft_solve).parameters
    assert list(q)[:3] == ["loc", "mat", "lookup"]


def test_argument_errors_raise_before_touching_the_gpu(monkeypatch):
    class FakeHandle:
        ptr = None
    monkeypatch.setattr(nat, "default_handle", lambda device=None: FakeHandle())
    with pytest.raises(ValueError, match="One of the following formats is expected"):
        sslap_b200.auction_solve()                                   # auction_solve.py:51-52
    with pytest.raises(ValueError, match="Buffer dtype mismatch"):
        sslap_b200.auction_solve(loc=np.zeros((3, 2), dtype=np.int32), val=np.ones(3, dtype=np.float32))
    with pytest.raises(AssertionError, match="Exactly one of the arguments"):
        sslap_b200.hopcroft_solve()                                   # feasibility_.pyx:232-233
    with pytest.raises(AssertionError):
        sslap_b200.hopcroft_solve(loc=np.zeros((1, 2), dtype=np.int32), mat=np.zeros((1, 1)))


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nat.Handle(0)
    src = ""
    for f in os.listdir(os.path.join(ROOT, "sslap_b200")):
        if f.endswith(".py"):
            src += open(os.path.join(ROOT, "sslap_b200", f)).read()
    assert "oracle" not in src.replace("cardinality oracle", ""), "the product package must not reference oracle/"
