"""TEST INFRASTRUCTURE ONLY — loader for the unmodified reference (sslap v0.2.5) installed under oracle/_ref/.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg / --impl reference) may import this module.
The product path (sslap_b200/) never does.

The reference's module init touches the long-removed aliases ``np.float`` / ``np.int``
(/root/reference/sslap/auction_.pyx:18,25,28 and feasibility_.pyx:13,17), so a 2-line runtime shim is applied
before import; no reference source is edited (see oracle/build_ref.sh for the install recipe).
"""
import os
import sys
import warnings

_REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(_REF_DIR, "sslap"))


def load():
    """Return the reference ``sslap`` module (auction_solve, hopcroft_solve)."""
    if not available():
        raise RuntimeError("oracle/_ref is missing: run oracle/build_ref.sh (needs /root/reference)")
    import numpy as np
    if not hasattr(np, "float"):
        np.float = np.float64  # noqa: shim for auction_.pyx:18
    if not hasattr(np, "int"):
        np.int = np.int_       # noqa: shim for auction_.pyx:25 (Windows branch symbol lookup only)
    if _REF_DIR not in sys.path:
        sys.path.insert(0, _REF_DIR)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import sslap  # the reference package, NOT sslap_b200
    return sslap
