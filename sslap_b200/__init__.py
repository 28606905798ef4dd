"""sslap_b200 — B200-native (sm_100a CUDA) drop-in for the hot path of OllieBoyne/sslap v0.2.5.

Same two names as the reference package (/root/reference/sslap/__init__.py:1-4).  Importing the package is cheap;
the CUDA library is loaded on first use and there is no CPU fallback.
"""
from .auction_solve import auction_solve
from .check_feasible import hopcroft_solve
from .batch import auction_solve_batch, pack_problems

__version__ = "0.1.0"
__all__ = ["auction_solve", "hopcroft_solve", "auction_solve_batch", "pack_problems"]
