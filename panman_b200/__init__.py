"""panman_b200 -- B200-native Fitch/Sankoff construction pass of PanMAN behind a C ABI.

The product is panman_b200/libpanman_b200.so (CUDA, sm_100a; sources in panman_b200/csrc, interface in
include/panman_b200.h). This Python package is a thin ctypes mirror of that interface for tests and the benchmark,
plus the synthetic workload generator named by BASELINE.json. There is no CPU compute path here.
"""
from .api import (ALGO_FITCH, ALGO_SANKOFF, FLAG_BLOCK_MODE, FLAG_WANT_STATES, Context, Group, PanmanError, Result, Runs, Timings,
                  column_range, pack_nibbles)
from .lib import LIB_PATH, build_library, load_library

__all__ = ["ALGO_FITCH", "ALGO_SANKOFF", "FLAG_BLOCK_MODE", "FLAG_WANT_STATES", "Context", "Group", "column_range", "PanmanError", "Result", "Runs", "Timings",
           "pack_nibbles", "LIB_PATH", "build_library", "load_library"]
