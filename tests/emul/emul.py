"""TEST INFRASTRUCTURE ONLY -- builds and binds tests/emul/emul_kernels.cpp (CPU emulation of the CUDA kernels'
logic, sharing plane_math.h / tree_program.cpp with the product; see the header of that file)."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle.oracle import MutLists

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libemul.so")
SRCS = [os.path.join(HERE, "emul_kernels.cpp"), os.path.join(ROOT, "panman_b200", "csrc", "tree_program.cpp")]
DEPS = SRCS + [os.path.join(ROOT, "panman_b200", "csrc", h) for h in ("plane_math.h", "tree_program.h")]


def build():
    if os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in DEPS):
        return
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SO] + SRCS)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class Emulator:
    def __init__(self):
        build()
        self.L = C.CDLL(SO)
        self.L.emul_run.restype = C.c_longlong

    def run(self, tree, algo, leaf_codes, parent_code, root_override=None, fwd_root_ref=None, leaf_present=None,
            block_mode=0, chunk_nodes=8, col_base=0, want_states=True, inline_nodes=3, level_mode=0):
        codes = np.ascontiguousarray(leaf_codes, np.uint8)
        n_rows, n_cols = codes.shape
        pc = np.ascontiguousarray(parent_code, np.uint8)
        ro = None if root_override is None else np.ascontiguousarray(root_override, np.int8)
        fr = None if fwd_root_ref is None else np.ascontiguousarray(fwd_root_ref, np.int8)
        lp = None if leaf_present is None else np.ascontiguousarray(leaf_present, np.uint8)
        off = np.zeros(tree.n_nodes + 1, np.int64)
        stats = np.zeros(8, np.int32)

        def call(pos, tc, states):
            return self.L.emul_run(
                C.c_int(algo), C.c_int(block_mode), C.c_int(tree.n_nodes), C.c_int(tree.root), _p(tree.child_off, C.c_int32),
                _p(tree.child_idx, C.c_int32), _p(tree.leaf_row, C.c_int32), C.c_int(chunk_nodes), C.c_longlong(n_cols),
                _p(codes, C.c_uint8), _p(lp, C.c_uint8), _p(pc, C.c_uint8), _p(ro, C.c_int8), _p(fr, C.c_int8),
                C.c_longlong(col_base), _p(off, C.c_longlong), _p(pos, C.c_int32), _p(tc, C.c_uint8), _p(states, C.c_uint8),
                _p(stats, C.c_int32), C.c_int(inline_nodes), C.c_int(level_mode))

        n = call(None, None, None)
        if n < 0:
            return int(n), None, None, stats
        pos = np.empty(max(n, 1), np.int32)
        tc = np.empty(max(n, 1), np.uint8)
        states = np.empty((tree.n_nodes, n_cols), np.uint8) if want_states else None
        n = call(pos, tc, states)
        return 0, MutLists(off.copy(), pos[:n].copy(), tc[:n].copy()), states, stats


def expand_runs_mismatches(tree, codes, parent_code, runs, chunk_nodes=8, inline_nodes=3):
    """Plane words in which the emulated expand_runs_kernel (on `runs`, a panman_b200.Runs of `codes`) differs from the emulated
    pack_leaves_kernel: 0 = both ingest paths build the same leaf matrix for this tree program."""
    build()
    L = C.CDLL(SO)
    L.emul_expand_runs.restype = C.c_longlong
    codes = np.ascontiguousarray(codes, np.uint8)
    pc = np.ascontiguousarray(parent_code, np.uint8)
    ev = np.ascontiguousarray(runs.events(), np.uint32) if runs.n_events else np.zeros(1, np.uint32)
    off = np.ascontiguousarray(runs.item_offsets(), np.int64)
    return int(L.emul_expand_runs(C.c_int(tree.n_nodes), C.c_int(tree.root), _p(tree.child_off, C.c_int32), _p(tree.child_idx, C.c_int32),
                                  _p(tree.leaf_row, C.c_int32), C.c_int(chunk_nodes), C.c_int(inline_nodes), C.c_longlong(codes.shape[1]),
                                  _p(codes, C.c_uint8), _p(pc, C.c_uint8), _p(ev, C.c_uint32), _p(off, C.c_longlong),
                                  C.c_int(runs.info.n_segments), C.c_int(runs.info.seg_rows)))
