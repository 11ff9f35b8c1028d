"""Structural invariants of the tree program (panman_b200/csrc/tree_program.cpp) over random trees and parameters:
the chunks partition the ops, every reference names a child, external rows / chain children / parent states come from
earlier tickets in both passes and from lower levels, parked parent states are the right ones (stack replay), chain
segments are flagged consistently. Checked by tests/emul emul_prog_check; hypothesis drives the shapes."""
import ctypes as C

import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from tests.emul.emul import Emulator

_emu = Emulator()
_emu.L.emul_prog_check.restype = C.c_int


def _tree_from_parents(parents):
    """parents[i] < i for node i + 1: any rooted tree on len(parents) + 1 nodes; returns CSR arrays with node 0 the root."""
    n = len(parents) + 1
    kids = [[] for _ in range(n)]
    for i, p in enumerate(parents):
        kids[p].append(i + 1)
    off = np.zeros(n + 1, np.int32)
    idx = []
    for v in range(n):
        idx += kids[v]
        off[v + 1] = len(idx)
    leaf_row = np.full(n, -1, np.int32)
    r = 0
    for v in range(n):
        if not kids[v]:
            leaf_row[v] = r
            r += 1
    return n, off, np.asarray(idx, np.int32), leaf_row


def _check(parents, chunk_nodes, inline_nodes, tail):
    n, off, idx, leaf_row = _tree_from_parents(parents)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    return _emu.L.emul_prog_check(C.c_int(n), C.c_int(0), p(off), p(idx), p(leaf_row), C.c_int(chunk_nodes), C.c_int(inline_nodes),
                                  C.c_int(tail))


@st.composite
def trees(draw):
    n = draw(st.integers(2, 400))
    shape = draw(st.sampled_from(["uniform", "deep", "bushy"]))
    parents = []
    for i in range(1, n):
        if shape == "uniform":
            parents.append(draw(st.integers(0, i - 1)))
        elif shape == "deep":  # long chains with occasional branches: chain segments
            parents.append(i - 1 if draw(st.integers(0, 9)) else draw(st.integers(0, i - 1)))
        else:
            parents.append(draw(st.integers(0, min(i - 1, 3))))
    return parents


@settings(max_examples=300, deadline=None)
@given(trees(), st.sampled_from([1, 2, 3, 5, 16, 64, 1000]), st.sampled_from([0, 1, 3, 9]), st.sampled_from([0, 3, 50]))
def test_tree_program_invariants(parents, chunk_nodes, inline_nodes, tail):
    assert _check(parents, chunk_nodes, inline_nodes, tail) == 0


def test_caterpillar_and_star_shapes():
    for n in (2, 3, 50, 3000):
        assert _check(list(range(n - 1)), 4, 1, 5) in (0,)             # a pure chain of unary nodes ending in one leaf
        assert _check([0] * (n - 1), 4, 1, 5) == 0                      # a star
        spine = []                                                      # a caterpillar: spine node i has leaf 2i+1... built as parents
        for i in range(1, n):
            spine.append(i - 1 if i % 2 else max(0, i - 2))
        assert _check(spine, 4, 0, 5) == 0


# ---------------------------------------------------------------- emulation of the kernels' logic on the same random shapes
def _flat_tree(parents):
    from oracle.oracle import FlatTree

    n = len(parents) + 1
    kids = [[] for _ in range(n)]
    for i, p in enumerate(parents):
        kids[p].append(i + 1)
    names, k = [], 0
    for v in range(n):
        if kids[v]:
            k += 1
            names.append(f"node_{k}")
        else:
            names.append(f"L{v}")
    return FlatTree.from_children(names, kids, 0)


@settings(max_examples=60, deadline=None)
@given(trees(), st.sampled_from([1, 2, 4, 9, 40]), st.sampled_from([0, 1, 3]), st.integers(0, 1), st.sampled_from([0.0, 0.03, 0.5]),
       st.integers(0, 2 ** 31 - 1))
def test_emulation_matches_port_on_random_shapes(parents, chunk_nodes, inline_nodes, algo, noise, seed):
    """Whatever the tree shape (unary chains, stars, deep backbones cut into chain segments) and however noisy the columns,
    the emulated kernels -- speculation included, which asserts its own exactness -- produce the oracle's lists and states."""
    from oracle.oracle import PortOracle

    tree = _flat_tree(parents)
    if tree.n_leaves < 1 or tree.n_nodes - tree.n_leaves < 1:
        return
    rng = np.random.default_rng(seed)
    n_cols = int(rng.choice([7, 64, 1030]))
    base = rng.integers(0, 5, size=n_cols)
    codes = np.repeat(base[None, :], tree.n_leaves, 0)
    codes = np.where(rng.random(codes.shape) < noise, rng.integers(0, 16, size=codes.shape), codes).astype(np.uint8)
    pc = rng.integers(0, 16, size=n_cols).astype(np.uint8)
    ro = np.where(rng.random(n_cols) < 0.3, rng.integers(0, 16, size=n_cols), -1).astype(np.int8) if seed % 2 else None
    want, want_states = PortOracle().run(tree, algo, codes, pc, ro, None, None, 0, n_threads=1, want_states=True)
    rc, got, states, _ = _emu.run(tree, algo, codes, pc, ro, None, None, 0, chunk_nodes=chunk_nodes, inline_nodes=inline_nodes,
                                  level_mode=seed % 3 == 0)
    assert rc == 0
    assert got.same_as(want) and np.array_equal(states, want_states)
