"""CPU emulation of the CUDA kernels' logic (tests/emul) against the golden vectors and the oracle port.
Validates, without a GPU, what the kernels compute: bit-plane math, the tree program (chunks / REF_ACC / fslots /
levels for any chunk size), staging + ordered gather. The real kernels are checked on the GPU in test_gpu_*.py."""
import numpy as np
import pytest

from oracle.oracle import random_tree
from tests.emul.emul import Emulator
from tests.golden_util import load_cases


@pytest.fixture(scope="module")
def emu():
    return Emulator()


@pytest.mark.parametrize("chunk_nodes,inline_nodes", [(1, 0), (3, 1), (8, 3), (8, 0), (1000, 3)])
def test_emulation_matches_golden(emu, chunk_nodes, inline_nodes):
    for c in load_cases():
        rc, got, states, _ = emu.run(c["tree"], c["algo"], c["codes"], c["parent_code"], c["root_override"],
                                     c["fwd_root_ref"], c["leaf_present"], c["block"], chunk_nodes=chunk_nodes,
                                     inline_nodes=inline_nodes, level_mode=chunk_nodes % 2)
        assert rc == 0
        assert got.same_as(c["expect"]), f"golden case {c['id']} chunk_nodes={chunk_nodes}"
        assert np.array_equal(states, c["states"]), f"golden case {c['id']} states"


@pytest.mark.parametrize("algo", [0, 1])
def test_emulation_vs_port_random(emu, port, algo):
    rng = np.random.default_rng(11 + algo)
    for trial in range(40):
        kind = ["binary", "polytomy", "unary", "caterpillar"][trial % 4]
        tree = random_tree(int(rng.integers(1, 120)), 2000 + trial, kind, max_arity=[3, 6, 20, 300][(trial // 4) % 4])
        n_cols = int(rng.choice([1, 5, 32, 100, 1023, 1024, 1025, 2500]))
        block = int(trial % 5 == 4)
        nst = 3 if block else 16
        base = rng.integers(0, min(nst, 5), size=n_cols)
        codes = np.repeat(base[None, :], tree.n_leaves, 0)
        noise = rng.random(codes.shape) < [0.01, 0.1, 0.6][trial % 3]
        codes = np.where(noise, rng.integers(0, nst, size=codes.shape), codes).astype(np.uint8)
        pc = rng.integers(0, nst, size=n_cols).astype(np.uint8)
        ro = np.where(rng.random(n_cols) < 0.3, rng.integers(0, nst, size=n_cols), -1).astype(np.int8) if trial % 2 else None
        fr = np.where(rng.random(n_cols) < 0.3, rng.integers(0, nst, size=n_cols), -1).astype(np.int8) if (
            algo == 0 and not block and trial % 3 == 0) else None
        lp = None
        if trial % 4 == 1 and tree.n_leaves > 1:
            lp = (rng.random(tree.n_leaves) < 0.7).astype(np.uint8)
            lp[0] = 1
        want, want_states = port.run(tree, algo, codes, pc, ro, fr, lp, block, n_threads=2, want_states=True)
        rc, got, states, stats = emu.run(tree, algo, codes, pc, ro, fr, lp, block, chunk_nodes=int(rng.choice([1, 2, 5, 16, 64])),
                                         col_base=7, inline_nodes=int(rng.choice([0, 1, 3, 10])), level_mode=trial % 2)
        assert rc == 0
        want.pos = want.pos + 7
        assert got.same_as(want), (algo, trial, kind, stats)
        assert np.array_equal(states, want_states), (algo, trial)


def test_emulation_sankoff_root_undefined(emu):
    # every leaf omitted and no override: the reference would assert (fitchSankoff.cpp:505) -> PMB_ERR_SANKOFF_ROOT
    tree = random_tree(6, 1, "binary")
    codes = np.zeros((6, 10), np.uint8)
    rc, *_ = emu.run(tree, 1, codes, np.zeros(10, np.uint8), None, None, np.zeros(6, np.uint8))
    assert rc == -4
    rc, got, _, _ = emu.run(tree, 1, codes, np.zeros(10, np.uint8), np.full(10, 2, np.int8), None, np.zeros(6, np.uint8))
    assert rc == 0 and got.node_offsets[-1] == 10  # only the forced root differs from the '-' consensus


def test_speculation_bounds_selftest(emu):
    """plane_math.h FitchInterval / fitch_candidates_step against plain integer arithmetic, column by column."""
    import ctypes as C

    emu.L.emul_speculation_selftest.argtypes = [C.c_ulonglong, C.c_int]
    assert emu.L.emul_speculation_selftest(2024, 50000) == 0


@pytest.mark.parametrize("noise", [0.0, 0.02, 0.7])
def test_emulation_chain_segments(emu, port, noise):
    """Deep trees cut into chain segments that are evaluated speculatively (tree_program.h): conserved columns resolve
    after a couple of ops, noisy ones never do and fall back to waiting -- both must equal the oracle."""
    rng = np.random.default_rng(int(noise * 100) + 3)
    for trial in range(12):
        kind = ["caterpillar", "unary", "polytomy"][trial % 3]
        tree = random_tree(int(rng.integers(40, 260)), 4000 + trial, kind, max_arity=3)
        n_cols = int(rng.choice([40, 1024, 1500]))
        base = rng.integers(0, 5, size=n_cols)
        codes = np.repeat(base[None, :], tree.n_leaves, 0)
        codes = np.where(rng.random(codes.shape) < noise, rng.integers(0, 16, size=codes.shape), codes).astype(np.uint8)
        pc = rng.integers(0, 16, size=n_cols).astype(np.uint8)
        ro = np.where(rng.random(n_cols) < 0.3, rng.integers(0, 16, size=n_cols), -1).astype(np.int8) if trial % 2 else None
        fr = np.where(rng.random(n_cols) < 0.3, rng.integers(0, 16, size=n_cols), -1).astype(np.int8) if trial % 4 == 0 else None
        for algo in (0, 1):
            want, want_states = port.run(tree, algo, codes, pc, ro, fr if algo == 0 else None, None, 0, n_threads=2, want_states=True)
            rc, got, states, stats = emu.run(tree, algo, codes, pc, ro, fr if algo == 0 else None, None, 0,
                                             chunk_nodes=int(rng.choice([1, 2, 4, 9])), inline_nodes=int(rng.choice([0, 1, 3])),
                                             level_mode=trial % 2)
            assert rc == 0, (rc, trial, algo)
            assert got.same_as(want) and np.array_equal(states, want_states), (trial, algo, kind)
            if kind == "caterpillar":
                assert stats[4] > 0, "the tree was expected to be cut into chain segments"
                if noise == 0.0:  # both passes speculate (Fitch and Sankoff) and conserved columns always resolve
                    assert stats[5] > 0 and stats[6] == stats[5]


def test_run_merge_logic_matches_oracle(emu, port):
    """The block-decomposed logic of the record-parallel run-merge kernels (pieces start where (index - run start) % 6 == 0;
    the latest break is carried across blocks by an exclusive max-scan, the piece numbers by a prefix sum; a piece's length
    is read off the flags ahead of it) against the oracle's greedy merge (reference src/panman.cpp:1445-1466), incl. column
    breaks checked against the oracle's PanGraph gap-list rule (:1261). Tiny blocks make runs straddle many of them."""
    import ctypes as C

    rng = np.random.default_rng(17)
    f = emu.L.emul_merge_runs
    f.restype = C.c_longlong
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    for trial in range(300):
        block = C.c_longlong([3, 8, 64, 2048][trial % 4])
        n = int(rng.integers(1, 200))
        # positions: long consecutive stretches with occasional jumps; types change now and then
        step = np.where(rng.random(n) < [0.05, 0.3, 0.8][trial % 3], rng.integers(2, 9, size=n), 1)
        pos = np.cumsum(step).astype(np.int32) + int(rng.integers(0, 50))
        ty = np.cumsum(rng.random(n) < 0.1) % 3
        tc = ((ty << 4) | rng.integers(0, 16, size=n)).astype(np.uint8)
        out_p, out_i, out_n = np.empty(n, np.int32), np.empty(n, np.uint8), np.empty(n, np.uint32)
        k = f(C.c_longlong(n), P(pos, C.c_int32), P(tc, C.c_uint8), None, C.c_longlong(0), P(out_p, C.c_int32), P(out_i, C.c_uint8),
              P(out_n, C.c_uint32), block)
        wp, wi, wn = port.merge_msa(pos, tc)
        assert k == len(wp) and np.array_equal(out_p[:k], wp) and np.array_equal(out_i[:k], wi) and np.array_equal(out_n[:k], wn), trial
        # gap-slot columns: column c = (gap position j, slot k); consecutive columns merge only inside one position
        widths = rng.integers(1, 12, size=40)
        col_j = np.repeat(np.arange(len(widths)), widths)
        col_k = np.concatenate([np.arange(w) for w in widths])
        brk = (col_k == 0).astype(np.uint8)
        cols = np.sort(rng.choice(len(col_j), size=min(n, len(col_j)), replace=False)).astype(np.int32)
        tcc = tc[:len(cols)]
        out_p, out_i, out_n = np.empty(len(cols), np.int32), np.empty(len(cols), np.uint8), np.empty(len(cols), np.uint32)
        k = f(C.c_longlong(len(cols)), P(cols, C.c_int32), P(tcc, C.c_uint8), P(brk, C.c_uint8), C.c_longlong(0), P(out_p, C.c_int32),
              P(out_i, C.c_uint8), P(out_n, C.c_uint32), block)
        ob, op_, og, mi, nu = port.merge_pangraph(1, np.zeros(len(cols), np.int32), col_j[cols], col_k[cols], tcc)
        assert k == len(op_) and np.array_equal(col_j[out_p[:k]], op_) and np.array_equal(col_k[out_p[:k]], og), trial
        assert np.array_equal(out_i[:k], mi) and np.array_equal(out_n[:k], nu), trial
