"""The clade-run encoding of the leaf matrix at the boundary (pmb_runs_*, include/panman_b200.h).

Host half (no device): pmb_runs_encode followed by a numpy restatement of what expand_runs_kernel does with the events must
give the matrix back, whatever the tree shape, width or content. Device half (-m gpu): every entry that takes the encoding
returns bit for bit what the nibble-matrix entry returns and what the oracle says -- lists, states, on one context, on a
column range, over a group."""
import numpy as np
import pytest

import panman_b200 as pb
from oracle.oracle import random_tree
from tests.golden_util import merge_all_nodes


def _matrix(rng, tree, n_cols, noise, nst=16):
    base = rng.integers(0, min(nst, 5), size=n_cols)
    codes = np.repeat(base[None, :], tree.n_leaves, 0)
    codes = np.where(rng.random(codes.shape) < noise, rng.integers(0, nst, size=codes.shape), codes).astype(np.uint8)
    return codes, base.astype(np.uint8)


def _same(res, want):
    return (np.array_equal(res.node_offsets, want.node_offsets) and np.array_equal(res.pos, want.pos)
            and np.array_equal(res.type_code, want.type_code))


@pytest.mark.parametrize("kind", ["binary", "polytomy", "unary", "caterpillar"])
def test_encode_round_trip(kind):
    rng = np.random.default_rng(hash(kind) % 1000)
    for trial in range(6):
        tree = random_tree(int(rng.integers(1, 700)), 40 + trial, kind, max_arity=[3, 30][trial % 2])
        n_cols = int(rng.choice([1, 2, 31, 1023, 1024, 1025, 2049, 4100]))
        codes, pc = _matrix(rng, tree, n_cols, [0.0, 0.02, 1.0][trial % 3])
        if trial == 4:
            pc = rng.integers(0, 16, size=n_cols).astype(np.uint8)  # a parent code unrelated to the rows is legal
        c4 = pb.pack_nibbles(codes)
        stride = c4.shape[1] + 3  # rows need not lie back to back
        wide = np.full((tree.n_leaves, stride), 0xEE, np.uint8)
        wide[:, :c4.shape[1]] = c4
        if n_cols % 2:
            wide[:, c4.shape[1] - 1] |= 0xF0  # the unused high nibble of an odd width is ignored
        r = pb.Runs(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row, n_cols, wide, stride, pc,
                    n_threads=[1, 3][trial % 2])
        I = r.info
        assert (I.n_cols, I.n_rows, I.n_tiles) == (n_cols, tree.n_leaves, (n_cols + 1023) // 1024)
        assert I.n_segments * I.seg_rows >= tree.n_leaves > (I.n_segments - 1) * I.seg_rows
        assert np.array_equal(r.decode(pc), codes), (kind, trial)
        if trial % 3 == 0:  # every leaf equals the parent code: nothing to say
            assert r.n_events == 0
        r.close()


@pytest.mark.parametrize("kind", ["binary", "polytomy", "caterpillar"])
def test_expansion_builds_the_planes_of_the_matrix_path(kind):
    """expand_runs_kernel restated lane by lane (tests/emul) on the encoder's events against pack_leaves_kernel's planes, with the
    leaf slots of real tree programs (several chunk sizes): the two ingest paths hand the pass kernels the same leaf matrix."""
    from tests.emul.emul import expand_runs_mismatches

    rng = np.random.default_rng(11)
    for trial in range(5):
        tree = random_tree(int(rng.integers(2, 500)), 300 + trial, kind, max_arity=[3, 12][trial % 2])
        n_cols = int(rng.choice([5, 1000, 1024, 1025, 2500]))
        codes, pc = _matrix(rng, tree, n_cols, [0.01, 0.3, 1.0][trial % 3])
        runs = pb.Runs.of_tree(tree, n_cols, pb.pack_nibbles(codes), pc)
        for chunk_nodes, inline_nodes in ((8, 3), (1, 0), (64, 2)):
            assert expand_runs_mismatches(tree, codes, pc, runs, chunk_nodes, inline_nodes) == 0, (kind, trial, chunk_nodes)
        runs.close()


def test_encode_is_small_on_clade_structured_columns():
    """A substitution on a branch changes a whole clade = one run of consecutive leaves in depth-first order: two events."""
    from panman_b200 import synth

    tree = synth.make_tree(2000, 9, "binary")
    codes4, pc = synth.simulate_msa(tree, 0, 4096, synth.MsaSpec(9, 1e-4, 0.05, 1e-4), device="cpu")
    r = pb.Runs.of_tree(tree, 4096, codes4.numpy(), pc.numpy())
    assert np.array_equal(r.decode(pc.numpy()), synth.unpack_nibbles(codes4, 4096).numpy())
    assert r.nbytes * 20 < codes4.numel()  # measured: about 1/60 of the nibble matrix
    r.close()


def test_encode_rejects_bad_arguments():
    tree = random_tree(20, 1, "binary")
    codes = np.zeros((tree.n_leaves, 8), np.uint8)
    pc = np.zeros(16, np.uint8)
    with pytest.raises(pb.PanmanError):  # stride smaller than the row
        pb.Runs(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row, 16, codes, 7, pc)
    lr = tree.leaf_row.copy()
    lr[lr >= 0] = 0  # every leaf claims row 0
    with pytest.raises(pb.PanmanError):
        pb.Runs(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, lr, 16, codes, 8, pc)
    with pytest.raises(pb.PanmanError):
        pb.Runs(tree.n_nodes, tree.n_nodes + 5, tree.child_off, tree.child_idx, tree.leaf_row, 16, codes, 8, pc)


# ------------------------------------------------------------------------------------------------------------ device half
@pytest.mark.gpu
@pytest.mark.parametrize("algo", [0, 1])
def test_runs_entry_matches_matrix_entry_and_oracle(port, algo):
    ctx = pb.Context(0)
    rng = np.random.default_rng(300 + algo)
    for trial in range(16):
        kind = ["binary", "polytomy", "unary", "caterpillar"][trial % 4]
        tree = random_tree(int(rng.integers(1, 900)), 5000 + trial, kind, max_arity=[3, 6, 40][trial % 3])
        n_cols = int(rng.choice([1, 33, 1024, 1025, 3000, 9000]))
        block = int(trial % 5 == 4)
        nst = 3 if block else 16
        codes, pc = _matrix(rng, tree, n_cols, [0.01, 0.1, 0.7][trial % 3], nst)
        ro = np.where(rng.random(n_cols) < 0.3, rng.integers(0, nst, size=n_cols), -1).astype(np.int8) if trial % 2 else None
        fr = np.where(rng.random(n_cols) < 0.3, rng.integers(0, nst, size=n_cols), -1).astype(np.int8) if (
            algo == 0 and not block and trial % 3 == 0) else None
        lp = None
        if trial % 4 == 1 and tree.n_leaves > 1:
            lp = (rng.random(tree.n_leaves) < 0.7).astype(np.uint8)
            lp[0] = 1
        want, want_states = port.run(tree, algo, codes, pc, ro, fr, lp, block, n_threads=4, want_states=True)
        ctx.set_option("chunk_nodes", int(rng.choice([0, 1, 7, 64])))
        ctx.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
        c4 = pb.pack_nibbles(codes)
        runs = pb.Runs.of_tree(tree, n_cols, c4, pc)
        flags = pb.FLAG_WANT_STATES | (pb.FLAG_BLOCK_MODE if block else 0)
        res = ctx.run_runs(algo, runs, pc, ro, fr, lp, col_base=0, flags=flags)
        assert _same(res, want), (trial, "run_runs vs oracle")
        assert np.array_equal(res.states, want_states), (trial, "states")
        dense = ctx.run_nuc(algo, n_cols, tree.n_leaves, c4, c4.shape[1], pc, ro, fr, lp, 0, flags)
        assert _same(res, dense) and np.array_equal(res.states, dense.states), (trial, "run_runs vs run_nuc")
        if n_cols > 2048:  # a tile-aligned column range of the encoded batch, positions offset by col_base
            a, b = 1024, 1024 * (n_cols // 1024) if trial % 2 else n_cols
            ctx.upload_runs(runs, pc[a:b], None if ro is None else ro[a:b], None if fr is None else fr[a:b], lp, col_begin=a,
                            n_cols=b - a, col_base=a)
            ctx.run_resident(algo, flags & pb.FLAG_BLOCK_MODE)
            part = ctx.download()
            w2, _ = port.run(tree, algo, codes[:, a:b], pc[a:b], None if ro is None else ro[a:b], None if fr is None else fr[a:b], lp,
                             block, n_threads=4)
            assert np.array_equal(part.node_offsets, w2.node_offsets) and np.array_equal(part.pos, w2.pos + a) \
                and np.array_equal(part.type_code, w2.type_code), (trial, "column range")
        runs.close()
    ctx.close()


@pytest.mark.gpu
def test_runs_errors():
    ctx = pb.Context(0)
    rng = np.random.default_rng(5)
    tree = random_tree(50, 1, "binary")
    codes, pc = _matrix(rng, tree, 3000, 0.05)
    c4 = pb.pack_nibbles(codes)
    runs = pb.Runs.of_tree(tree, 3000, c4, pc)
    # the encoding belongs to a depth-first leaf order (any tree that walks the rows in the same order may use it): the same
    # topology with the rows numbered backwards is another one
    ctx.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, np.where(tree.leaf_row >= 0, tree.n_leaves - 1 - tree.leaf_row, -1))
    with pytest.raises(pb.PanmanError) as e:
        ctx.upload_runs(runs, pc)
    assert e.value.code == -1
    ctx.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
    for a, n in ((512, 1024), (0, 1000), (2048, 2000), (-1024, 1024)):  # ranges must follow the tiles
        with pytest.raises(pb.PanmanError):
            ctx.upload_runs(runs, pc[:max(n, 1)], col_begin=a, n_cols=n)
    ctx.upload_runs(runs, pc)  # and the context is still usable
    ctx.run_resident(0)
    assert ctx.download().n_mut > 0
    runs.close()
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [[0], [0, 0, 0], "all"])
def test_runs_over_a_group(port, devices):
    import torch

    if devices == "all":
        if torch.cuda.device_count() < 2:
            pytest.skip("one GPU visible")
        devices = list(range(torch.cuda.device_count()))
    rng = np.random.default_rng(77)
    tree = random_tree(400, 12, "binary")
    n_cols = 1024 * 2 * len(devices) + 700
    codes, pc = _matrix(rng, tree, n_cols, 0.03)
    c4 = pb.pack_nibbles(codes)
    g = pb.Group(devices)
    g.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
    runs = pb.Runs.of_tree(tree, n_cols, c4, pc)
    for algo in (0, 1):
        ro = codes[0].astype(np.int8) if algo == 1 else None
        want, _ = port.run(tree, algo, codes, pc, ro, None, None, 0, n_threads=8)
        assert _same(g.run_runs(algo, runs, pc, ro), want), (devices, algo, "whole batch")
        got = g.merge_runs()
        for x, y in zip(got, merge_all_nodes(port, want.node_offsets, want.pos, want.type_code)):
            assert np.array_equal(x, y), (devices, algo, "NucMut fields")
        # every rank encodes and uploads its own range (what one process per GPU does)
        shards = []
        for i in range(len(devices)):
            a, b = g.column_range(n_cols, i)
            sc4 = pb.pack_nibbles(codes[:, a:b])
            shards.append(pb.Runs.of_tree(tree, b - a, sc4, pc[a:b]))
            g.upload_shard_runs(i, n_cols, shards[-1], pc[a:b], None if ro is None else ro[a:b])
        for _ in range(3):
            g.run_async(algo)
        g.wait()
        assert _same(g.download(), want), (devices, algo, "per-rank shards")
        for s in shards:
            s.close()
    runs.close()
    g.close()
