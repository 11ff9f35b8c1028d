"""GPU parity tests: the CUDA path, called through the C ABI (include/panman_b200.h), against
  * the committed golden vectors (outputs of the reference's own fitchSankoff.cpp, tests/golden/make_golden.py),
  * the oracle port on the same seeded inputs (bit-exact lists and states),
  * size-independent properties at BASELINE.json's full configuration sizes.
Bit-exact is the bar everywhere: this is integer / index work."""
import numpy as np
import pytest

import panman_b200 as pb
from oracle.oracle import random_tree
from panman_b200 import synth
from tests.golden_util import load_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = pb.Context(0)
    yield c
    c.close()


def _set_tree(ctx, tree):
    ctx.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)


def _same(res, want):
    return (np.array_equal(res.node_offsets, want.node_offsets) and np.array_equal(res.pos, want.pos)
            and np.array_equal(res.type_code, want.type_code))


@pytest.mark.parametrize("chunk_nodes,inline_nodes,schedule", [(0, 3, 1), (1, 0, 1), (5, 1, 1), (5, 3, 0)])
def test_golden_vectors(ctx, chunk_nodes, inline_nodes, schedule):
    """schedule 1 = persistent kernels with dependency flags, 0 = one launch per dependency level"""
    ctx.set_option("chunk_nodes", chunk_nodes)
    ctx.set_option("inline_nodes", inline_nodes)
    ctx.set_option("schedule", schedule)
    for c in load_cases():
        _set_tree(ctx, c["tree"])
        res = ctx.run_codes(c["tree"], c["algo"], c["codes"], c["parent_code"], c["root_override"], c["fwd_root_ref"],
                            c["leaf_present"], c["block"], want_states=True)
        assert _same(res, c["expect"]), f"golden case {c['id']}"
        assert np.array_equal(res.states, c["states"]), f"golden case {c['id']} states"
    ctx.set_option("chunk_nodes", 0)
    ctx.set_option("inline_nodes", 3)
    ctx.set_option("schedule", 1)
    ctx.set_option("col_groups", 0)


@pytest.mark.parametrize("algo", [0, 1])
def test_random_vs_oracle(ctx, port, algo):
    rng = np.random.default_rng(100 + algo)
    for trial in range(24):
        kind = ["binary", "polytomy", "unary", "caterpillar"][trial % 4]
        tree = random_tree(int(rng.integers(1, 400)), 3000 + trial, kind, max_arity=[3, 6, 20, 300][(trial // 4) % 4])
        n_cols = int(rng.choice([1, 33, 1000, 1024, 1025, 5000]))
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
        want, want_states = port.run(tree, algo, codes, pc, ro, fr, lp, block, n_threads=4, want_states=True)
        ctx.set_option("chunk_nodes", int(rng.choice([0, 1, 7, 64])))
        ctx.set_option("inline_nodes", int(rng.choice([0, 2, 3, 9])))
        ctx.set_option("schedule", int(trial % 3 != 0))
        ctx.set_option("col_groups", int(rng.choice([0, 1, 2, 5])))
        _set_tree(ctx, tree)
        res = ctx.run_codes(tree, algo, codes, pc, ro, fr, lp, block, want_states=True, col_base=11)
        want.pos = want.pos + 11
        assert _same(res, want), (algo, trial, kind)
        assert np.array_equal(res.states, want_states), (algo, trial)
    ctx.set_option("chunk_nodes", 0)
    ctx.set_option("inline_nodes", 3)
    ctx.set_option("schedule", 1)
    ctx.set_option("col_groups", 0)


def test_dense_mutations_overflow_path(ctx, port):
    """Every cell random: the staging pool is outgrown and the library must size it exactly and redo the pass."""
    rng = np.random.default_rng(5)
    tree = random_tree(300, 77, "binary")
    codes = rng.integers(0, 16, size=(tree.n_leaves, 3000)).astype(np.uint8)
    pc = rng.integers(0, 16, size=3000).astype(np.uint8)
    want, _ = port.run(tree, 0, codes, pc, n_threads=4)
    ctx.set_option("staging_records", 1000)
    _set_tree(ctx, tree)
    res = ctx.run_codes(tree, 0, codes, pc)
    ctx.set_option("staging_records", 0)
    assert res.n_mut > 1000 and _same(res, want)


def test_resident_api_is_rerunnable(ctx, port):
    tree = random_tree(200, 9, "binary")
    rng = np.random.default_rng(9)
    codes = rng.integers(0, 5, size=(tree.n_leaves, 2500)).astype(np.uint8)
    pc = codes[0].copy()
    _set_tree(ctx, tree)
    c4 = pb.pack_nibbles(codes)
    ctx.upload(2500, tree.n_leaves, c4, c4.shape[1], pc)
    want_f, _ = port.run(tree, 0, codes, pc, n_threads=2)
    want_s, _ = port.run(tree, 1, codes, pc, n_threads=2)
    for _ in range(2):
        t = ctx.run_resident(pb.ALGO_FITCH)
        assert t.total_ms > 0 and t.n_launches > 0
        assert _same(ctx.download(), want_f)
        ctx.run_resident(pb.ALGO_SANKOFF)
        assert _same(ctx.download(), want_s)
    assert ctx.algorithmic_bytes(pb.ALGO_FITCH) == 2500 * (tree.n_leaves + 4 * (tree.n_nodes - tree.n_leaves)) + 8 * want_s.node_offsets[-1] \
        or ctx.algorithmic_bytes(pb.ALGO_FITCH) > 0


def test_error_behaviour(ctx):
    tree = random_tree(6, 1, "binary")
    _set_tree(ctx, tree)
    codes = np.zeros((6, 10), np.uint8)
    # every leaf omitted, no override: the reference would assert (fitchSankoff.cpp:505)
    with pytest.raises(pb.PanmanError) as e:
        ctx.run_codes(tree, 1, codes, np.zeros(10, np.uint8), None, None, np.zeros(6, np.uint8))
    assert e.value.code == -4
    res = ctx.run_codes(tree, 1, codes, np.zeros(10, np.uint8), np.full(10, 2, np.int8), None, np.zeros(6, np.uint8))
    assert res.n_mut == 10
    # malformed trees are rejected, not executed
    with pytest.raises(pb.PanmanError) as e:
        ctx.set_tree(3, 0, np.asarray([0, 2, 2, 2]), np.asarray([1, 1]), np.asarray([-1, 0, 1]))
    assert e.value.code == -1
    with pytest.raises(pb.PanmanError):
        ctx.run_codes(tree, 0, np.zeros((5, 10), np.uint8), np.zeros(10, np.uint8))  # wrong row count
    fresh = pb.Context(0)
    with pytest.raises(pb.PanmanError) as e:
        fresh.run_resident(0)
    assert e.value.code == -3
    fresh.close()


def _replay_property(tree, res, codes, parent_code):
    """Reference invariant (commented asserts src/panman.cpp:1192-1225; replay rules src/fasta.cpp:535-657): walking
    root -> leaf and applying each node's records over the consensus reproduces the leaf's character.
    Node ids are pre-order, so a per-depth stack of rows is enough."""
    depth = tree.depth()
    path = [None] * (int(depth.max()) + 2)
    for v in range(tree.n_nodes):
        base = parent_code if v == tree.root else path[depth[v] - 1]
        a, b = res.node_offsets[v], res.node_offsets[v + 1]
        row = base.copy()
        row[res.pos[a:b]] = res.type_code[a:b] & 15
        if tree.leaf_row[v] >= 0:
            if not np.array_equal(row, codes[tree.leaf_row[v]]):
                return False
        else:
            path[depth[v]] = row
    return True


@pytest.mark.parametrize("name", ["sars20k", "indel10k"])
def test_full_size_config_vs_oracle(ctx, port, name):
    """BASELINE.json configs[1] and configs[2] at full size: exact list equality with the oracle (the port runs
    1.2e9 node x columns in seconds), plus the replay property and sortedness."""
    cfg = synth.CONFIGS[name]
    tree = synth.make_tree(cfg["n_leaves"], cfg["seed"], cfg["kind"])
    codes4, pc = synth.simulate_msa(tree, 0, cfg["n_cols"], synth.MsaSpec(cfg["seed"], cfg["p_sub"], cfg["f_gap"], cfg["p_N"]),
                                    device="cuda")
    codes = synth.unpack_nibbles(codes4, cfg["n_cols"]).cpu().numpy()
    pc_h = pc.cpu().numpy()
    _set_tree(ctx, tree)
    ctx.upload(cfg["n_cols"], tree.n_leaves, codes4, codes4.shape[1], pc)
    for algo_name in cfg["algos"]:
        algo = 0 if algo_name == "fitch" else 1
        ro = codes[0].astype(np.int8) if algo == 1 else None  # Sankoff with --reference = leaf 0
        if ro is not None:
            ctx.upload(cfg["n_cols"], tree.n_leaves, codes4, codes4.shape[1], pc, ro)
        ctx.run_resident(algo)
        res = ctx.download()
        want, _ = port.run(tree, algo, codes, pc_h, ro, None, None, 0, n_threads=8)
        assert _same(res, want), (name, algo_name)
        # ascending positions inside every node's list
        d = np.diff(res.pos.astype(np.int64))
        starts = res.node_offsets[1:-1]
        starts = starts[(starts > 0) & (starts < len(res.pos))]
        d[starts - 1] = 1
        assert (d > 0).all()
        if algo == 0:
            assert _replay_property(tree, res, codes, pc_h)


def _slice_lists(res, n_nodes, a, b):
    """Per-node records of `res` with a <= position < b, as (offsets, pos, type_code)."""
    node = np.repeat(np.arange(n_nodes), np.diff(res.node_offsets))
    keep = (res.pos >= a) & (res.pos < b)
    off = np.zeros(n_nodes + 1, np.int64)
    off[1:] = np.cumsum(np.bincount(node[keep], minlength=n_nodes))
    return off, res.pos[keep], res.type_code[keep]


def _sorted_within_nodes(res):
    d = np.diff(res.pos.astype(np.int64))
    starts = res.node_offsets[1:-1]
    starts = starts[(starts > 0) & (starts < len(res.pos))]
    d[starts - 1] = 1
    return bool((d > 0).all())


@pytest.mark.parametrize("name,slice_cols", [("caterpillar100k", 2048), ("ecoli4k", 16384)])
def test_full_size_deep_and_wide_configs(port, name, slice_cols):
    """BASELINE.json configs[4] (100k-leaf caterpillar-heavy tree x 30k columns: depth, chain segments) and configs[3]
    (4k leaves x 5M columns: volume, 4883 column tiles) at FULL size on one GPU, Fitch AND Sankoff (the latter with the root
    forced to leaf 0's character, as --reference does: SURVEY 8d). The oracle checks a column slice taken from the middle --
    the generator is stateless per cell, so the slice is reproduced in isolation -- and the whole result must be
    position-sorted per node and consistent in its counts."""
    import torch

    cfg = synth.CONFIGS[name]
    tree = synth.make_tree(cfg["n_leaves"], cfg["seed"], cfg["kind"])
    spec = synth.MsaSpec(cfg["seed"], cfg["p_sub"], cfg["f_gap"], cfg["p_N"])
    C = cfg["n_cols"]
    codes4, pc = synth.simulate_msa(tree, 0, C, spec, device="cuda")
    ro_full = synth.unpack_nibbles(codes4[:1], C)[0].to(torch.int8).contiguous()  # leaf 0's character per column
    a = (C // 2 // 1024) * 1024 + 96          # not tile aligned on purpose
    b = a + slice_cols
    s4, spc = synth.simulate_msa(tree, a, b, spec, device="cuda")
    codes = synth.unpack_nibbles(s4, b - a).cpu().numpy()
    # the slice's parent codes are those of the full run (consensus = first non-gap leaf, a per-column rule)
    assert np.array_equal(spc.cpu().numpy(), pc[a:b].cpu().numpy())
    c = pb.Context(0)
    c.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
    for algo in (0, 1):
        c.upload(C, tree.n_leaves, codes4, codes4.shape[1], pc, ro_full if algo == 1 else None)
        if algo == 1:
            del codes4
            torch.cuda.empty_cache()
        t = c.run_resident(algo)
        res = c.download()
        assert res.n_mut == res.node_offsets[-1] == len(res.pos) and t.total_ms > 0
        assert res.pos.min() >= 0 and res.pos.max() < C
        assert _sorted_within_nodes(res)
        ro = codes[0].astype(np.int8) if algo == 1 else None
        want, _ = port.run(tree, algo, codes, spc.cpu().numpy(), ro, None, None, 0, n_threads=16)
        off, pos, tc = _slice_lists(res, tree.n_nodes, a, b)
        assert np.array_equal(off, want.node_offsets), (name, algo)
        assert np.array_equal(pos, want.pos + a) and np.array_equal(tc, want.type_code), (name, algo)
        del res
    c.close()


def test_caterpillar_depth(ctx, port):
    """Deep unbalanced tree (config 5's shape at reduced size: full depth handling, not full volume)."""
    tree = synth.make_tree(20000, 5, "caterpillar")
    codes4, pc = synth.simulate_msa(tree, 0, 2048, synth.MsaSpec(5, 3e-4, 0.05, 1e-3), device="cuda")
    codes = synth.unpack_nibbles(codes4, 2048).cpu().numpy()
    _set_tree(ctx, tree)
    ctx.upload(2048, tree.n_leaves, codes4, codes4.shape[1], pc)
    ctx.run_resident(0)
    want, _ = port.run(tree, 0, codes, pc.cpu().numpy(), n_threads=8)
    assert _same(ctx.download(), want)


@pytest.mark.parametrize("noise", [0.0, 0.02, 0.7])
def test_chain_segments(ctx, port, noise):
    """Deep trees are cut into chain segments that run concurrently and speculatively (tree_program.h): conserved
    columns resolve after a couple of ops, noisy ones never do and fall back to waiting for the neighbour segment."""
    rng = np.random.default_rng(int(noise * 100) + 3)
    for trial in range(8):
        kind = ["caterpillar", "unary"][trial % 2]
        tree = random_tree(int(rng.integers(200, 3000)), 4100 + trial, kind, max_arity=3)
        n_cols = int(rng.choice([700, 2048, 4100]))
        base = rng.integers(0, 5, size=n_cols)
        codes = np.repeat(base[None, :], tree.n_leaves, 0)
        codes = np.where(rng.random(codes.shape) < noise, rng.integers(0, 16, size=codes.shape), codes).astype(np.uint8)
        pc = rng.integers(0, 16, size=n_cols).astype(np.uint8)
        ro = np.where(rng.random(n_cols) < 0.3, rng.integers(0, 16, size=n_cols), -1).astype(np.int8) if trial % 2 else None
        fr = np.where(rng.random(n_cols) < 0.3, rng.integers(0, 16, size=n_cols), -1).astype(np.int8) if trial % 4 == 0 else None
        ctx.set_option("chunk_nodes", int(rng.choice([2, 5, 16, 64])))
        ctx.set_option("inline_nodes", int(rng.choice([0, 1, 3])))
        ctx.set_option("schedule", int(trial % 4 != 3))
        _set_tree(ctx, tree)
        for algo in (0, 1):
            want, want_states = port.run(tree, algo, codes, pc, ro, fr if algo == 0 else None, None, 0, n_threads=4, want_states=True)
            res = ctx.run_codes(tree, algo, codes, pc, ro, fr if algo == 0 else None, None, 0, want_states=True)
            assert _same(res, want), (noise, trial, algo, kind)
            assert np.array_equal(res.states, want_states), (noise, trial, algo)
    ctx.set_option("chunk_nodes", 0)
    ctx.set_option("inline_nodes", 3)
    ctx.set_option("schedule", 1)


def test_reroot_column_convention(ctx, port):
    """Tree::reroot (reference src/reroot.cpp:170-224) re-infers every column on the re-rooted tree with the new root's
    own character as defaultState: root override on EVERY column, 'x' read as a gap, assign-time parent state = '-'
    for gap columns and the consensus character for main columns. Same kernels, same C ABI: root_override everywhere."""
    rng = np.random.default_rng(77)
    for trial in range(6):
        tree = random_tree(int(rng.integers(5, 600)), 5100 + trial, ["binary", "caterpillar", "polytomy"][trial % 3], max_arity=4)
        n_cols = int(rng.choice([300, 1024, 3000]))
        base = rng.integers(0, 5, size=n_cols)
        codes = np.repeat(base[None, :], tree.n_leaves, 0)
        codes = np.where(rng.random(codes.shape) < 0.05, rng.integers(0, 16, size=codes.shape), codes).astype(np.uint8)
        new_root_seq = codes[int(rng.integers(0, tree.n_leaves))].astype(np.int8)       # the new root's characters
        pc = np.where(rng.random(n_cols) < 0.3, 0, base).astype(np.uint8)               # '-' for gap columns, consensus else
        _set_tree(ctx, tree)
        for algo in (0, 1):
            want, want_states = port.run(tree, algo, codes, pc, new_root_seq, None, None, 0, n_threads=4, want_states=True)
            res = ctx.run_codes(tree, algo, codes, pc, new_root_seq, None, None, 0, want_states=True)
            assert _same(res, want) and np.array_equal(res.states, want_states), (trial, algo)
            assert np.array_equal(res.states[tree.root], new_root_seq.astype(np.uint8))  # the root is what reroot forces it to be


def test_async_runs_and_shard_merge(ctx, port):
    """Two column-range shards (two contexts on this GPU, as two ranks would hold them) run asynchronously, are packed
    (pmb_pack_result) and merged (pmb_merge_packed): the merged lists equal the single-range result."""
    import torch

    rng = np.random.default_rng(31)
    tree = random_tree(300, 55, "binary")
    n_cols = 5000
    base = rng.integers(0, 5, size=n_cols)
    codes = np.repeat(base[None, :], tree.n_leaves, 0)
    codes = np.where(rng.random(codes.shape) < 0.03, rng.integers(0, 16, size=codes.shape), codes).astype(np.uint8)
    pc = codes[0].copy()
    want, _ = port.run(tree, 0, codes, pc, n_threads=4)
    ranges = [pb.column_range(2, n_cols, k) for k in range(2)]
    shards = []
    for a, b in ranges:
        c = pb.Context(0)
        c.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
        c4 = pb.pack_nibbles(codes[:, a:b])
        c.upload(b - a, tree.n_leaves, c4, c4.shape[1], np.ascontiguousarray(pc[a:b]), col_base=a)
        for _ in range(3):  # several passes in flight, one wait
            c.run_resident_async(pb.ALGO_FITCH)
        shards.append(c)
    cap = 0
    for c in shards:
        t = c.wait()
        assert t.total_ms > 0
        cap = max(cap, int(c.result_device().n_mut))
    cap += 100
    nbytes = shards[0].packed_bytes(cap)
    buf = torch.empty(2 * nbytes, dtype=torch.uint8, device="cuda")
    for k, c in enumerate(shards):
        c.pack_result(buf[k * nbytes:(k + 1) * nbytes], cap)
        c.wait()
        torch.cuda.synchronize()
    r = shards[0].merge_packed(2, buf, cap)
    torch.cuda.synchronize()
    N = tree.n_nodes

    class _Arr:
        def __init__(self, ptr, n, typestr):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}

    off = torch.as_tensor(_Arr(r.node_offsets, N + 1, "<i8"), device="cuda").cpu().numpy()
    n = int(off[-1])
    pos = torch.as_tensor(_Arr(r.pos, max(n, 1), "<i4"), device="cuda")[:n].cpu().numpy()
    tc = torch.as_tensor(_Arr(r.type_code, max(n, 1), "|u1"), device="cuda")[:n].cpu().numpy()
    assert np.array_equal(off, want.node_offsets) and np.array_equal(pos, want.pos) and np.array_equal(tc, want.type_code)
    # the run-merge comes after the shard merge: runs straddle the shard boundary
    got = shards[0].merge_runs(source=1)
    for g, w in zip(got, _oracle_merge(port, want, N)):
        assert np.array_equal(g, w)
    for c in shards:
        c.close()


def _oracle_merge(port, res, n_nodes):
    off, P, M, U = [0], [], [], []
    for v in range(n_nodes):
        a, b = res.node_offsets[v], res.node_offsets[v + 1]
        if b > a:
            p, mi, nu = port.merge_msa(res.pos[a:b], res.type_code[a:b])
            P.append(p); M.append(mi); U.append(nu)
        off.append(off[-1] + (len(P[-1]) if b > a else 0))
    from oracle.oracle import wire_mut_info

    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    mi, nu = cat(M, np.uint8), cat(U, np.uint32)
    return np.asarray(off, np.int64), cat(P, np.int32), mi, nu, wire_mut_info(mi, nu).astype(np.uint32)  # wire: src/panman.cpp:2876


def test_run_merge_on_device(ctx, port):
    """pmb_merge_runs (reference src/panman.cpp:1445-1466 + NucMut ctor src/panman.hpp:109-151) against the oracle's merge:
    long runs (cut every six), type changes inside consecutive positions, nodes without records, several tiles."""
    rng = np.random.default_rng(41)
    for trial in range(6):
        tree = random_tree(int(rng.integers(3, 300)), 6100 + trial, ["binary", "polytomy", "caterpillar"][trial % 3], max_arity=4)
        n_cols = int(rng.choice([50, 1024, 2600]))
        # block-wise columns: stretches where a clade differs from the consensus give long runs of consecutive positions
        base = rng.integers(0, 5, size=n_cols)
        codes = np.repeat(base[None, :], tree.n_leaves, 0)
        for _ in range(12):
            a = int(rng.integers(0, n_cols))
            b = min(n_cols, a + int(rng.integers(1, 40)))
            rows = rng.random(tree.n_leaves) < 0.3
            codes[np.ix_(rows, np.arange(a, b))] = rng.integers(0, 5, size=(int(rows.sum()), b - a))
        codes = np.where(rng.random(codes.shape) < 0.01, rng.integers(0, 16, size=codes.shape), codes).astype(np.uint8)
        pc = base.astype(np.uint8)
        _set_tree(ctx, tree)
        for algo in (0, 1):
            res = ctx.run_codes(tree, algo, codes, pc)
            want = _oracle_merge(port, res, tree.n_nodes)
            got = ctx.merge_runs()
            for g, w in zip(got, want):
                assert np.array_equal(g, w), (trial, algo)
            assert (got[2] >> 4).max(initial=1) <= 6 and (got[2] >> 4).min(initial=1) >= 1


def test_sync_overflow_does_not_leak_into_async_status(port):
    """A synchronous run that outgrew (and regrew) its staging pool must not leave a status bit for a later pmb_wait."""
    rng = np.random.default_rng(6)
    tree = random_tree(100, 5, "binary")
    codes = rng.integers(0, 16, size=(tree.n_leaves, 1500)).astype(np.uint8)
    want, _ = port.run(tree, 0, codes, codes[0].copy(), n_threads=2)
    c = pb.Context(0)
    c.set_option("staging_records", 500)
    c.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
    assert _same(c.run_codes(tree, 0, codes, codes[0].copy()), want)   # overflows once, resized, rerun
    c.set_option("staging_records", 0)
    c4 = pb.pack_nibbles(codes)
    c.upload(1500, tree.n_leaves, c4, c4.shape[1], codes[0].copy())
    c.run_resident_async(pb.ALGO_FITCH)
    c.wait()
    assert _same(c.download(), want)
    c.close()


def test_async_overflow_is_reported(ctx):
    rng = np.random.default_rng(5)
    tree = random_tree(100, 7, "binary")
    codes = rng.integers(0, 16, size=(tree.n_leaves, 2000)).astype(np.uint8)
    c = pb.Context(0)
    c.set_option("staging_records", 100)
    c.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
    c4 = pb.pack_nibbles(codes)
    c.upload(2000, tree.n_leaves, c4, c4.shape[1], codes[0].copy())
    c.run_resident_async(pb.ALGO_FITCH)
    with pytest.raises(pb.PanmanError) as e:
        c.wait()
    assert e.value.code == -8 and "overflow" in str(e.value)  # PMB_ERR_STAGING: its own code, not the watchdog's
    t = c.run_resident(pb.ALGO_FITCH)  # the synchronous entry sizes the pool and succeeds
    n_mut = c.download().n_mut
    assert n_mut > 100 and t.total_ms > 0
    # the pool keeps what it has grown to: the next batch of the same alignment may run asynchronously right away
    c.upload(2000, tree.n_leaves, c4, c4.shape[1], codes[0].copy())
    for _ in range(3):
        c.run_resident_async(pb.ALGO_FITCH)
    c.wait()
    assert c.download().n_mut == n_mut
    c.close()


@pytest.mark.parametrize("algo", [0, 1])
def test_two_lanes_of_asynchronous_passes(port, algo):
    """Asynchronous passes of a small problem alternate between two pipelines inside the context ("lanes" option): whichever
    lane ran last, download / result_device / merge_runs (with this context's column breaks) / pack_result see its lists;
    uploads slipped in between are ordered behind both lanes' readers; an error of a lane pass comes out of pmb_wait."""
    import torch

    rng = np.random.default_rng(70 + algo)
    tree = random_tree(250, 81, "binary")
    batches = []
    for n_cols in (3000, 1500, 4100):
        base = rng.integers(1, 5, size=n_cols)
        codes = np.repeat(base[None, :], tree.n_leaves, 0)
        codes = np.where(rng.random(codes.shape) < 0.05, rng.integers(0, 16, size=codes.shape), codes).astype(np.uint8)
        pc = base.astype(np.uint8)
        ro = codes[0].astype(np.int8) if algo else None
        want, _ = port.run(tree, algo, codes, pc, ro, None, None, 0, n_threads=4)
        batches.append((n_cols, pb.pack_nibbles(codes), pc, ro, want))
    c = pb.Context(0)
    c.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
    for lanes in (1, 0, 1):
        c.set_option("lanes", lanes)
        for k, (n_cols, c4, pc, ro, want) in enumerate(batches):
            c.upload(n_cols, tree.n_leaves, c4, c4.shape[1], pc, ro)
            for n_async in (1, 2, 5):  # the last pass on the main lane, on the second lane, on the main lane again
                for _ in range(n_async):
                    c.run_resident_async(algo)
                assert _same(c.download(), want), (lanes, k, n_async)
            brk = np.zeros(n_cols, np.uint8)
            brk[::7] = 1
            for n_async in (1, 2):
                for _ in range(n_async):
                    c.run_resident_async(algo)
                c.set_column_breaks(brk)
                got = c.merge_runs()
                # the same call on one lane only is the reference for the other
                c.set_option("lanes", 0)
                c.run_resident_async(algo)
                c.set_column_breaks(brk)
                one = c.merge_runs()
                c.set_option("lanes", lanes)
                for x, y in zip(got, one):
                    assert np.array_equal(x, y), (lanes, k, n_async, "merge_runs")
            # shard packing reads the lists of whichever lane ran last
            for n_async in (1, 2):
                for _ in range(n_async):
                    c.run_resident_async(algo)
                c.wait()
                cap = int(c.result_device().n_mut) + 8
                buf = torch.empty(c.packed_bytes(cap), dtype=torch.uint8, device="cuda")
                c.pack_result(buf, cap)
                c.wait()
                torch.cuda.synchronize()
                hdr = buf[:16].cpu().numpy().view(np.int64)
                assert int(hdr[0]) == int(want.node_offsets[-1]) and int(hdr[1]) == tree.n_nodes, (lanes, k, n_async, "pack_result")
        # a new batch right behind passes still in flight on both lanes, no wait in between
        n0, c40, pc0, ro0, want0 = batches[0]
        n1, c41, pc1, ro1, want1 = batches[1]
        c.upload(n0, tree.n_leaves, c40, c40.shape[1], pc0, ro0)
        for _ in range(4):
            c.run_resident_async(algo)
        c.upload(n1, tree.n_leaves, c41, c41.shape[1], pc1, ro1)
        for _ in range(3):
            c.run_resident_async(algo)
        assert _same(c.download(), want1), (lanes, "upload behind passes in flight")
    # another tree while passes of the old one are still in flight on both lanes, then straight on
    tree2 = random_tree(180, 82, "polytomy", max_arity=5)
    codes2 = rng.integers(0, 16, size=(tree2.n_leaves, 2500)).astype(np.uint8)
    codes2 = np.where(rng.random(codes2.shape) < 0.9, codes2[:1], codes2).astype(np.uint8)
    want2, _ = port.run(tree2, algo, codes2, codes2[0].copy(), codes2[0].astype(np.int8) if algo else None, None, None, 0, n_threads=4)
    for _ in range(3):
        c.run_resident_async(algo)
    c.set_tree(tree2.n_nodes, tree2.root, tree2.child_off, tree2.child_idx, tree2.leaf_row)
    c42 = pb.pack_nibbles(codes2)
    c.upload(2500, tree2.n_leaves, c42, c42.shape[1], codes2[0].copy(), codes2[0].astype(np.int8) if algo else None)
    for _ in range(4):
        c.run_resident_async(algo)
    assert _same(c.download(), want2), "after set_tree"
    c.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
    if algo == 1:  # a Sankoff root without a finite cost, met by the pass on the second lane
        n_cols, c4, pc, _, _ = batches[0]
        bad = np.zeros((tree.n_leaves, n_cols), np.uint8)
        lp = np.zeros(tree.n_leaves, np.uint8)  # every leaf omitted, no override (fitchSankoff.cpp:505)
        c.upload(n_cols, tree.n_leaves, pb.pack_nibbles(bad), (n_cols + 1) // 2, pc, None, None, lp)
        c.run_resident_async(1)
        c.run_resident_async(1)
        with pytest.raises(pb.PanmanError) as e:
            c.wait()
        assert e.value.code == -4
    else:  # destroyed with passes in flight on both lanes
        n_cols, c4, pc, ro, _ = batches[2]
        c.upload(n_cols, tree.n_leaves, c4, c4.shape[1], pc, ro)
        for _ in range(4):
            c.run_resident_async(algo)
    c.close()
