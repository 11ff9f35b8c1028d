"""GPU parity tests of the column-sharded multi-GPU entry (pmb_group_*, include/panman_b200.h): every rank runs the pass on
its column range, packs its lists into rank 0's mailbox (peer memory), rank 0 concatenates per node and run-merges after
that. Checked bit-exact against the oracle on the WHOLE alignment: merged lists and merged NucMut fields.

A group may list the same device several times (several ranks on one GPU): the hand-shake (stream memory operations on
the arrive / credit words), the double-buffered mailbox and the merge are then exactly those of a multi-GPU run, so these
tests are meaningful on a one-GPU box; with several GPUs visible the same cases also run across all of them, and the
two-process test maps rank 0's mailbox through CUDA IPC as under torchrun."""
import os
import socket

import numpy as np
import pytest

import panman_b200 as pb
from oracle.oracle import PortOracle, random_tree
from panman_b200 import synth
from tests.golden_util import merge_all_nodes

pytestmark = pytest.mark.gpu


def _n_devices():
    import torch

    return torch.cuda.device_count()


def _case(seed, n_leaves=300, n_cols=5000, kind="binary", noise=0.03):
    rng = np.random.default_rng(seed)
    tree = random_tree(n_leaves, seed + 7, kind, max_arity=4)
    base = rng.integers(0, 5, size=n_cols)
    codes = np.repeat(base[None, :], tree.n_leaves, 0)
    for _ in range(10):  # clade-wide stretches: long runs of consecutive positions, some across range boundaries
        a = int(rng.integers(0, n_cols))
        b = min(n_cols, a + int(rng.integers(1, 60)))
        rows = rng.random(tree.n_leaves) < 0.3
        codes[np.ix_(rows, np.arange(a, b))] = rng.integers(0, 5, size=(int(rows.sum()), b - a))
    for edge in range(1024, n_cols, 1024):  # and one across every possible boundary
        rows = rng.random(tree.n_leaves) < 0.5
        codes[np.ix_(rows, np.arange(edge - 4, min(n_cols, edge + 5)))] = 3
    codes = np.where(rng.random(codes.shape) < noise, rng.integers(0, 16, size=codes.shape), codes).astype(np.uint8)
    return tree, codes, base.astype(np.uint8)


def _same(res, want):
    return (np.array_equal(res.node_offsets, want.node_offsets) and np.array_equal(res.pos, want.pos)
            and np.array_equal(res.type_code, want.type_code))


def _check_group(devices, tree, codes, pc, port, algos=(0, 1), steps=5):
    g = pb.Group(devices)
    g.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
    c4 = pb.pack_nibbles(codes)
    n_cols = codes.shape[1]
    for algo in algos:
        ro = codes[0].astype(np.int8) if algo == 1 else None
        want, _ = port.run(tree, algo, codes, pc, ro, None, None, 0, n_threads=8)
        # end to end: host buffers in, merged host lists out (sizes its own mailbox on the first call)
        res = g.run_nuc(algo, n_cols, tree.n_leaves, c4, c4.shape[1], pc, ro)
        assert _same(res, want), (devices, algo, "run_nuc")
        # resident: several steps back to back (credits, both mailbox parities), one wait
        g.upload(n_cols, tree.n_leaves, c4, c4.shape[1], pc, ro)
        for _ in range(steps):
            g.run_async(algo)
        g.wait()
        assert _same(g.download(), want), (devices, algo, "resident")
        got = g.merge_runs()
        for x, y in zip(got, merge_all_nodes(port, want.node_offsets, want.pos, want.type_code)):
            assert np.array_equal(x, y), (devices, algo, "NucMut fields")
    g.close()


@pytest.mark.parametrize("world", [1, 2, 3])
def test_group_ranks_on_one_device(port, world):
    tree, codes, pc = _case(11 + world)
    _check_group([0] * world, tree, codes, pc, port)


def test_group_more_ranks_than_tiles(port):
    """2 000 columns are two tiles: ranks 2 and 3 own empty ranges and sit the pass out (no shard, no hand-shake)."""
    tree, codes, pc = _case(5, n_leaves=120, n_cols=2000)
    _check_group([0] * 4, tree, codes, pc, port, steps=3)
    assert pb.column_range(4, 2000, 2) == (2000, 2000)


def test_group_over_all_devices(port):
    n = _n_devices()
    if n < 2:
        pytest.skip("one GPU visible: the multi-device mapping (peer access over NVLink) needs two")
    tree, codes, pc = _case(23, n_leaves=500, n_cols=1024 * 2 * n + 300, kind="caterpillar")
    _check_group(list(range(n)), tree, codes, pc, port)
    tree, codes, pc = _case(24, n_leaves=200, n_cols=1024 * n)
    _check_group(list(range(n)), tree, codes, pc, port, steps=8)


def test_group_capacity_and_staging_errors(port):
    """A mailbox too small for a shard and a staging pool too small for an asynchronous pass are reported by the wait with
    their own codes (PMB_ERR_CAPACITY / PMB_ERR_STAGING) and the group recovers once more is reserved."""
    tree, codes, pc = _case(31, n_leaves=150, n_cols=3000, noise=0.2)
    want, _ = port.run(tree, 0, codes, pc, n_threads=4)
    c4 = pb.pack_nibbles(codes)
    g = pb.Group([0, 0])
    g.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
    g.reserve(64)
    g.upload(3000, tree.n_leaves, c4, c4.shape[1], pc)
    g.run_async(0)
    with pytest.raises(pb.PanmanError) as e:
        g.wait()
    assert e.value.code == -9
    g.reserve(int(want.node_offsets[-1]) + 16)
    g.run_async(0)
    g.run_async(0)
    g.wait()
    assert _same(g.download(), want)
    g.ctx(1).set_option("staging_records", 256)
    g.upload(3000, tree.n_leaves, c4, c4.shape[1], pc)
    g.run_async(0)
    with pytest.raises(pb.PanmanError) as e:
        g.wait()
    assert e.value.code == -8 and "rank 1" in str(e.value)
    g.run_async(0)  # the pool was grown by the failed wait
    g.wait()
    assert _same(g.download(), want)
    g.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port_no, q):
    """One process per rank, as under torchrun: handles exchanged through torch.distributed, mailbox mapped by CUDA IPC."""
    import torch
    import torch.distributed as dist

    from panman_b200.distributed import connect_group

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    tree, codes, pc = _case(41, n_leaves=400, n_cols=7000)
    a, b = pb.column_range(world, codes.shape[1], rank)
    g = pb.Group([dev], rank_base=rank, world=world)
    g.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
    c4 = pb.pack_nibbles(codes[:, a:b])
    ok = True
    for algo in (0, 1):
        ro = codes[0].astype(np.int8) if algo == 1 else None
        g.upload_shard(0, codes.shape[1], tree.n_leaves, c4, c4.shape[1], np.ascontiguousarray(pc[a:b]),
                       None if ro is None else np.ascontiguousarray(ro[a:b]))
        if algo == 0:  # size the mailbox from a first pass of every rank, then connect once
            t = g.ctx(0).run_resident(algo)
            assert t.total_ms > 0
            connect_group(g, dist, int(g.ctx(0).result_device().n_mut * 2) + 1024)
        for _ in range(4):
            g.run_async(algo)
        g.wait()
        res = g.download()
        if rank == 0:
            port = PortOracle()
            want, _ = port.run(tree, algo, codes, pc, ro, None, None, 0, n_threads=4)
            ok = ok and _same(res, want)
            got = g.merge_runs()
            ok = ok and all(np.array_equal(x, y) for x, y in zip(got, merge_all_nodes(port, want.node_offsets, want.pos, want.type_code)))
        else:
            ok = ok and res.n_mut == 0
        dist.barrier()
    q.put((rank, bool(ok)))
    dist.barrier()
    g.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_group_one_process_per_rank(world):
    """world processes (on as many GPUs as are visible, several per GPU otherwise) form one group through
    pmb_group_export / pmb_group_connect."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port_no, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got == {r: True for r in range(world)}


@pytest.mark.parametrize("name", ["sars20k", "ecoli4k"])
def test_group_full_size_configs(port, name):
    """BASELINE.json configs[1] whole, and a 64-tile slice of configs[3] (4k leaves: the full 5M columns are checked through
    properties in test_gpu_parity.py), sharded over all visible GPUs (two ranks on one GPU otherwise): merged lists and
    merged NucMut fields against the oracle on the whole range."""
    import torch

    cfg = synth.CONFIGS[name]
    n_cols = cfg["n_cols"] if name == "sars20k" else 65536 + 777
    tree = synth.make_tree(cfg["n_leaves"], cfg["seed"], cfg["kind"])
    codes4, pc = synth.simulate_msa(tree, 0, n_cols, synth.MsaSpec(cfg["seed"], cfg["p_sub"], cfg["f_gap"], cfg["p_N"]), device="cuda")
    codes = synth.unpack_nibbles(codes4, n_cols).cpu().numpy()
    h4, hpc = codes4.cpu().numpy(), pc.cpu().numpy()
    del codes4
    torch.cuda.empty_cache()
    n = _n_devices()
    devices = list(range(n)) if n > 1 else [0, 0]
    g = pb.Group(devices)
    g.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
    for algo in (0, 1):
        ro = codes[0].astype(np.int8) if algo == 1 else None
        res = g.run_nuc(algo, n_cols, tree.n_leaves, h4, h4.shape[1], hpc, ro)
        want, _ = port.run(tree, algo, codes, hpc, ro, None, None, 0, n_threads=16)
        assert _same(res, want), (name, algo)
        got = g.merge_runs()
        for x, y in zip(got, merge_all_nodes(port, want.node_offsets, want.pos, want.type_code)):
            assert np.array_equal(x, y), (name, algo)
    g.close()
