"""Multi-rank path on the CPU. The data path of a pmb_group (peer-memory gather, device merge) needs GPUs and is covered
by tests/test_group_gpu.py; what runs without one is everything around it: the column ranges the library hands out, the
handle exchange bench.py uses (panman_b200.distributed, here over world_size-2 gloo), and the sharding semantics
themselves -- every rank runs the (oracle) pass on ITS range, rank 0 concatenates the shards per node in rank order and
run-merges AFTER that, and the result must equal the single-range result (reference: sort + merge after all columns,
src/panman.cpp:1445-1466)."""
import os
import pickle
import socket

import numpy as np
import torch.multiprocessing as mp

import panman_b200 as pb
from oracle.oracle import PortOracle, random_tree
from panman_b200.distributed import agree_max, exchange_bytes
from tests.golden_util import concat_shards, merge_all_nodes


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs():
    rng = np.random.default_rng(3)
    tree = random_tree(150, 42, "binary")
    n_cols = 3000
    base = rng.integers(0, 5, size=n_cols)
    codes = np.repeat(base[None, :], tree.n_leaves, 0)
    codes = np.where(rng.random(codes.shape) < 0.05, rng.integers(0, 16, size=codes.shape), codes).astype(np.uint8)
    return tree, codes, codes[0].copy()


def test_column_ranges():
    """Contiguous, ascending, 1024-aligned, covering; ranks beyond the tile count get empty ranges at the END."""
    for world, n_cols in ((2, 5000), (8, 30000), (8, 5_000_000), (4, 100), (8, 3000), (3, 1024), (1, 7)):
        r = [pb.column_range(world, n_cols, k) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == n_cols
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        assert all(a % 1024 == 0 or a == n_cols for a, _ in r)
        sizes = [b - a for a, b in r]
        assert all(s >= 0 for s in sizes) and sum(sizes) == n_cols
        empties = [s == 0 for s in sizes]
        assert empties == sorted(empties), "empty ranges must come last"
        tiles = [(s + 1023) // 1024 for s in sizes]
        assert max(tiles) - min(t for t in tiles) <= 1
    assert pb.column_range(2, 5000, 0) == (0, 3072) and pb.column_range(2, 5000, 1) == (3072, 5000)
    assert pb.column_range(4, 100, 0) == (0, 100) and pb.column_range(4, 100, 3) == (100, 100)


def test_concatenated_shards_match_single_range():
    tree, codes, pc = _inputs()
    port = PortOracle()
    whole, _ = port.run(tree, 0, codes, pc, n_threads=2)
    parts = []
    for k in range(3):
        a, b = pb.column_range(3, codes.shape[1], k)
        part, _ = port.run(tree, 0, codes[:, a:b], pc[a:b], n_threads=1)
        parts.append((part.node_offsets, part.pos + a, part.type_code))
    off, pos, tc = concat_shards(parts)
    assert np.array_equal(off, whole.node_offsets)
    assert np.array_equal(pos, whole.pos) and np.array_equal(tc, whole.type_code)
    # the <= 6 run-merge after the concatenation equals the merge of the whole; per shard it would not (runs straddle)
    want = merge_all_nodes(port, whole.node_offsets, whole.pos, whole.type_code)
    got = merge_all_nodes(port, off, pos, tc)
    assert all(np.array_equal(x, y) for x, y in zip(got, want))
    per_shard = sum(len(merge_all_nodes(port, o, p, t)[1]) for o, p, t in parts)
    assert per_shard >= len(want[1])


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tree, codes, pc = _inputs()
    a, b = pb.column_range(world, codes.shape[1], rank)
    lists, _ = PortOracle().run(tree, 0, codes[:, a:b], pc[a:b], n_threads=1)
    # the exchange bench.py performs for the mailbox handles: fixed-size blobs, all-gathered in rank order
    cap = agree_max(dist, int(lists.node_offsets[-1]))
    blob = pickle.dumps((rank, a, b, int(lists.node_offsets[-1])))
    blobs = exchange_bytes(dist, blob.ljust(128, b"\0"))
    assert len(blobs) == 128 * world
    # the shards themselves (on the GPU box: packed into rank 0's mailbox by the packing kernel)
    payload = pickle.dumps((lists.node_offsets, lists.pos + a, lists.type_code))
    size = agree_max(dist, len(payload))
    shards = exchange_bytes(dist, payload.ljust(size, b"\0"))
    if rank == 0:
        q.put((cap, [pickle.loads(blobs[k * 128:(k + 1) * 128]) for k in range(world)],
               [pickle.loads(shards[k * size:(k + 1) * size]) for k in range(world)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_exchange_and_merge():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    cap, heads, shards = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    tree, codes, pc = _inputs()
    assert [h[0] for h in heads] == [0, 1] and heads[0][2] == heads[1][1] and heads[1][2] == codes.shape[1]
    assert cap == max(h[3] for h in heads)
    whole, _ = PortOracle().run(tree, 0, codes, pc, n_threads=2)
    off, pos, tc = concat_shards(shards)
    assert np.array_equal(off, whole.node_offsets) and np.array_equal(pos, whole.pos) and np.array_equal(tc, whole.type_code)
