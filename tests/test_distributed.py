"""Multi-rank path on the CPU: world_size-2 gloo processes shard the columns, run the (oracle) pass on their range and
gather the lists to rank 0 exactly as bench.py does over NCCL; the merged lists must equal the single-range result."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp

from oracle.oracle import PortOracle, random_tree
from panman_b200.distributed import column_ranges, gather_lists, merge_lists


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


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tree, codes, pc = _inputs()
    a, b = column_ranges(codes.shape[1], world)[rank]
    lists, _ = PortOracle().run(tree, 0, codes[:, a:b], pc[a:b], n_threads=1)
    off = torch.from_numpy(lists.node_offsets)
    pos = torch.from_numpy(lists.pos + a)  # col_base
    tc = torch.from_numpy(lists.type_code)
    out = gather_lists(dist, rank, world, off, pos, tc)
    if rank == 0:
        q.put([t.numpy() for t in out])
    dist.barrier()
    dist.destroy_process_group()


def test_column_ranges_are_tile_aligned():
    r = column_ranges(5000, 2)
    assert r == [(0, 2048), (2048, 5000)]
    r = column_ranges(30000, 8)
    assert r[0][0] == 0 and r[-1][1] == 30000 and all(a % 1024 == 0 for a, _ in r)
    assert all(r[i][1] == r[i + 1][0] for i in range(7))
    assert column_ranges(100, 4)[0] == (0, 0) or sum(b - a for a, b in column_ranges(100, 4)) == 100


def test_merge_lists_matches_single_range():
    tree, codes, pc = _inputs()
    port = PortOracle()
    whole, _ = port.run(tree, 0, codes, pc, n_threads=2)
    offs, poss, tcs = [], [], []
    for a, b in column_ranges(codes.shape[1], 3):
        part, _ = port.run(tree, 0, codes[:, a:b], pc[a:b], n_threads=1)
        offs.append(torch.from_numpy(part.node_offsets))
        poss.append(torch.from_numpy(part.pos + a))
        tcs.append(torch.from_numpy(part.type_code))
    off, pos, tc = merge_lists(offs, poss, tcs)
    assert np.array_equal(off.numpy(), whole.node_offsets)
    assert np.array_equal(pos.numpy(), whole.pos) and np.array_equal(tc.numpy(), whole.type_code)


def test_two_rank_gloo_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    off, pos, tc = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    tree, codes, pc = _inputs()
    whole, _ = PortOracle().run(tree, 0, codes, pc, n_threads=2)
    assert np.array_equal(off, whole.node_offsets) and np.array_equal(pos, whole.pos) and np.array_equal(tc, whole.type_code)
