"""Column-range sharding across ranks and the gather of per-rank mutation lists (torch.distributed plumbing).

Columns are independent units of the Fitch/Sankoff path, so rank r simply owns a contiguous column range and runs the
whole pass on it (tree replicated, no collective inside the kernels). The one exchange is at the end: every rank's
per-node lists go to rank 0, where the merged list of a node is the concatenation of the ranks' lists in rank order --
each is already in ascending position and ranges are disjoint and ordered, so nothing is sorted. The reference's <=6
run-merge (src/panman.cpp:1445-1466) must run after this gather: a run can straddle a shard boundary.
Works on CUDA tensors over NCCL (device pointers of pmb_result_device, zero-copy) and on CPU tensors over gloo (tests).
"""
import torch


def column_ranges(n_cols: int, world: int, granule: int = 1024):
    """Contiguous ranges aligned to the kernels' 1024-column tiles (the last one takes the remainder)."""
    tiles = (n_cols + granule - 1) // granule
    out = []
    for r in range(world):
        a = tiles * r // world * granule
        b = min(n_cols, tiles * (r + 1) // world * granule)
        out.append((a, max(a, b)))
    return out


def gather_lists(dist, rank: int, world: int, node_offsets: torch.Tensor, pos: torch.Tensor, type_code: torch.Tensor):
    """node_offsets int64 [N+1], pos int32 [n], type_code uint8 [n] of this rank. Returns (offsets, pos, type_code) of
    the whole alignment on rank 0, None elsewhere."""
    dev = node_offsets.device
    N = node_offsets.numel() - 1
    n = int(pos.numel())
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([n], dtype=torch.int64, device=dev))
    cl = [int(c.item()) for c in counts]
    offs = [torch.empty(N + 1, dtype=torch.int64, device=dev) for _ in range(world)] if rank == 0 else None
    dist.gather(node_offsets.contiguous(), offs, dst=0)
    if rank != 0:
        if n:
            for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, pos.contiguous(), 0),
                                             dist.P2POp(dist.isend, type_code.contiguous(), 0)]):
                w.wait()
        return None
    poss = [pos] + [torch.empty(cl[k], dtype=torch.int32, device=dev) for k in range(1, world)]
    tcs = [type_code] + [torch.empty(cl[k], dtype=torch.uint8, device=dev) for k in range(1, world)]
    ops = []
    for k in range(1, world):
        if cl[k]:
            ops += [dist.P2POp(dist.irecv, poss[k], k), dist.P2POp(dist.irecv, tcs[k], k)]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return merge_lists(offs, poss, tcs)


def merge_lists(offs, poss, tcs):
    """Concatenate per-node lists of several column ranges in range order."""
    dev = offs[0].device
    N = offs[0].numel() - 1
    cnt = torch.stack([o[1:] - o[:-1] for o in offs])  # ranges x N
    before = torch.cumsum(cnt, 0) - cnt
    merged_off = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    merged_off[1:] = torch.cumsum(cnt.sum(0), 0)
    total = int(merged_off[-1])
    mpos = torch.empty(total, dtype=torch.int32, device=dev)
    mtc = torch.empty(total, dtype=torch.uint8, device=dev)
    for k in range(len(offs)):
        nk = int(poss[k].numel())
        if nk == 0:
            continue
        shift = merged_off[:-1] + before[k] - offs[k][:-1]
        idx = torch.repeat_interleave(shift, cnt[k]) + torch.arange(nk, device=dev)
        mpos[idx] = poss[k]
        mtc[idx] = tcs[k]
    return merged_off, mpos, mtc
