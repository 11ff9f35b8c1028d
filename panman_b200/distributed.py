"""One process per GPU (torchrun): forming a pmb_group across the ranks of a torch.distributed job.

The library owns the data path (column ranges, the gather of the per-rank lists into rank 0's mailbox over NVLink, the
merge); all it needs from the job launcher is ONE all-gather of the ranks' 128-byte mailbox handles -- the role
ncclGetUniqueId + broadcast plays for NCCL. This module does that exchange with torch.distributed (NCCL on the GPU box,
gloo in the CPU tests) and nothing else; bench.py and the multi-GPU tests call it.
"""
import torch

from .api import Group
from .lib import GROUP_HANDLE_BYTES


def exchange_bytes(dist, blob: bytes, device="cpu") -> bytes:
    """All-gather of equally sized byte strings in rank order (NCCL needs a CUDA `device`, gloo the CPU)."""
    world = dist.get_world_size()
    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(device)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return b"".join(bytes(t.cpu().numpy().tobytes()) for t in out)


def agree_max(dist, value: int, device="cpu") -> int:
    """The largest `value` over the ranks (every rank must reserve the same mailbox capacity)."""
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())


def connect_group(group: Group, dist, capacity: int, device="cpu"):
    """reserve -> export -> all-gather -> connect, the same on every rank. `capacity` (records per shard) is maximised
    over the ranks first."""
    if group.world == 1:
        return
    group.reserve(agree_max(dist, capacity, device))
    handles = exchange_bytes(dist, group.export(), device)
    assert len(handles) == group.world * GROUP_HANDLE_BYTES
    group.connect(handles)
    dist.barrier()
