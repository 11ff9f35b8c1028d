#!/usr/bin/env python
"""Host-to-device ceiling of the box with N ranks copying at once (launch with torch.distributed.run, one rank per GPU):
every rank copies 1 GiB of its own page-locked memory to its own GPU, plain cudaMemcpyAsync through torch, nothing of the
library involved. Prints GB/s per rank when the ranks copy alone (one after the other) and when they all copy together.
  python -m torch.distributed.run --nproc-per-node 8 tools/h2d_probe.py"""
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << 30
    src = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    src.fill_(7)
    dst = torch.empty(n, dtype=torch.uint8, device="cuda")

    def copy_rate(reps=4):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        return reps * n / (time.perf_counter() - t0) / 1e9

    alone = torch.zeros(world, dtype=torch.float64, device="cuda")
    for r in range(world):  # one rank at a time
        if world > 1:
            dist.barrier()
        if r == rank:
            alone[r] = copy_rate()
    if world > 1:
        dist.barrier()
    together = torch.zeros(world, dtype=torch.float64, device="cuda")
    together[rank] = copy_rate()
    if world > 1:
        dist.all_reduce(alone)
        dist.all_reduce(together)
    if rank == 0:
        print("H2D GB/s per rank, copying alone   :", [round(float(x), 1) for x in alone])
        print("H2D GB/s per rank, all at once     :", [round(float(x), 1) for x in together], "sum", round(float(together.sum()), 1))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
