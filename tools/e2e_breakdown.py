#!/usr/bin/env python
"""Where the end-to-end time of pmb_run_nuc goes (config 2 by default): upload / pass / download with host timers,
for pinned-host, pageable-host and device-resident inputs, plus the raw H2D copy rate of the same bytes."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import panman_b200 as pb  # noqa: E402
from panman_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="sars20k")
    args = ap.parse_args()
    cfg = synth.CONFIGS[args.config]
    tree = synth.make_tree(cfg["n_leaves"], cfg["seed"], cfg["kind"])
    C = cfg["n_cols"]
    codes4, pc = synth.simulate_msa(tree, 0, C, synth.MsaSpec(cfg["seed"], cfg["p_sub"], cfg["f_gap"], cfg["p_N"]), device="cuda")
    h_pin = torch.empty(codes4.shape, dtype=torch.uint8, pin_memory=True).copy_(codes4)
    h_page = h_pin.clone()  # pageable
    h_pc = torch.empty(pc.shape, dtype=torch.uint8, pin_memory=True).copy_(pc)
    torch.cuda.synchronize()
    ctx = pb.Context(0)
    ctx.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
    dst = torch.empty_like(codes4)
    for _ in range(3):
        dst.copy_(h_pin, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        dst.copy_(h_pin, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"raw pinned H2D copy: {codes4.numel() / 1e6:.0f} MB in {dt * 1e3:.2f} ms = {codes4.numel() / dt / 1e9:.1f} GB/s")
    for name, src in (("pinned host", h_pin), ("pageable host", h_page), ("device", codes4)):
        up = run = down = 0.0
        n = 5
        for i in range(n + 2):
            t0 = time.perf_counter()
            ctx.upload(C, tree.n_leaves, src, src.shape[1], h_pc)
            t1 = time.perf_counter()
            ctx.run_resident(0)
            t2 = time.perf_counter()
            ctx.download(copy=False)
            t3 = time.perf_counter()
            if i >= 2:
                up += (t1 - t0) / n
                run += (t2 - t1) / n
                down += (t3 - t2) / n
        print(f"{name:14s}: upload {up * 1e3:.2f} ms  pass {run * 1e3:.2f} ms  download {down * 1e3:.2f} ms  "
              f"total {(up + run + down) * 1e3:.2f} ms")


if __name__ == "__main__":
    main()
