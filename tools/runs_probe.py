#!/usr/bin/env python
"""The ingest kernels alone on a slice of a named configuration: pack_leaves_kernel (nibble matrix already in device memory)
against expand_runs_kernel (clade-run events already in device memory is not an option of the ABI, so: events from
page-locked memory + expansion), CUDA events on the library's stream. Also the run to put under ncu:
  ncu --set full --clock-control none -k regex:expand_runs -c 3 python tools/runs_probe.py --config ecoli4k --cols 1250304"""
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
    ap.add_argument("--config", default="ecoli4k")
    ap.add_argument("--cols", type=int, default=1250304)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    cfg = synth.CONFIGS[a.config]
    C = min(a.cols, cfg["n_cols"])
    dev = torch.device("cuda", 0)
    tree = synth.make_tree(cfg["n_leaves"], cfg["seed"], cfg["kind"])
    codes4, pc = synth.simulate_msa(tree, 0, C, synth.MsaSpec(cfg["seed"], cfg["p_sub"], cfg["f_gap"], cfg["p_N"]), device=dev)
    h_codes = torch.empty(codes4.shape, dtype=torch.uint8, pin_memory=True).copy_(codes4)
    h_pc = torch.empty(pc.shape, dtype=torch.uint8, pin_memory=True).copy_(pc)
    torch.cuda.synchronize()
    ctx = pb.Context(0)
    ctx.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
    lib = torch.cuda.ExternalStream(ctx.stream_handle(), device=dev)
    t0 = time.perf_counter()
    runs = pb.Runs.of_tree(tree, C, h_codes.numpy(), h_pc.numpy())
    enc = time.perf_counter() - t0
    plane_gb = tree.n_leaves * ((C + 1023) // 1024) * 512 / 1e9

    def timed(f):
        f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(lib)
        for _ in range(a.reps):
            f()
        e1.record(lib)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.reps

    ms_dev = timed(lambda: ctx.upload(C, tree.n_leaves, codes4, codes4.shape[1], pc))
    ms_host = timed(lambda: ctx.upload(C, tree.n_leaves, h_codes, h_codes.shape[1], h_pc))
    ms_runs = timed(lambda: ctx.upload_runs(runs, h_pc))
    t = ctx.run_resident(pb.ALGO_FITCH)
    print(f"{a.config} {tree.n_leaves} leaves x {C} columns: plane matrix {plane_gb:.2f} GB, nibble matrix {codes4.numel() / 1e9:.2f} GB, "
          f"events {runs.nbytes / 1e6:.1f} MB ({runs.n_events / C:.2f} per column), host encode {enc * 1e3:.0f} ms")
    print(f"  ingest from device nibbles (pack_leaves_kernel)        {ms_dev:8.3f} ms  = {plane_gb / ms_dev * 1e3:7.0f} GB/s of planes written")
    print(f"  ingest from page-locked nibbles (copy + pack)         {ms_host:8.3f} ms  = {codes4.numel() / ms_host / 1e6:7.1f} GB/s over the link")
    print(f"  ingest from page-locked events (copy + expand_runs)   {ms_runs:8.3f} ms  = {plane_gb / ms_runs * 1e3:7.0f} GB/s of planes written")
    print(f"  one Fitch pass on it                                  {t.total_ms:8.3f} ms")
    runs.close()
    ctx.close()


if __name__ == "__main__":
    main()
