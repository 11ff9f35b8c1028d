#!/usr/bin/env python
"""Tuning sweep on one GPU: build a named workload once, then time pmb_run_resident under several option sets.
  python tools/sweep.py --config sars20k --algo fitch --grid "chunk_nodes=32,63,128 schedule=0,1 inline_nodes=3"
Prints one line per option set: device ms per phase (median of --steps), node*col/s, fraction of the HBM roofline."""
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import panman_b200 as pb  # noqa: E402
from panman_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="sars20k")
    ap.add_argument("--algo", default="fitch")
    ap.add_argument("--cols", type=int, default=0)
    ap.add_argument("--leaves", type=int, default=0)
    ap.add_argument("--steps", type=int, default=7)
    ap.add_argument("--grid", default="chunk_nodes=0 schedule=1")
    args = ap.parse_args()
    cfg = dict(synth.CONFIGS[args.config])
    if args.cols:
        cfg["n_cols"] = args.cols
    if args.leaves:
        cfg["n_leaves"] = args.leaves
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    tree = synth.make_tree(cfg["n_leaves"], cfg["seed"], cfg["kind"])
    C = cfg["n_cols"]
    codes4, pc = synth.simulate_msa(tree, 0, C, synth.MsaSpec(cfg["seed"], cfg["p_sub"], cfg["f_gap"], cfg["p_N"]), device="cuda")
    algo = pb.ALGO_FITCH if args.algo == "fitch" else pb.ALGO_SANKOFF
    ro = synth.unpack_nibbles(codes4[:1], C)[0].to(torch.int8).contiguous() if algo == pb.ALGO_SANKOFF else None
    torch.cuda.synchronize()
    keys, vals = [], []
    for tok in args.grid.split():
        k, v = tok.split("=")
        keys.append(k)
        vals.append([int(x) for x in v.split(",")])
    print(f"# {args.config} {args.algo}: {tree.n_nodes} nodes x {C} cols; peak {peak} GB/s", flush=True)
    ref_sig = None
    for combo in itertools.product(*vals):
        ctx = pb.Context(0)
        for k, v in zip(keys, combo):
            ctx.set_option(k, v)
        ctx.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
        ctx.upload(C, tree.n_leaves, codes4, codes4.shape[1], pc, ro)
        ts = []
        try:
            for i in range(3 + args.steps):
                t = ctx.run_resident(algo)
                if i >= 3:
                    ts.append((t.forward_ms, t.backward_ms, t.compact_ms, t.total_ms))
        except pb.PanmanError as e:
            print(dict(zip(keys, combo)), "ERROR", e, flush=True)
            ctx.close()
            continue
        # pipelined: asynchronous passes back to back, as bench.py times them (consecutive passes overlap)
        lib = torch.cuda.ExternalStream(ctx.stream_handle())
        lib_end = torch.cuda.ExternalStream(ctx.result_stream_handle())
        for _ in range(3):
            ctx.run_resident_async(algo)
        ctx.wait()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(lib)
        for _ in range(20):
            ctx.run_resident_async(algo)
        ctx.join()
        e1.record(lib)
        ctx.wait()
        torch.cuda.synchronize()
        piped = e0.elapsed_time(e1) / 20
        res = ctx.download()
        sig = (int(res.n_mut), int(res.pos.astype(np.int64).sum()), int(res.type_code.astype(np.int64).sum()))
        if ref_sig is None:
            ref_sig = sig
        m = np.median(np.asarray(ts), 0)
        ab = ctx.algorithmic_bytes(algo)
        print(f"{dict(zip(keys, combo))} fwd {m[0]:.3f} bwd {m[1]:.3f} cmp {m[2]:.3f} total {m[3]:.3f} ms, pipelined {piped:.4f} ms = {ab / (piped * 1e-3) / 1e9 / peak:.3f} | "
              f"{tree.n_nodes * C / (m[3] * 1e-3):.3e} node*col/s | roofline {ab / (m[3] * 1e-3) / 1e9 / peak:.3f} | "
              f"launches {t.n_launches} levels {t.n_levels} | same_result {sig == ref_sig}", flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
