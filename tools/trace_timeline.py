#!/usr/bin/env python
"""Debug: per-item timeline of one pass (library option "trace") -> utilisation over time and waiting statistics.
  python tools/trace_timeline.py --config sars20k [--chunk-nodes 63] [--col-groups 1]"""
import argparse
import ctypes as C
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
    ap.add_argument("--chunk-nodes", type=int, default=0)
    ap.add_argument("--col-groups", type=int, default=1)
    ap.add_argument("--bins", type=int, default=24)
    ap.add_argument("--last", type=int, default=12)
    ap.add_argument("--window", type=float, default=30.0)
    ap.add_argument("--dump", default="")
    args = ap.parse_args()
    cfg = synth.CONFIGS[args.config]
    tree = synth.make_tree(cfg["n_leaves"], cfg["seed"], cfg["kind"])
    Cn = cfg["n_cols"]
    codes4, pc = synth.simulate_msa(tree, 0, Cn, synth.MsaSpec(cfg["seed"], cfg["p_sub"], cfg["f_gap"], cfg["p_N"]), device="cuda")
    algo = pb.ALGO_FITCH if args.algo == "fitch" else pb.ALGO_SANKOFF
    ro = synth.unpack_nibbles(codes4[:1], Cn)[0].to(torch.int8).contiguous() if algo == pb.ALGO_SANKOFF else None
    torch.cuda.synchronize()
    ctx = pb.Context(0)
    ctx.set_option("chunk_nodes", args.chunk_nodes)
    ctx.set_option("col_groups", args.col_groups)
    ctx.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
    ctx.upload(Cn, tree.n_leaves, codes4, codes4.shape[1], pc, ro)
    for _ in range(3):
        ctx.run_resident(algo)
    ctx.set_option("trace", 1)
    t = ctx.run_resident(algo)
    print(f"timings fwd {t.forward_ms:.3f} bwd {t.backward_ms:.3f} cmp {t.compact_ms:.3f} total {t.total_ms:.3f} ms")
    buf = np.zeros(64 * 1024 * 1024 // 8, np.uint64)
    ctx.L.pmb_debug_trace.restype = C.c_longlong
    ctx.L.pmb_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
    n = ctx.L.pmb_debug_trace(ctx.h, buf.ctypes.data, len(buf))
    rec = buf[:n].reshape(-1, 4)
    if args.dump:
        np.save(args.dump, rec)
    half = len(rec) // 2
    origin = rec[rec[:, 0] > 0, 0].min()
    for name, r in (("forward", rec[:half]), ("backward", rec[half:])):
        r = r[r[:, 0] > 0]
        t0 = (r[:, 0] - origin).astype(np.float64) / 1e3
        t1 = (r[:, 1] - origin).astype(np.float64) / 1e3
        waited = (r[:, 3] & 0xFFFFFFFF).astype(np.float64) / 1e3
        dur = t1 - t0
        print(f"== {name}: {len(r)} items, span {t0.min():.1f} .. {t1.max():.1f} us; item duration mean {dur.mean():.1f} "
              f"p50 {np.median(dur):.1f} p99 {np.percentile(dur, 99):.1f} max {dur.max():.1f} us; "
              f"waiting: {int((waited > 0).sum())} items, total {waited.sum() / 1e3:.2f} ms, max {waited.max():.1f} us")
        lo, hi = t0.min(), t1.max()
        edges = np.linspace(lo, hi, args.bins + 1)
        line_a, line_w = [], []
        for a, b in zip(edges[:-1], edges[1:]):
            ov = np.clip(np.minimum(t1, b) - np.maximum(t0, a), 0, None)
            active = ov.sum() / (b - a)
            # waiting time is not located inside the item; approximate by spreading it over the item's span
            wv = (ov * (waited / np.maximum(dur, 1e-9))).sum() / (b - a)
            line_a.append(active)
            line_w.append(wv)
        print("   time(us)  " + " ".join(f"{e:6.0f}" for e in edges[:-1]))
        print("   warps     " + " ".join(f"{x:6.0f}" for x in line_a))
        print("   ~waiting  " + " ".join(f"{x:6.0f}" for x in line_w))
        # the items that finish last: which chunks are they, how long did they run and wait
        chunk = (r[:, 2] >> np.uint64(32)).astype(np.int64)
        tile = (r[:, 2] & np.uint64(0xFFFFFFFF)).astype(np.int64)
        order = np.argsort(-t1)[:args.last]
        print(f"   last {args.last} items to finish (end us, start us, waited us, chunk, tile):")
        print("   " + "  ".join(f"({t1[i]:.0f},{t0[i]:.0f},{waited[i]:.0f},c{chunk[i]},t{tile[i]})" for i in order))
        # per-chunk summary of the late finishers
        late = t1 > (t1.max() - args.window)
        cs, counts = np.unique(chunk[late], return_counts=True)
        print(f"   chunks with items finishing in the last {args.window:.0f} us: {len(cs)}; "
              + ", ".join(f"c{c}x{n}" for c, n in list(zip(cs, counts))[:40]))
    ctx.close()


if __name__ == "__main__":
    main()
