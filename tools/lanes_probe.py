#!/usr/bin/env python
"""How much would K independent pass pipelines on one GPU give a small problem? K contexts on the same device, the same
batch resident in each, asynchronous passes issued round-robin; wall time per pass over all of them.
  python tools/lanes_probe.py --config indel10k --algo fitch --lanes 1,2,3"""
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
    ap.add_argument("--config", default="indel10k")
    ap.add_argument("--algo", default="fitch")
    ap.add_argument("--lanes", default="1,2,3")
    ap.add_argument("--grid", default="0")
    ap.add_argument("--passes", type=int, default=60)
    a = ap.parse_args()
    cfg = synth.CONFIGS[a.config]
    tree = synth.make_tree(cfg["n_leaves"], cfg["seed"], cfg["kind"])
    C = cfg["n_cols"]
    codes4, pc = synth.simulate_msa(tree, 0, C, synth.MsaSpec(cfg["seed"], cfg["p_sub"], cfg["f_gap"], cfg["p_N"]), device="cuda")
    algo = pb.ALGO_FITCH if a.algo == "fitch" else pb.ALGO_SANKOFF
    ro = synth.unpack_nibbles(codes4[:1], C)[0].to(torch.int8).contiguous() if algo == pb.ALGO_SANKOFF else None
    for grid in [int(x) for x in a.grid.split(",")]:
        for K in [int(x) for x in a.lanes.split(",")]:
            ctxs = []
            for _ in range(K):
                c = pb.Context(0)
                if grid:
                    c.set_option("grid_pct", grid)
                c.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
                c.upload(C, tree.n_leaves, codes4, codes4.shape[1], pc, ro)
                c.run_resident(algo)
                ctxs.append(c)
            for i in range(3 * K):
                ctxs[i % K].run_resident_async(algo)
            for c in ctxs:
                c.wait()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(a.passes):
                ctxs[i % K].run_resident_async(algo)
            for c in ctxs:
                c.wait()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / a.passes
            alg = ctxs[0].algorithmic_bytes(algo)
            print(f"{a.config} {a.algo} grid_pct {grid or 'auto'} lanes {K}: {dt * 1e3:.4f} ms per pass, {alg / dt / 1e9:.0f} GB/s algorithmic", flush=True)
            for c in ctxs:
                c.close()


if __name__ == "__main__":
    main()
