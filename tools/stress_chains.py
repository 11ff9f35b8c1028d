#!/usr/bin/env python
"""Stress test of the chain-segment paths on one GPU: random deep trees, small chunks, mixed noise; every result is
compared with the oracle port. Prints the first mismatch / error and exits non-zero.
  python tools/stress_chains.py --seconds 60 --seed 1"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import panman_b200 as pb  # noqa: E402
from oracle.oracle import PortOracle, random_tree  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    port = PortOracle()
    ctx = pb.Context(0)
    t0 = time.time()
    n = 0
    while time.time() - t0 < args.seconds:
        kind = ["caterpillar", "unary", "binary"][n % 3]
        tree = random_tree(int(rng.integers(50, 4000)), int(rng.integers(1 << 30)), kind, max_arity=3)
        n_cols = int(rng.choice([700, 2048, 4100, 9000]))
        noise = float(rng.choice([0.0, 0.005, 0.02, 0.2, 0.7]))
        base = rng.integers(0, 5, size=n_cols)
        codes = np.repeat(base[None, :], tree.n_leaves, 0)
        codes = np.where(rng.random(codes.shape) < noise, rng.integers(0, 16, size=codes.shape), codes).astype(np.uint8)
        pc = rng.integers(0, 16, size=n_cols).astype(np.uint8)
        ro = np.where(rng.random(n_cols) < 0.3, rng.integers(0, 16, size=n_cols), -1).astype(np.int8) if n % 2 else None
        opts = dict(chunk_nodes=int(rng.choice([2, 5, 16, 64, 0])), inline_nodes=int(rng.choice([0, 1, 3])),
                    schedule=int(rng.random() < 0.8), bwd_tail=int(rng.choice([0, 20])))
        for k, v in opts.items():
            ctx.set_option(k, v)
        ctx.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
        for algo in (0, 1):
            want, _ = port.run(tree, algo, codes, pc, ro, None, None, 0, n_threads=8)
            try:
                res = ctx.run_codes(tree, algo, codes, pc, ro, None, None, 0)
            except pb.PanmanError as e:
                print("ERROR", e, dict(n=n, kind=kind, leaves=tree.n_leaves, cols=n_cols, noise=noise, algo=algo, **opts), flush=True)
                sys.exit(1)
            ok = (np.array_equal(res.node_offsets, want.node_offsets) and np.array_equal(res.pos, want.pos)
                  and np.array_equal(res.type_code, want.type_code))
            if not ok:
                print("MISMATCH", dict(n=n, kind=kind, leaves=tree.n_leaves, cols=n_cols, noise=noise, algo=algo, **opts), flush=True)
                sys.exit(2)
        n += 1
    print(f"stress ok: {n} cases in {time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    main()
