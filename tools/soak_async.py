#!/usr/bin/env python
"""Soak test of the multi-stream pass pipeline: many asynchronous passes back to back (forward i+1 beside backward i, the
compaction of i beside both, two ticket sets and two set matrices alternating), uploads of OTHER batches slipped in between
without waiting, several contexts on one GPU at once -- every result compared with a synchronous run of the same batch.
  python tools/soak_async.py [--seconds 120]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import panman_b200 as pb  # noqa: E402
from oracle.oracle import random_tree  # noqa: E402


def batch(rng, tree, n_cols, noise):
    base = rng.integers(0, 5, size=n_cols)
    codes = np.repeat(base[None, :], tree.n_leaves, 0)
    codes = np.where(rng.random(codes.shape) < noise, rng.integers(0, 16, size=codes.shape), codes).astype(np.uint8)
    return pb.pack_nibbles(codes), base.astype(np.uint8), codes[0].astype(np.int8)


def same(a, b):
    return np.array_equal(a.node_offsets, b.node_offsets) and np.array_equal(a.pos, b.pos) and np.array_equal(a.type_code, b.type_code)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    args = ap.parse_args()
    rng = np.random.default_rng(2024)
    t_end = time.time() + args.seconds
    rounds = passes = regrown = 0
    while time.time() < t_end:
        kind = ["binary", "caterpillar", "polytomy"][rounds % 3]
        tree = random_tree(int(rng.integers(50, 3000)), 7000 + rounds, kind, max_arity=5)
        ctxs = [pb.Context(0) for _ in range(int(rng.integers(1, 4)))]
        for c in ctxs:
            c.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
            if rounds % 4 == 3:
                c.set_option("chunk_nodes", int(rng.choice([2, 7, 40])))
        batches = [batch(rng, tree, int(rng.choice([700, 3000, 9000, 20000])), float(rng.choice([0.001, 0.05, 0.5]))) for _ in range(3)]
        want = {}
        for algo in (0, 1):
            for bi, (c4, pc, ro) in enumerate(batches):
                n_cols = len(pc)
                ctxs[0].upload(n_cols, tree.n_leaves, c4, c4.shape[1], pc, ro if algo else None)
                ctxs[0].run_resident(algo)
                want[(algo, bi)] = ctxs[0].download()
        for algo in (0, 1):
            order = rng.permutation(len(batches) * 3) % len(batches)
            for bi in order:
                c4, pc, ro = batches[bi]
                n_cols = len(pc)
                k = int(rng.integers(1, 12))
                for c in ctxs:  # every context: upload (no wait for the passes still in flight), k passes back to back
                    c.upload(n_cols, tree.n_leaves, c4, c4.shape[1], pc, ro if algo else None)
                    for _ in range(k):
                        c.run_resident_async(algo)
                    passes += k
                for c in ctxs:
                    try:
                        c.wait()
                    except pb.PanmanError as e:  # a dense batch outgrew the pool: it has been grown, run again (documented)
                        assert e.code == -8, e
                        regrown += 1
                        c.run_resident_async(algo)
                        c.wait()
                    got = c.download()
                    assert same(got, want[(algo, bi)]), (rounds, algo, bi, kind)
        for c in ctxs:
            c.close()
        rounds += 1
    # the multi-rank hand-shake under the same treatment: hundreds of steps in flight over both mailbox parities
    steps = 0
    for trial in range(4):
        tree = random_tree(int(rng.integers(100, 1500)), 7500 + trial, ["binary", "caterpillar"][trial % 2], max_arity=3)
        c4, pc, ro = batch(rng, tree, int(rng.choice([5000, 12000])), 0.02)
        one = pb.Context(0)
        one.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
        g = pb.Group([0] * int(rng.integers(2, 6)))
        g.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
        for algo in (0, 1):
            one.upload(len(pc), tree.n_leaves, c4, c4.shape[1], pc, ro if algo else None)
            one.run_resident(algo)
            w = one.download()
            got = g.run_nuc(algo, len(pc), tree.n_leaves, c4, c4.shape[1], pc, ro if algo else None)
            assert same(got, w), ("group run_nuc", trial, algo)
            g.upload(len(pc), tree.n_leaves, c4, c4.shape[1], pc, ro if algo else None)
            for _ in range(300):
                g.run_async(algo)
            steps += 300
            g.wait()
            assert same(g.download(), w), ("group", trial, algo)
        g.close()
        one.close()
    print(f"soak_async ok: {steps} group steps identical too; {rounds} rounds, {passes} asynchronous passes ({regrown} reruns after a staging-pool overflow), every result "
          "identical to the synchronous run")


if __name__ == "__main__":
    main()
