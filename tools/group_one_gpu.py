#!/usr/bin/env python
"""A column-sharded group with several ranks on ONE GPU (the hand-shake, mailbox and merge are those of a multi-GPU run):
for timing the gather / merge kernels under ncu and for checking the merged result against a single-context run.
  python tools/group_one_gpu.py --ranks 8 --config ecoli4k --cols 1250000 --steps 3"""
import argparse
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
    ap.add_argument("--ranks", type=int, default=8)
    ap.add_argument("--config", default="ecoli4k")
    ap.add_argument("--cols", type=int, default=1250000)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    cfg = dict(synth.CONFIGS[args.config])
    C = args.cols or cfg["n_cols"]
    tree = synth.make_tree(cfg["n_leaves"], cfg["seed"], cfg["kind"])
    codes4, pc = synth.simulate_msa(tree, 0, C, synth.MsaSpec(cfg["seed"], cfg["p_sub"], cfg["f_gap"], cfg["p_N"]), device="cuda")
    one = pb.Context(0)
    one.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
    one.upload(C, tree.n_leaves, codes4, codes4.shape[1], pc)
    one.run_resident(0)
    want = one.download()
    want_nm = one.merge_runs()
    one.close()
    g = pb.Group([0] * args.ranks)
    g.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
    g.reserve(int(want.n_mut) // args.ranks * 2 + 4096)
    g.upload(C, tree.n_leaves, codes4, codes4.shape[1], pc)
    for _ in range(args.steps):
        g.run_async(0)
    g.wait()
    res = g.download()
    ok = np.array_equal(res.node_offsets, want.node_offsets) and np.array_equal(res.pos, want.pos) and np.array_equal(res.type_code, want.type_code)
    nm = g.merge_runs()
    ok_nm = all(np.array_equal(a, b) for a, b in zip(nm, want_nm))
    print(f"{args.ranks} ranks on one GPU, {C} columns, {res.n_mut} records: merged lists identical = {ok}, NucMut fields identical = {ok_nm}")
    g.close()
    assert ok and ok_nm


if __name__ == "__main__":
    main()
