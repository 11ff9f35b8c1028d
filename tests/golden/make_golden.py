"""Generates tests/golden/fitch_sankoff_golden.npz from the VERBATIM reference build
(oracle/_ref/libpanman_ref.so = /root/reference/src/fitchSankoff.cpp compiled as is).

The reference ships no golden vectors, so these are outputs of the reference's own
code run in the build container:  python tests/golden/make_golden.py
Each case stores the inputs in the pmb_run_nuc convention and the expected per-node mutation lists and
assigned states. Cases cover: binary / polytomous / unary / caterpillar trees, all 16 codes, absent leaves,
root override (defaultState), forward root reference (refState), block (3-state) variants, single-leaf trees,
all-gap columns, and 1..200 columns (ragged against the 32-column packing granule).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import RefOracle, build, random_tree, ref_run_columns  # noqa: E402


def make_cases():
    rng = np.random.default_rng(20261018)
    cases = []
    kinds = ["binary", "polytomy", "unary", "caterpillar"]
    col_counts = [1, 31, 32, 33, 64, 100, 200, 7]
    for i in range(32):
        kind = kinds[i % 4]
        n_leaves = [1, 2, 3, 5, 17, 40, 64, 90][(i // 4) % 8]
        tree = random_tree(n_leaves, 1000 + i, kind, max_arity=6)
        n_cols = col_counts[i % 8]
        algo = (i // 2) % 2
        block = 1 if i % 7 == 3 else 0
        nst = 3 if block else 16
        alpha = [nst, min(nst, 5), 2, 1][i % 4]
        # mostly-conserved columns with sparse changes, like real alignments; plus fully random ones
        base = rng.integers(0, alpha, size=n_cols)
        codes = np.repeat(base[None, :], tree.n_leaves, 0)
        noise = rng.random(codes.shape) < [0.02, 0.2, 0.5, 1.0][(i // 3) % 4]
        codes = np.where(noise, rng.integers(0, nst, size=codes.shape), codes).astype(np.uint8)
        if i % 5 == 0:
            codes[:, 0] = 0  # an all-gap column
        parent_code = np.where(rng.random(n_cols) < 0.7, codes[0], rng.integers(0, nst, size=n_cols)).astype(np.uint8)
        root_override = np.full(n_cols, -1, np.int8)
        fwd_root_ref = None
        leaf_present = None
        if i % 3 == 1:
            root_override = np.where(rng.random(n_cols) < 0.5, rng.integers(0, nst, size=n_cols), -1).astype(np.int8)
        if algo == 0 and not block and i % 4 == 2:
            fwd_root_ref = np.where(rng.random(n_cols) < 0.5, rng.integers(0, nst, size=n_cols), -1).astype(np.int8)
        if i % 6 == 4 and tree.n_leaves > 1:
            leaf_present = (rng.random(tree.n_leaves) < 0.7).astype(np.uint8)
            leaf_present[int(rng.integers(0, tree.n_leaves))] = 1  # Sankoff root must stay defined
        cases.append(dict(newick=tree.to_newick(), tree=tree, algo=algo, block=block, codes=codes,
                          parent_code=parent_code, root_override=root_override, fwd_root_ref=fwd_root_ref,
                          leaf_present=leaf_present))
    return cases


def main():
    build()
    ref = RefOracle()
    out = {}
    cases = make_cases()
    for k, c in enumerate(cases):
        muts, states = ref_run_columns(ref, c["tree"], c["algo"], c["codes"], c["parent_code"], c["root_override"],
                                       c["fwd_root_ref"], c["leaf_present"], c["block"])
        p = f"c{k:02d}_"
        out[p + "newick"] = np.frombuffer(c["newick"].encode(), np.uint8)
        out[p + "algo_block"] = np.asarray([c["algo"], c["block"]], np.int32)
        out[p + "codes"] = c["codes"]
        out[p + "parent_code"] = c["parent_code"]
        out[p + "root_override"] = c["root_override"]
        if c["fwd_root_ref"] is not None:
            out[p + "fwd_root_ref"] = c["fwd_root_ref"]
        if c["leaf_present"] is not None:
            out[p + "leaf_present"] = c["leaf_present"]
        out[p + "node_offsets"] = muts.node_offsets
        out[p + "pos"] = muts.pos
        out[p + "type_code"] = muts.type_code
        out[p + "states"] = states
    out["n_cases"] = np.asarray([len(cases)], np.int32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fitch_sankoff_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(cases), "cases;",
          sum(int(out[f"c{k:02d}_pos"].size) for k in range(len(cases))), "records")


if __name__ == "__main__":
    main()
