"""Generates tests/golden/sars20_pangraph.npz: BASELINE.json configs[0] (test/sars_20.json + test/sars_20.nwk of the
reference: 20 SARS-CoV-2 genomes as a PanGraph of 10 blocks) turned into per-block column batches in the pmb_run_nuc
convention, with the expected per-node mutation lists and assigned states computed by the VERBATIM reference build
(oracle/_ref/libpanman_ref.so, one column at a time through oracle.ref_run_columns).

  python tests/golden/make_sars20_golden.py        (needs /root/reference; the .npz is committed, this script made it)

What is restated here (test infrastructure, not product): how the reference's PanGraph path builds the per-sequence
aligned strings of a block -- src/panman.cpp:6216-6258 (JSON fields: consensus `sequence`, `gaps` {pos: len},
`mutate` [pos(1-based), char], `insert` [[pos, offset], string], `delete` [pos(1-based), len]) and :1006-1045
(consensus + '-' filled gap slots; substitutions, insertions into gap slots, deletions) -- and the column drivers
:1048-1232: one column per consensus position j in [0, len] ("main", parent state = consensus char, '-' at j = len)
and one per gap slot (j, k) (parent state '-'); sequences whose path lacks the block are OMITTED (missing-leaf
semantics), sequences that have it contribute their character ('-' = code 0).

Block columns are in the reference's own order: oracle.RefPgOrder runs the reference's chaining.cpp / rotation.cpp
(compiled verbatim into oracle/_ref/libpanman_pgorder.so) the way Pangraph::Pangraph drives them (src/panman.cpp:6259-6465).
One rule of that path is pinned by no executable reference code: without --reference the Fitch main-column branch forces
the root to the state of whichever owner of the block a tbb::concurrent_unordered_map walks last (src/panman.cpp:1131-1138;
the `reference.length()` guard of :1057 is missing at :1132). TBB is not installed here; the walk order is restated in
tests/pangraph_util.py (tbb_walk_key: ascending bit-reversed tbb_hasher of the sequence name). Gap columns (guarded) have no
override.
The tree is bifurcating, so the reference would run Fitch; the Sankoff lists (its polytomy branch, with the override
rule of that branch: none without --reference) are stored as well.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import RefOracle, RefPgOrder, build, parse_newick, ref_run_columns  # noqa: E402
from tests.pangraph_util import build_batches  # noqa: E402

REF_TEST = "/root/reference/test"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sars20_pangraph.npz")


def main():
    build()
    ref = RefOracle()
    newick = open(os.path.join(REF_TEST, "sars_20.nwk")).readline().strip()
    tree = parse_newick(newick)
    pg = json.load(open(os.path.join(REF_TEST, "sars_20.json")))
    order = RefPgOrder().order(pg)
    print("block columns:", order["topo_ids"])
    bcodes, batches = build_batches(pg, tree, order)
    out = dict(newick=np.frombuffer(newick.encode(), np.uint8), n_blocks=np.asarray([len(batches)], np.int32),
               block_ids=np.frombuffer("\n".join(order["topo_ids"]).encode(), np.uint8))
    # ---- block level: one 3-state column per block (0 absent, 1 forward, 2 reverse); parent state absent; no --reference
    out["blk_codes"] = bcodes
    for algo in (0, 1):
        want, states = ref_run_columns(ref, tree, algo, bcodes, np.zeros(bcodes.shape[1], np.uint8), None, None, None, 1)
        out[f"blk_a{algo}_off"], out[f"blk_a{algo}_pos"], out[f"blk_a{algo}_tc"] = want.node_offsets, want.pos, want.type_code
        out[f"blk_a{algo}_states"] = states
    # ---- nucleotide level, block by block
    total_cols = 0
    for i, bt in enumerate(batches):
        p = f"b{i}_"
        codes, present, parent_code, root_override = bt["codes"], bt["present"], bt["parent_code"], bt["root_override"]
        out[p + "codes"], out[p + "present"], out[p + "parent_code"], out[p + "root_override"] = codes, present, parent_code, root_override
        out[p + "col_j"], out[p + "col_k"] = bt["col_j"], bt["col_k"]
        for algo in (0, 1):
            ro = root_override if algo == 0 else None
            want, states = ref_run_columns(ref, tree, algo, codes, parent_code, ro, None, present, 0)
            out[p + f"a{algo}_off"], out[p + f"a{algo}_pos"], out[p + f"a{algo}_tc"] = want.node_offsets, want.pos, want.type_code
            out[p + f"a{algo}_states"] = states
            print(f"block {i} {bt['id']}: {int(present.sum())} sequences x {codes.shape[1]} columns, algo {algo}: {len(want.pos)} records")
        total_cols += codes.shape[1]
    print("nucleotide columns in total:", total_cols)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
