#!/usr/bin/env python
"""End-to-end timing of the -M construction flow through the host adaptor (libpanman_b200_host): FASTA text + Newick in,
Node::nucMutation fields out -- reader + consensus, packing, pmb_run_nuc (upload, pass, download), run-merge -- next to
the reference's own flow (verbatim fitchSankoff.cpp behind the restated string-keyed drivers) on a column sample.
  python tools/msa_flow_bench.py --leaves 5000 --cols 30000"""
import argparse
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import panman_b200 as pb  # noqa: E402
from oracle.oracle import CHAR_OF, RefOracle, have_ref  # noqa: E402
from panman_b200 import synth  # noqa: E402
from panman_b200.host import load_host_library  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--leaves", type=int, default=5000)
    ap.add_argument("--cols", type=int, default=30000)
    ap.add_argument("--low-mem", action="store_true")
    ap.add_argument("--ref-cols", type=int, default=256)
    args = ap.parse_args()
    tree = synth.make_tree(args.leaves, 2, "binary")
    codes4, pc = synth.simulate_msa(tree, 0, args.cols, synth.MsaSpec(2, 3e-5, 0.01, 1e-3), device="cuda")
    codes = synth.unpack_nibbles(codes4, args.cols).cpu().numpy()
    names = tree.names()
    leaf_names = [names[v] for v in tree.leaves]
    rows = CHAR_OF[codes]
    fasta = b"".join(b">" + n.encode() + b"\n" + bytes(r) + b"\n" for n, r in zip(leaf_names, rows))
    newick = tree.to_newick() if hasattr(tree, "to_newick") else None
    if newick is None:
        from oracle.oracle import FlatTree

        kids = [[] for _ in range(tree.n_nodes)]
        for v in range(tree.n_nodes):
            kids[v] = list(tree.child_idx[tree.child_off[v]:tree.child_off[v + 1]])
        newick = FlatTree.from_children(names, kids, tree.root).to_newick()
    L = load_host_library()
    ctx = pb.Context(0)
    err = C.create_string_buffer(512)
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        h = L.pmh_build_from_msa(ctx.h, fasta, len(fasta), newick.encode(), b"", int(args.low_mem), err, 512)
        dt = time.perf_counter() - t0
        if not h:
            raise SystemExit(err.value.decode())
        sec = list(np.ctypeslib.as_array(L.pmh_build_seconds(h), (4,)))
        n = L.pmh_build_n_tuples(h)
        L.pmh_build_free(h)
        if best is None or dt < best[0]:
            best = (dt, sec, n)
    dt, sec, n = best
    units = tree.n_nodes * args.cols
    print(f"-M flow, {args.leaves} leaves x {args.cols} columns ({len(fasta) / 1e6:.0f} MB FASTA), {n} records: total {dt * 1e3:.1f} ms = "
          f"{units / dt:.3e} node*col/s | reader+consensus {sec[0] * 1e3:.1f} ms, pack {sec[1] * 1e3:.1f} ms, "
          f"pmb_run_nuc {sec[2] * 1e3:.1f} ms, run-merge {sec[3] * 1e3:.1f} ms")
    if have_ref():
        ref = RefOracle()

        class T:
            pass

        t = T()
        t.n_nodes, t.names, t.parent, t.child_off, t.child_idx = tree.n_nodes, names, tree.parent, tree.child_off, tree.child_idx
        hh = ref.tree(t)
        nc = min(args.cols, args.ref_cols)
        threads = len(os.sched_getaffinity(0))
        t0 = time.perf_counter()
        ref.msa_run(hh, t, int(args.low_mem), leaf_names, [bytes(r[:nc]) for r in rows], bytes(CHAR_OF[pc.cpu().numpy()][:nc]), "",
                    n_threads=threads)
        rdt = time.perf_counter() - t0
        print(f"reference drivers on the first {nc} columns, {threads} threads: {rdt:.2f} s = {tree.n_nodes * nc / rdt:.3e} node*col/s "
              f"(passes only, no reader)")
    ctx.close()


if __name__ == "__main__":
    main()
