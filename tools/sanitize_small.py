#!/usr/bin/env python
"""Small end-to-end cases, every kernel path once -- Fitch / Sankoff, plain / presence mask / block mode / states, chain
segments (speculating and waiting), level schedule, overflow retry, shard pack / merge, run-merge, the multi-rank group,
clade-run encoded input --
for compute-sanitizer where the pool allows it, and always under the library's own guard bytes: PMB_DEBUG_CANARY=1 puts 256
guard bytes around every device buffer and poisons its body, pmb_debug_check_canaries() counts the guard bytes a kernel
overwrote (out-of-bounds writes), and the comparison with the oracle catches reads of memory nobody initialised.
  python tools/sanitize_small.py > profiles/r02_canary_check.log"""
import ctypes
import os
import sys

os.environ.setdefault("PMB_DEBUG_CANARY", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import panman_b200 as pb  # noqa: E402
from oracle.oracle import PortOracle, random_tree  # noqa: E402


def main():
    rng = np.random.default_rng(3)
    port = PortOracle()
    ctx = pb.Context(0)
    n = 0
    for kind, leaves, chunk, sched, noise in (("binary", 150, 0, 1, 0.05), ("caterpillar", 300, 8, 1, 0.0), ("caterpillar", 200, 5, 1, 0.7),
                                              ("polytomy", 120, 3, 0, 0.1), ("unary", 90, 1, 1, 0.02)):
        tree = random_tree(leaves, 900 + n, kind, max_arity=5)
        n_cols = [1500, 1024, 33, 2100, 700][n % 5]
        base = rng.integers(0, 5, size=n_cols)
        codes = np.repeat(base[None, :], tree.n_leaves, 0)
        codes = np.where(rng.random(codes.shape) < noise, rng.integers(0, 16, size=codes.shape), codes).astype(np.uint8)
        pc = rng.integers(0, 16, size=n_cols).astype(np.uint8)
        ro = np.where(rng.random(n_cols) < 0.3, rng.integers(0, 16, size=n_cols), -1).astype(np.int8)
        lp = (rng.random(tree.n_leaves) < 0.8).astype(np.uint8)
        lp[0] = 1
        ctx.set_option("chunk_nodes", chunk)
        ctx.set_option("schedule", sched)
        ctx.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
        for algo in (0, 1):
            for present in (None, lp):
                want, ws = port.run(tree, algo, codes, pc, ro, None, present, 0, n_threads=2, want_states=True)
                res = ctx.run_codes(tree, algo, codes, pc, ro, None, present, 0, want_states=True)
                assert np.array_equal(res.pos, want.pos) and np.array_equal(res.type_code, want.type_code)
                assert np.array_equal(res.states, ws)
                # the same batch clade-run encoded (expand_runs_kernel instead of pack_leaves_kernel)
                c4 = pb.pack_nibbles(codes)
                runs = pb.Runs.of_tree(tree, n_cols, c4, pc)
                r2 = ctx.run_runs(algo, runs, pc, ro, None, present, flags=pb.FLAG_WANT_STATES)
                assert np.array_equal(r2.pos, want.pos) and np.array_equal(r2.type_code, want.type_code) and np.array_equal(r2.states, ws)
                runs.close()
        n += 1
    # overflow retry + async + pack/merge
    tree = random_tree(100, 5, "binary")
    codes = rng.integers(0, 16, size=(tree.n_leaves, 1500)).astype(np.uint8)
    ctx.set_option("chunk_nodes", 0)
    ctx.set_option("schedule", 1)
    ctx.set_option("staging_records", 500)
    ctx.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
    want, _ = port.run(tree, 0, codes, codes[0].copy(), n_threads=2)
    res = ctx.run_codes(tree, 0, codes, codes[0].copy())
    assert np.array_equal(res.pos, want.pos)
    ctx.set_option("staging_records", 0)
    c4 = pb.pack_nibbles(codes)
    ctx.upload(1500, tree.n_leaves, c4, c4.shape[1], codes[0].copy())
    ctx.run_resident_async(0)
    ctx.wait()
    cap = int(ctx.result_device().n_mut) + 10
    buf = torch.empty(2 * ctx.packed_bytes(cap), dtype=torch.uint8, device="cuda")
    for k in range(2):
        ctx.pack_result(buf[k * ctx.packed_bytes(cap):(k + 1) * ctx.packed_bytes(cap)], cap)
    ctx.merge_packed(2, buf, cap)
    ctx.merge_status()
    ctx.merge_runs(source=1)
    ctx.run_resident(0)
    ctx.merge_runs(source=0)
    torch.cuda.synchronize()
    lib = pb.load_library()
    lib.pmb_debug_check_canaries.restype = ctypes.c_longlong
    bad = lib.pmb_debug_check_canaries()
    print(f"single-context paths: guard bytes overwritten = {bad} (-1 = canaries off)")
    assert bad in (0, -1)
    ctx.close()
    # column-sharded group: three ranks on this GPU, asynchronous steps back to back (mailbox slots of both parities, the
    # stream-memory-operation hand-shake, the record-parallel merge), then the run-merge of the merged lists
    codes = rng.integers(0, 5, size=(tree.n_leaves, 3500)).astype(np.uint8)
    codes = np.where(rng.random(codes.shape) < 0.9, codes[:1], codes).astype(np.uint8)
    pc = codes[0].copy()
    c4 = pb.pack_nibbles(codes)
    for algo in (0, 1):
        want, _ = port.run(tree, algo, codes, pc, codes[0].astype(np.int8) if algo else None, None, None, 0, n_threads=2)
        g = pb.Group([0, 0, 0])
        g.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
        res = g.run_nuc(algo, 3500, tree.n_leaves, c4, c4.shape[1], pc, codes[0].astype(np.int8) if algo else None)
        assert np.array_equal(res.pos, want.pos) and np.array_equal(res.type_code, want.type_code)
        g.upload(3500, tree.n_leaves, c4, c4.shape[1], pc, codes[0].astype(np.int8) if algo else None)
        for _ in range(4):
            g.run_async(algo)
        g.wait()
        res = g.download()
        assert np.array_equal(res.pos, want.pos) and np.array_equal(res.type_code, want.type_code)
        g.merge_runs()
        runs = pb.Runs.of_tree(tree, 3500, c4, pc)
        res = g.run_runs(algo, runs, pc, codes[0].astype(np.int8) if algo else None)
        assert np.array_equal(res.pos, want.pos) and np.array_equal(res.type_code, want.type_code)
        runs.close()
        bad = ctypes.c_longlong(pb.load_library().pmb_debug_check_canaries()).value
        print(f"group algo {algo}: guard bytes overwritten = {bad}")
        assert bad in (0, -1)
        g.close()
    print("sanitize_small ok; PMB_DEBUG_CANARY =", os.environ.get("PMB_DEBUG_CANARY"))


if __name__ == "__main__":
    main()
