"""Debug helper: run one trial of tests/test_gpu_parity.py::test_random_vs_oracle in isolation.
usage: python tools/debug_trials.py ALGO TRIAL [schedule] [chunk_nodes] [inline_nodes]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import panman_b200 as pb  # noqa: E402
from oracle.oracle import PortOracle, random_tree  # noqa: E402

algo, want_trial = int(sys.argv[1]), int(sys.argv[2])
over = [int(x) for x in sys.argv[3:]]
rng = np.random.default_rng(100 + algo)
port = PortOracle()
for trial in range(24):
    kind = ["binary", "polytomy", "unary", "caterpillar"][trial % 4]
    tree = random_tree(int(rng.integers(1, 400)), 3000 + trial, kind, max_arity=[3, 6, 20, 300][(trial // 4) % 4])
    n_cols = int(rng.choice([1, 33, 1000, 1024, 1025, 5000]))
    block = int(trial % 5 == 4)
    nst = 3 if block else 16
    base = rng.integers(0, min(nst, 5), size=n_cols)
    codes = np.repeat(base[None, :], tree.n_leaves, 0)
    noise = rng.random(codes.shape) < [0.01, 0.1, 0.6][trial % 3]
    codes = np.where(noise, rng.integers(0, nst, size=codes.shape), codes).astype(np.uint8)
    pc = rng.integers(0, nst, size=n_cols).astype(np.uint8)
    ro = np.where(rng.random(n_cols) < 0.3, rng.integers(0, nst, size=n_cols), -1).astype(np.int8) if trial % 2 else None
    fr = np.where(rng.random(n_cols) < 0.3, rng.integers(0, nst, size=n_cols), -1).astype(np.int8) if (
        algo == 0 and not block and trial % 3 == 0) else None
    lp = None
    if trial % 4 == 1 and tree.n_leaves > 1:
        lp = (rng.random(tree.n_leaves) < 0.7).astype(np.uint8)
        lp[0] = 1
    ck = int(rng.choice([0, 1, 7, 64]))
    inl = int(rng.choice([0, 2, 3, 9]))
    sch = int(trial % 3 != 0)
    if trial != want_trial:
        continue
    if len(over) > 0:
        sch = over[0]
    if len(over) > 1:
        ck = over[1]
    if len(over) > 2:
        inl = over[2]
    ws = len(over) > 3 and over[3] == 0
    want, want_states = port.run(tree, algo, codes, pc, ro, fr, lp, block, n_threads=4, want_states=True)
    ctx = pb.Context(0)
    ctx.set_option("chunk_nodes", ck)
    ctx.set_option("inline_nodes", inl)
    ctx.set_option("schedule", sch)
    ctx.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
    tag = f"algo {algo} trial {trial} {kind} nodes {tree.n_nodes} cols {n_cols} chunk {ck} inline {inl} sched {sch} block {block} lp {lp is not None} states {not ws}"
    try:
        res = ctx.run_codes(tree, algo, codes, pc, ro, fr, lp, block, want_states=not ws, col_base=11)
    except pb.PanmanError as e:
        print("FAIL", tag, "|", str(e)[:90])
        sys.exit(1)
    want.pos = want.pos + 11
    same = np.array_equal(res.node_offsets, want.node_offsets) and np.array_equal(res.pos, want.pos) and np.array_equal(
        res.type_code, want.type_code) and (ws or np.array_equal(res.states, want_states))
    print("OK  " if same else "DIFF", tag)
