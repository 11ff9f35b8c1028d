"""Tree::reroot on the library (f-4; reference src/reroot.cpp:4-261).

* the topology transform (Tree::transform / transformHelper, src/panman.cpp:5831-5906) of the C++ host library against an
  independent recursive Python restatement, on random trees, compared by NAME (child lists in order);
* on a GPU: pmh_pangraph_reroot -- block columns and nucleotide columns re-inferred with the new root forced -- against the
  reference's own passes (the verbatim fitchSankoff.cpp build, or the port pinned to it) driven as src/reroot.cpp drives
  them, and the merged NucMut / BlockMut fields against the reference's own struct."""
import json

import numpy as np
import pytest

from oracle.oracle import FlatTree, block_mut_from_nuc, random_tree
from tests.pangraph_util import random_pangraph


def _py_transform(tree: FlatTree, tip: int):
    """Literal restatement with per-node child lists and parent pointers (recursive, like the reference). Returns
    (root name, {name: [child names]})."""
    kids = {v: [int(c) for c in tree.child_idx[tree.child_off[v]:tree.child_off[v + 1]]] for v in range(tree.n_nodes)}
    parent = {v: int(tree.parent[v]) for v in range(tree.n_nodes)}
    names = dict(enumerate(tree.names))
    root = tree.root
    dead = set()

    def helper(node):  # transformHelper, src/panman.cpp:5831-5865
        if node == root:
            if len(kids[node]) > 1:
                return node
            only = kids[node][0]
            dead.add(node)
            return only
        par = parent[node]
        kids[par].remove(node)
        parent[node] = -1
        new_child = helper(par)
        kids[node].append(new_child)
        parent[new_child] = node
        return node

    par = parent[tip]
    if par == -1 or par == root:  # :5868-5877
        return names[root], {names[v]: [names[c] for c in kids[v]] for v in kids}
    kids[par].remove(tip)
    parent[tip] = -1
    n_internal = sum(1 for v in range(tree.n_nodes) if tree.child_off[v + 1] > tree.child_off[v])
    new_root = tree.n_nodes
    names[new_root] = f"node_{n_internal + 1}"
    kids[new_root] = [tip]
    parent[tip] = new_root
    sib = helper(par)
    kids[new_root].append(sib)
    parent[sib] = new_root
    return names[new_root], {names[v]: [names[c] for c in kids[v]] for v in kids if v not in dead}


def _by_name(t):
    return t.names[t.root], {t.names[v]: [t.names[c] for c in t.child_idx[t.child_off[v]:t.child_off[v + 1]]] for v in range(t.n_nodes)}


def test_transform_matches_restatement():
    import sys

    from panman_b200.host import reroot_newick

    sys.setrecursionlimit(20000)
    rng = np.random.default_rng(54)
    for trial in range(60):
        kind = ["binary", "polytomy", "caterpillar", "unary"][trial % 4]
        tree = random_tree(int(rng.integers(2, 120)), 5400 + trial, kind, max_arity=5)
        tip = int(rng.choice(tree.leaves))
        if tree.child_off[tree.root + 1] - tree.child_off[tree.root] == 1 and tree.parent[tip] != tree.root:
            # a unary root loses its only child: the reference reads children[0] of an empty vector (src/panman.cpp:5836);
            # the library reports it
            with pytest.raises(ValueError, match="no other child"):
                reroot_newick(tree.to_newick(), tree.names[tip])
            continue
        want_root, want = _py_transform(tree, tip)
        got = reroot_newick(tree.to_newick(), tree.names[tip])
        got_root, got_kids = _by_name(got)
        assert got_root == want_root, (trial, kind)
        assert got_kids == want, (trial, kind)
        # ids are a pre-order walk, leaf rows travel with the names
        assert got.root == 0 and all(got.parent[v] < v for v in range(1, got.n_nodes))
        rows = {tree.names[v]: int(tree.leaf_row[v]) for v in tree.leaves}
        assert {got.names[v]: int(got.leaf_row[v]) for v in range(got.n_nodes) if got.leaf_row[v] >= 0} == rows


def test_transform_errors():
    from panman_b200.host import reroot_newick

    with pytest.raises(ValueError, match="not found"):
        reroot_newick("((a,b),c);", "zzz")
    with pytest.raises(ValueError, match="not a tip"):
        reroot_newick("((a,b),c);", "node_2")
    t = reroot_newick("((a,b),c);", "c")  # the root's child: nothing changes (src/panman.cpp:5873-5877)
    assert t.names == ["node_1", "node_2", "a", "b", "c"]
    t = reroot_newick("((a,b),c);", "a")  # the old root keeps one child and disappears
    assert _by_name(t) == ("node_3", {"node_3": ["a", "node_2"], "a": [], "node_2": ["b", "c"], "b": [], "c": []})


def _flat(host_tree) -> FlatTree:
    t = FlatTree(list(host_tree.names), host_tree.parent.astype(np.int32), host_tree.child_off.astype(np.int32),
                 host_tree.child_idx.astype(np.int32), int(host_tree.root), host_tree.leaf_row.astype(np.int32))
    return t


def _replay(tree, off, pos, tc, parent_code):
    """What getSequenceFromReference yields per leaf row for one block: consensus + the mutations on the path root -> tip."""
    rows = np.zeros((tree.n_leaves, len(parent_code)), np.uint8)
    cur = {tree.root: parent_code.copy()}
    for v in range(tree.n_nodes):  # creation order: parents before children
        base = cur[v] if v == tree.root else cur[int(tree.parent[v])].copy()
        a, b = off[v], off[v + 1]
        base[pos[a:b]] = tc[a:b] & 15
        cur[v] = base
        if tree.leaf_row[v] >= 0:
            rows[tree.leaf_row[v]] = base
    return rows


@pytest.mark.gpu
def test_pangraph_reroot_matches_reference_passes(port, refnm):
    import panman_b200 as pb
    from oracle.oracle import RefOracle, have_ref, ref_run_columns
    from panman_b200.host import PanGraphBuild

    ref = RefOracle() if have_ref() else None
    rng = np.random.default_rng(90)
    ctx = pb.Context(0)
    for trial in range(6):
        tree = random_tree(int(rng.integers(4, 40)), 9100 + trial, ["binary", "polytomy", "caterpillar"][trial % 3], max_arity=4)
        text = random_pangraph(tree, rng, n_blocks=int(rng.integers(1, 5)), max_len=int(rng.choice([30, 120, 400])))
        build = PanGraphBuild(text.encode(), tree.to_newick())
        old = _flat(build.tree)
        built = build.run(ctx, 1 if old.has_polytomy() else 0)
        tip = int(rng.choice(old.leaves))
        row = int(old.leaf_row[tip])
        results = build.reroot(ctx, old.names[tip])
        new = _flat(build.tree)
        assert _by_name(new) == _py_transform(old, tip)
        # block columns: states of every leaf, root forced to the new root's state (src/reroot.cpp:54-122)
        nb = build.n_blocks
        bro = build.block_states[row].astype(np.int8)
        want, _ = port.run(new, 0, build.block_states, np.zeros(nb, np.uint8), bro, None, None, 1, n_threads=2)
        off, pos, tc = results[0]
        assert np.array_equal(off, want.node_offsets) and np.array_equal(pos, want.pos) and np.array_equal(tc, want.type_code)
        if ref is not None:
            wr, _ = ref_run_columns(ref, new, 0, build.block_states, np.zeros(nb, np.uint8), bro, None, None, 1)
            assert np.array_equal(off, wr.node_offsets) and np.array_equal(pos, wr.pos) and np.array_equal(tc, wr.type_code)
        got_bm = build.blockmut()
        for v in range(new.n_nodes):
            a, b = off[v], off[v + 1]
            bt, inv = block_mut_from_nuc(tc[a:b] >> 4, tc[a:b] & 15)  # the reference's (BlockMutationType, inversion) pair
            assert got_bm[v] == [(int(pos[a + i]), -1, int(bt[i] == 1), int(inv[i])) for i in range(b - a)]
        # nucleotide columns: every leaf with the characters the built PanMAT yields, root forced (src/reroot.cpp:134-224)
        for bt, (boff, bpos, btc), (off, pos, tc) in zip(build.batches, built[1:], results[1:]):
            rows = _replay(old, boff, bpos, btc, bt["parent_code"])
            present = bt["present"].astype(bool)
            assert np.array_equal(rows[present], bt["codes"][present])  # the replay property of the -P build
            ro = rows[row].astype(np.int8)
            want, _ = port.run(new, 0, rows, bt["parent_code"], ro, None, None, 0, n_threads=2)
            assert np.array_equal(off, want.node_offsets), (trial, bt["id"])
            assert np.array_equal(pos, want.pos) and np.array_equal(tc, want.type_code), (trial, bt["id"])
            if ref is not None and rows.shape[1] <= 200:
                wr, _ = ref_run_columns(ref, new, 0, rows, bt["parent_code"], ro, None, None, 0)
                assert np.array_equal(off, wr.node_offsets) and np.array_equal(pos, wr.pos) and np.array_equal(tc, wr.type_code)
        # Node::nucMutation against the reference's own NucMut struct driven by the loops of src/reroot.cpp:226-261
        got = build.nucmut()
        for v in range(new.n_nodes):
            tup = []
            for b, (bt, (off, pos, tc)) in enumerate(zip(build.batches, results[1:])):
                for k in range(off[v], off[v + 1]):
                    tup.append((b, int(bt["col_j"][pos[k]]), int(bt["col_k"][pos[k]]), int(tc[k]) >> 4, int(tc[k]) & 15))
            want_v = []
            for gap in (0, 1):
                sel = [t for t in tup if (t[2] >= 0) == bool(gap)]
                if not sel:
                    continue
                rb, rsb, rp, rg, rinfo, rnucs = refnm.merge_pangraph(gap, *[[t[i] for t in sel] for i in range(5)])
                want_v += [(int(rp[i]), int(rg[i]), int(rb[i]), int(rsb[i]), int(rinfo[i]), int(rnucs[i])) for i in range(len(rb))]
            assert got[v] == want_v, (trial, v)
        build.close()
    ctx.close()
