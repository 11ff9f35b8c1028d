"""The C++ PanGraph adaptor (panman_b200/host/pangraph.cpp, include/panman_b200_host.h): its JSON reader and per-block
column batches against the Python restatement used for the golden fixture (tests/pangraph_util.py) on random PanGraphs --
and on the reference's own test/sars_20.json where /root/reference exists -- then, on a GPU, the whole -P flow against the
oracle on the same batches."""
import json
import os

import numpy as np
import pytest

from oracle.oracle import random_tree
from tests.pangraph_util import build_batches, random_pangraph


def _same_batches(got, want_states, want_batches):
    assert np.array_equal(got.block_states, want_states)
    assert len(got.batches) == len(want_batches)
    for g, w in zip(got.batches, want_batches):
        for k in ("codes", "present", "parent_code", "root_override", "col_j", "col_k"):
            assert np.array_equal(g[k], w[k]), (g["id"], k)


@pytest.mark.parametrize("duplicates,circular,shuffle", [(False, False, False), (True, False, False), (False, True, False),
                                                         (False, False, True), (True, True, True)])
def test_adaptor_batches_match_restatement(refpg, duplicates, circular, shuffle):
    """Block columns (the order chain_align gives, duplicated blocks, rotated circular paths: against the reference's own
    compiled chaining.cpp / rotation.cpp) and the per-column batches (against the Python restatement on that order)."""
    from panman_b200.host import PanGraphBuild

    rng = np.random.default_rng(8 + 4 * duplicates + 2 * circular + shuffle)
    for trial in range(12):
        tree = random_tree(int(rng.integers(2, 40)), 8100 + trial, ["binary", "polytomy", "caterpillar"][trial % 3], max_arity=4)
        text = random_pangraph(tree, rng, n_blocks=int(rng.integers(1, 7)), max_len=int(rng.choice([8, 60, 300])),
                               duplicates=duplicates, circular=circular, shuffle=shuffle)
        pg = json.loads(text)
        order = refpg.order(pg)
        want_states, want_batches = build_batches(pg, tree, order)
        got = PanGraphBuild(text.encode(), tree.to_newick())
        assert got.tree.names == tree.names
        assert got.block_ids == order["topo_ids"], (trial, "block columns")
        rows = {tree.names[v]: int(tree.leaf_row[v]) for v in tree.leaves}
        for name, rot in order["rotation_index"].items():
            assert got.rotation_index[rows[name]] == rot, (trial, name)
        _same_batches(got, want_states, want_batches)
        got.close()


@pytest.mark.parametrize("reference", ["one", "many"])
def test_adaptor_reference_overrides(refpg, reference):
    """--reference: the root is forced to the reference sequence's state on every block and nucleotide column; where the
    string matches several sequence names, to the one the reference's map walks reach last."""
    from panman_b200.host import PanGraphBuild

    rng = np.random.default_rng(77)
    for trial in range(8):
        tree = random_tree(int(rng.integers(3, 30)), 8300 + trial, ["binary", "polytomy"][trial % 2], max_arity=4)
        text = random_pangraph(tree, rng, n_blocks=int(rng.integers(2, 6)), max_len=40, duplicates=trial % 3 == 0)
        pg = json.loads(text)
        names = [tree.names[v] for v in tree.leaves]
        ref = names[int(rng.integers(0, len(names)))] if reference == "one" else names[0][:1]  # the common first letter
        order = refpg.order(pg)
        want_states, want_batches, want_block_override = build_batches(pg, tree, order, ref)
        got = PanGraphBuild(text.encode(), tree.to_newick(), ref)
        _same_batches(got, want_states, want_batches)
        assert np.array_equal(got.block_override, want_block_override), trial
        got.close()


def test_adaptor_reports_malformed_input():
    """Nothing may abort the host process: malformed JSON, wrong kinds and sizes, unknown blocks come back as errors."""
    from panman_b200.host import PanGraphBuild

    tree = random_tree(3, 1, "binary")
    pg = json.loads(random_pangraph(tree, np.random.default_rng(1), n_blocks=2))
    with pytest.raises(ValueError, match="JSON"):
        PanGraphBuild(b'{"paths": [', tree.to_newick())
    with pytest.raises(ValueError, match="JSON"):
        PanGraphBuild(b"[" * 100 + b"]" * 100, tree.to_newick())
    bad = json.loads(json.dumps(pg))
    bad["blocks"][0]["gaps"] = {"x7": 2}
    with pytest.raises(ValueError, match="gap"):
        PanGraphBuild(json.dumps(bad).encode(), tree.to_newick())
    bad = json.loads(json.dumps(pg))
    bad["blocks"][0]["mutate"] = [[{"name": tree.names[int(tree.leaves[0])], "number": 1}, [[3]]]]
    with pytest.raises(ValueError, match="malformed"):
        PanGraphBuild(json.dumps(bad).encode(), tree.to_newick())
    bad = json.loads(json.dumps(pg))
    bad["blocks"][0]["insert"] = [[{"name": tree.names[int(tree.leaves[0])], "number": 1}, [[5, "ACG"]]]]
    with pytest.raises(ValueError, match="malformed"):
        PanGraphBuild(json.dumps(bad).encode(), tree.to_newick())
    bad = json.loads(json.dumps(pg))
    bad["paths"][0]["blocks"][0]["id"] = "NOPE"
    with pytest.raises(ValueError, match="unknown block"):
        PanGraphBuild(json.dumps(bad).encode(), tree.to_newick())


@pytest.mark.skipif(not os.path.exists("/root/reference/test/sars_20.json"), reason="the reference's test data is not on this machine")
def test_adaptor_on_sars20_matches_fixture():
    from panman_b200.host import PanGraphBuild
    from tests.golden_util import load_sars20

    tree, batches = load_sars20()
    text = open("/root/reference/test/sars_20.json", "rb").read()
    newick = open("/root/reference/test/sars_20.nwk").readline().strip()
    got = PanGraphBuild(text, newick)
    assert got.tree.names == tree.names
    assert np.array_equal(got.block_states, batches[0]["codes"])
    for g, w in zip(got.batches, batches[1:]):
        assert np.array_equal(g["codes"], w["codes"]) and np.array_equal(g["present"], w["present"])
        assert np.array_equal(g["parent_code"], w["parent_code"]) and np.array_equal(g["root_override"], w["root_override"][0])
        assert np.array_equal(g["col_j"], w["col_j"]) and np.array_equal(g["col_k"], w["col_k"])
    got.close()


@pytest.mark.gpu
@pytest.mark.parametrize("algo", [0, 1])
def test_pangraph_flow_matches_oracle(port, algo):
    import panman_b200 as pb
    from panman_b200.host import PanGraphBuild

    rng = np.random.default_rng(80 + algo)
    ctx = pb.Context(0)
    for trial in range(5):
        tree = random_tree(int(rng.integers(3, 60)), 8200 + trial, ["binary", "polytomy", "caterpillar"][trial % 3], max_arity=4)
        text = random_pangraph(tree, rng, n_blocks=int(rng.integers(1, 5)), max_len=int(rng.choice([40, 300, 1500])),
                               duplicates=trial % 2 == 1, circular=trial == 3, shuffle=trial >= 3)
        build = PanGraphBuild(text.encode(), tree.to_newick())
        results = build.run(ctx, algo)
        nb = build.n_blocks
        want, _ = port.run(tree, algo, build.block_states, np.zeros(nb, np.uint8), None, None, None, 1, n_threads=2)
        off, pos, tc = results[0]
        assert np.array_equal(off, want.node_offsets) and np.array_equal(pos, want.pos) and np.array_equal(tc, want.type_code)
        for bt, (off, pos, tc) in zip(build.batches, results[1:]):
            ro = bt["root_override"] if algo == 0 else None
            want, _ = port.run(tree, algo, bt["codes"], bt["parent_code"], ro, None, bt["present"], 0, n_threads=2)
            assert np.array_equal(off, want.node_offsets), (trial, bt["id"])
            assert np.array_equal(pos, want.pos) and np.array_equal(tc, want.type_code), (trial, bt["id"])
        # Node::nucMutation: the oracle's run-merge of the 6-tuples (block, -1, pos, gapPos, type, code), non-gap list then gap list
        got = build.nucmut()
        for v in range(tree.n_nodes):
            tup = []
            for b, (bt, (off, pos, tc)) in enumerate(zip(build.batches, results[1:])):
                for k in range(off[v], off[v + 1]):
                    tup.append((b, int(bt["col_j"][pos[k]]), int(bt["col_k"][pos[k]]), int(tc[k])))
            want_v = []
            for gap in (0, 1):
                sel = sorted(t for t in tup if (t[2] >= 0) == bool(gap))
                if not sel:
                    continue
                ob, op_, og, mi, nu = port.merge_pangraph(gap, [t[0] for t in sel], [t[1] for t in sel], [t[2] for t in sel],
                                                          np.asarray([t[3] for t in sel], np.uint8))
                want_v += [(int(op_[i]), int(og[i]), int(ob[i]), -1, int(mi[i]), int(nu[i])) for i in range(len(ob))]
            assert got[v] == want_v, (trial, v)
        # the Mutation lists the writer would store (src/panman.cpp:2854-2929): grouped by block, wire mutInfo, block mutations
        from tests.golden_util import writer_layout

        bm = build.blockmut()
        for v, w in enumerate(build.wire()):
            assert w == writer_layout(got[v], bm[v]), (trial, v)
        build.close()
    ctx.close()
