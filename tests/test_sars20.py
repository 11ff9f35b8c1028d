"""BASELINE.json configs[0]: the reference's own test data (test/sars_20.json + test/sars_20.nwk, 20 SARS-CoV-2 genomes
as a PanGraph of 10 blocks; 31 137 nucleotide columns + 10 block columns) through every engine. Expected lists and
states come from the verbatim reference build (tests/golden/make_sars20_golden.py). Leaves whose path lacks a block
are omitted for that block's batch (leaf_present), main columns carry the root override -- the PanGraph column
drivers of src/panman.cpp:873-963 and :1048-1232."""
import numpy as np
import pytest

from tests.golden_util import load_sars20


@pytest.fixture(scope="module")
def sars20():
    return load_sars20()


def test_fixture_shape(sars20):
    tree, batches = sars20
    assert tree.n_leaves == 20 and tree.n_nodes == 39 and len(batches) == 11
    assert sum(b["codes"].shape[1] for b in batches[1:]) == 31137  # config 1
    assert batches[0]["codes"].shape == (20, 10)


@pytest.mark.parametrize("algo", [0, 1])
def test_port_matches_reference_on_sars20(port, sars20, algo):
    tree, batches = sars20
    for b in batches:
        want, want_states = b["expect"][algo]
        got, states = port.run(tree, algo, b["codes"], b["parent_code"], b["root_override"][algo], None, b["present"], b["block"],
                               n_threads=4, want_states=True)
        assert got.same_as(want), (b["id"], algo)
        assert np.array_equal(states, want_states), (b["id"], algo)


@pytest.mark.parametrize("algo", [0, 1])
def test_emulation_matches_reference_on_sars20(sars20, algo):
    from tests.emul.emul import Emulator

    emu = Emulator()
    tree, batches = sars20
    for b in batches:
        want, want_states = b["expect"][algo]
        rc, got, states, _ = emu.run(tree, algo, b["codes"], b["parent_code"], b["root_override"][algo], None, b["present"],
                                     b["block"], chunk_nodes=4, inline_nodes=1)
        assert rc == 0
        assert got.same_as(want), (b["id"], algo)
        assert np.array_equal(states, want_states), (b["id"], algo)


@pytest.mark.gpu
@pytest.mark.parametrize("algo", [0, 1])
def test_gpu_matches_reference_on_sars20(sars20, algo):
    import panman_b200 as pb

    tree, batches = sars20
    ctx = pb.Context(0)
    ctx.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
    total = 0
    for b in batches:
        want, want_states = b["expect"][algo]
        res = ctx.run_codes(tree, algo, b["codes"], b["parent_code"], b["root_override"][algo], None, b["present"], b["block"],
                            want_states=True)
        assert np.array_equal(res.node_offsets, want.node_offsets), (b["id"], algo)
        assert np.array_equal(res.pos, want.pos) and np.array_equal(res.type_code, want.type_code), (b["id"], algo)
        assert np.array_equal(res.states, want_states), (b["id"], algo)
        total += res.n_mut
    assert total > 3000
    ctx.close()
