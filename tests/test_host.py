"""Host adaptor (C++, include/panman_b200_host.h): Newick conventions on the CPU, the -M construction flow on the GPU."""
import numpy as np
import pytest

from tests.golden_util import writer_layout

from oracle.oracle import CODE_OF, parse_newick as py_parse_newick, random_tree


def test_cpp_newick_matches_restated_conventions():
    import panman_b200 as pb

    pb.build_library()
    from panman_b200.host import parse_newick

    cases = ["((A:0.1,'B,x (1)':2):0.5,(C,D,E)0.9:1,F);", "(A,B);", "(A);", "((A,B),(C,(D,E)));\n", "  ((a:1,b:2):3,c:4);  "]
    for kind in ("binary", "polytomy", "unary", "caterpillar"):
        cases.append(random_tree(60, 5, kind, max_arity=5).to_newick())
    cases.append(random_tree(3000, 6, "caterpillar").to_newick())  # deep: the parser must not recurse
    for nw in cases:
        a = parse_newick(nw)
        b = py_parse_newick(nw.strip())
        assert a.names == b.names, nw[:60]
        assert np.array_equal(a.parent, b.parent) and np.array_equal(a.child_off, b.child_off)
        assert np.array_equal(a.child_idx, b.child_idx) and np.array_equal(a.leaf_row, b.leaf_row)
        assert a.polytomy == b.has_polytomy()
    t = parse_newick(cases[0])
    assert t.names[:4] == ["node_1", "node_2", "A", "B,x (1)"] and t.polytomy
    for bad in ["((A,B);", "A,B);", ""]:
        with pytest.raises(ValueError):
            parse_newick(bad)


def _fasta(names, rows, width=70, eol=b"\n", final_eol=True):
    out = []
    for n, r in zip(names, rows):
        out.append(b">" + n.encode() + b" some description")
        s = bytes(r)
        for i in range(0, len(s), width):
            out.append(s[i:i + width])
    return eol.join(out) + (eol if final_eol else b"")


def _msa_case(rng, trial, low_mem):
    """A random alignment + tree, and what the reference's -M branch makes of it: (tree, names, rows, reference id, consensus,
    the rows the passes see)."""
    alphabet = np.frombuffer(b"ACGTN-ACGTACGTRYKM", np.uint8)
    kind = ["binary", "polytomy", "caterpillar"][trial % 3]
    tree = random_tree(int(rng.integers(2, 80)), 700 + trial, kind, max_arity=4)
    names = [tree.names[v] for v in tree.leaves]
    n_cols = int(rng.integers(1, 2500))
    base = rng.choice(alphabet[:4], size=n_cols)
    rows = np.repeat(base[None, :], len(names), 0)
    noise = rng.random(rows.shape) < 0.1
    rows = np.where(noise, rng.choice(alphabet, size=rows.shape), rows).astype(np.uint8)
    if not low_mem:
        rows[:, rng.random(n_cols) < 0.03] = ord("-")  # all-gap columns: dropped by the -M branch
    else:
        rows = rows[:, (rows != ord("-")).any(0)]
    reference = names[int(rng.integers(0, len(names)))] if trial % 2 else ""
    order = sorted(range(len(names)), key=lambda i: names[i].encode())
    srt = rows[order]
    nongap = srt != ord("-")
    first = srt[nongap.argmax(0), np.arange(rows.shape[1])]
    if not low_mem:
        if reference:
            cons, use = rows[names.index(reference)].copy(), rows
        else:
            keep = nongap.any(0)
            cons, use = first[keep], rows[:, keep]
    else:
        use = rows
        if reference:
            r = rows[names.index(reference)]
            cons = np.where(r != ord("-"), r, np.where(nongap.any(0), first, 0)).astype(np.uint8)
        else:
            cons = first
    return tree, names, rows, reference, cons, use


@pytest.mark.parametrize("parallel_bytes", [1 << 40, 0])
@pytest.mark.parametrize("low_mem", [False, True])
def test_msa_prepare_on_the_host(low_mem, parallel_bytes):
    """The host-only half of the -M flow (pmh_msa_prepare: FASTA reader incl. wrapped lines and descriptions, consensus rule,
    all-gap column removal, nibble packing, per-column parameters) against the restated rules -- no device involved."""
    import panman_b200 as pb

    pb.build_library()
    from panman_b200.host import MsaPrepared, load_host_library

    # the reader cuts large files at header lines and parses the pieces in parallel: both paths must give the same
    import ctypes as C

    load_host_library().pmh_set_reader_parallel_bytes(C.c_int64(parallel_bytes))
    rng = np.random.default_rng(31 + int(low_mem))
    for trial in range(10):
        tree, names, rows, reference, cons, use = _msa_case(rng, trial, low_mem)
        # the MSA branch cuts lines at '\r' (CRLF files), the low-memory branch does not; wrapped or single-line records
        text = _fasta(names, rows, width=[70, 7, 10 ** 9][trial % 3], eol=b"\r\n" if (not low_mem and trial % 4 == 1) else b"\n",
                      final_eol=trial % 5 != 2)
        prep = MsaPrepared(text, tree.to_newick(), reference, low_mem)
        assert prep.tree.names == tree.names
        assert prep.consensus == bytes(cons), trial
        assert prep.n_cols == len(cons)
        row_of = {tree.names[v]: int(tree.leaf_row[v]) for v in tree.leaves}
        want = np.zeros((tree.n_leaves, len(cons)), np.uint8)
        for n, r in zip(names, use):
            want[row_of[n]] = CODE_OF[r]
        assert np.array_equal(prep.codes, want), trial
        assert prep.present.all()
        assert np.array_equal(prep.parent_code, CODE_OF[np.frombuffer(bytes(cons), np.uint8)])
        refcodes = CODE_OF[use[names.index(reference)]].astype(np.int8) if reference else None
        if low_mem:
            assert prep.fwd_root_ref is None and (np.array_equal(prep.root_override, refcodes) if reference else prep.root_override is None)
        else:
            assert prep.root_override is None and (np.array_equal(prep.fwd_root_ref, refcodes) if reference else prep.fwd_root_ref is None)
    load_host_library().pmh_set_reader_parallel_bytes(C.c_int64(8 << 20))


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["ctx", "group"])
@pytest.mark.parametrize("low_mem", [False, True])
def test_msa_build_matches_reference_flow(port, ref, low_mem, engine):
    """End to end -M flow (reader, consensus, packing, pmb_run_nuc, run-merge) against the restated reference drivers
    around the verbatim fitchSankoff.cpp and the oracle's run-merge (reference src/panman.cpp:1274-1466, 1467-1649).
    engine "group": the same through pmh_msa_run_group, the column ranges over a three-rank pmb_group."""
    import panman_b200 as pb
    from panman_b200.host import MsaBuild

    ctx = pb.Context(0) if engine == "ctx" else pb.Group([0, 0, 0])
    rng = np.random.default_rng(21 + int(low_mem))
    alphabet = np.frombuffer(b"ACGTN-ACGTACGTRYKM", np.uint8)
    for trial in range(8):
        kind = ["binary", "polytomy", "caterpillar"][trial % 3]
        tree = random_tree(int(rng.integers(2, 80)), 700 + trial, kind, max_arity=4)
        names = [tree.names[v] for v in tree.leaves]
        n_cols = int(rng.integers(1, 2500))
        base = rng.choice(alphabet[:4], size=n_cols)
        rows = np.repeat(base[None, :], len(names), 0)
        noise = rng.random(rows.shape) < 0.1
        rows = np.where(noise, rng.choice(alphabet, size=rows.shape), rows).astype(np.uint8)
        if not low_mem:
            rows[:, rng.random(n_cols) < 0.03] = ord("-")  # all-gap columns: dropped by the -M branch
        else:
            rows = rows[:, (rows != ord("-")).any(0)]
            n_cols = rows.shape[1]
        reference = names[int(rng.integers(0, len(names)))] if trial % 2 else ""
        order = sorted(range(len(names)), key=lambda i: names[i].encode())
        srt = rows[order]
        nongap = srt != ord("-")
        first = srt[nongap.argmax(0), np.arange(rows.shape[1])]
        if not low_mem:
            if reference:
                cons, use = rows[names.index(reference)].copy(), rows
            else:
                keep = nongap.any(0)
                cons, use = first[keep], rows[:, keep]
        else:
            use = rows
            if reference:
                r = rows[names.index(reference)]
                cons = np.where(r != ord("-"), r, np.where(nongap.any(0), first, 0)).astype(np.uint8)
            else:
                cons = first
        build = MsaBuild(ctx, _fasta(names, rows), tree.to_newick(), reference, low_mem)
        assert build.tree.names == tree.names
        assert build.consensus == bytes(cons), trial
        h = ref.tree(tree)
        want, _ = ref.msa_run(h, tree, int(low_mem), names, [bytes(r) for r in use], bytes(cons), reference, n_threads=2)
        ref.free(h)
        assert np.array_equal(build.tuple_offsets, want.node_offsets), trial
        assert np.array_equal(build.tuple_pos, want.pos) and np.array_equal(build.tuple_type_code, want.type_code), trial
        for v in range(tree.n_nodes):
            p, tc = want.of(v)
            wp, wmi, wnu = port.merge_msa(p, tc)
            got = build.nucmut[v]
            assert [g[0] for g in got] == list(wp) and [g[4] for g in got] == list(wmi) and [g[5] for g in got] == list(wnu), (trial, v)
            assert all(g[1] == -1 and g[2] == 0 and g[3] == -1 for g in got)
            # what the writer would store for the node (src/panman.cpp:2854-2929); the root carries the block insertion (:1439-1440)
            assert build.wire[v] == writer_layout(got, [(0, -1, 1, 0)] if v == tree.root else []), (trial, v)
    ctx.close()
