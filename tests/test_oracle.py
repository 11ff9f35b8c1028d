"""CPU tests of the checkers themselves: the C port (oracle/fs_oracle.c) against the reference's own
fitchSankoff.cpp compiled verbatim (oracle/_ref), and against the committed golden vectors."""
import numpy as np
import pytest

from oracle.oracle import (CODE_OF, NO_DEFAULT, block_mut_from_nuc, parse_newick, random_tree, ref_run_columns)
from tests.golden_util import load_cases


def test_newick_conventions():
    # node_<k> numbering in order of '(' ; children in Newick order; quoted names; branch lengths ignored
    t = parse_newick("((A:0.1,'B,x (1)':2):0.5,(C,D,E)0.9:1,F);")
    assert t.names == ["node_1", "node_2", "A", "B,x (1)", "node_3", "C", "D", "E", "F"]
    assert list(t.parent) == [-1, 0, 1, 1, 0, 4, 4, 4, 0]
    assert t.has_polytomy()
    assert list(t.leaf_row) == [-1, -1, 0, 1, -1, 2, 3, 4, 5]
    t2 = parse_newick("(A,B);")
    assert not t2.has_polytomy() and t2.names == ["node_1", "A", "B"]
    for kind in ("binary", "polytomy", "unary", "caterpillar"):
        r = random_tree(50, 3, kind, max_arity=5)
        rt = parse_newick(r.to_newick())
        assert rt.names == r.names and np.array_equal(rt.child_idx, r.child_idx)


def test_port_matches_golden(port):
    for c in load_cases():
        got, states = port.run(c["tree"], c["algo"], c["codes"], c["parent_code"], c["root_override"], c["fwd_root_ref"],
                               c["leaf_present"], c["block"], n_threads=2, want_states=True)
        assert got.same_as(c["expect"]), f"golden case {c['id']}"
        assert np.array_equal(states, c["states"]), f"golden case {c['id']} states"


def test_port_vs_reference_columns(port, ref):
    rng = np.random.default_rng(0)
    n = 0
    for trial in range(300):
        kind = ["binary", "polytomy", "unary", "caterpillar"][trial % 4]
        t = random_tree(int(rng.integers(1, 40)), trial, kind, max_arity=5)
        h = ref.tree(t)
        for mode in range(4):
            nst = 16 if mode < 2 else 3
            codes = rng.integers(0, rng.integers(1, nst + 1), size=t.n_nodes)
            absent = rng.random(t.n_nodes) < (0.3 if trial % 3 == 0 else 0.0)
            if mode in (0, 2):
                lv = np.where(absent, -1, 1 << codes).astype(np.int32)
                ps = 1 << int(rng.integers(0, nst))
                dflt = (1 << int(rng.integers(0, nst))) if rng.random() < 0.4 else NO_DEFAULT
                fr = (1 << int(rng.integers(0, nst))) if (mode == 0 and rng.random() < 0.3) else -1
            else:
                lv = np.where(absent, -1, codes).astype(np.int32)
                ps = int(rng.integers(0, nst))
                dflt = int(rng.integers(0, nst)) if rng.random() < 0.4 else NO_DEFAULT
                fr = -1
            a = list(port.column(t, mode, lv, fr, ps, dflt))
            b = ref.column(h, t, mode, lv, fr, ps, dflt)
            assert a[0] == b[0]
            if a[0] != 0:
                continue  # the reference's assert(minPtr != -1): both report -2
            if mode >= 2:
                a[3], a[4] = block_mut_from_nuc(a[3], a[4])
            for x, y, name in zip(a[1:], b[1:], ["forward", "final", "mut_type", "mut_arg"]):
                assert np.array_equal(x, y), (trial, kind, mode, name)
            n += 1
        ref.free(h)
    assert n > 1000


def _msa_inputs(rng, tree, n_cols, with_reference):
    """Random MSA as the reference's -M reader would hold it: std::map id -> row (byte-wise id order)."""
    names = [tree.names[v] for v in tree.leaves]
    alphabet = np.frombuffer(b"ACGTN-ACGTACGTRYKM", np.uint8)
    base = rng.choice(alphabet[:4], size=n_cols)
    rows = np.repeat(base[None, :], len(names), 0)
    noise = rng.random(rows.shape) < 0.15
    rows = np.where(noise, rng.choice(alphabet, size=rows.shape), rows).astype(np.uint8)
    rows[:, rng.random(n_cols) < 0.05] = ord("-")  # all-gap columns (dropped when no --reference)
    reference = names[int(rng.integers(0, len(names)))] if with_reference else ""
    return names, rows, reference


def _consensus(names, rows, reference):
    """reference src/panman.cpp:1332-1362: the --reference row, else first non-gap in map order with all-gap
    columns removed from every sequence."""
    order = sorted(range(len(names)), key=lambda i: names[i].encode())
    if reference:
        return rows[names.index(reference)].copy(), rows
    srt = rows[order]
    nongap = srt != ord("-")
    keep = nongap.any(0)
    first = nongap.argmax(0)
    cons = srt[first, np.arange(rows.shape[1])]
    return cons[keep], rows[:, keep]


@pytest.mark.parametrize("algo", [0, 1])
def test_port_vs_reference_msa_batch(port, ref, algo):
    """The string-keyed restated MSA drivers around the verbatim functions (Fitch panman.cpp:1381-1435, Sankoff
    :1568-1613) against the array port fed through the pmb_run_nuc input convention."""
    rng = np.random.default_rng(7 + algo)
    for trial in range(12):
        kind = ["binary", "polytomy", "caterpillar"][trial % 3]
        tree = random_tree(int(rng.integers(2, 60)), 500 + trial, kind, max_arity=4)
        names, rows, reference = _msa_inputs(rng, tree, int(rng.integers(1, 120)), with_reference=trial % 2 == 1)
        if algo == 1 and not reference:
            # low-mem mode exits on an all-gap column without --reference (panman.cpp:1548-1551)
            rows = rows[:, (rows != ord("-")).any(0)]
        if algo == 0:
            cons, rows = _consensus(names, rows, reference)
        else:
            # :1527-1557: reference char, or first non-gap if the reference has a gap (stale '\0' if all gaps)
            cons0, _ = _consensus(names, rows, "")
            cons = cons0 if not reference else None
            if reference:
                r = rows[names.index(reference)]
                order = sorted(range(len(names)), key=lambda i: names[i].encode())
                srt = rows[order]
                nongap = srt != ord("-")
                first = srt[nongap.argmax(0), np.arange(rows.shape[1])]
                cons = np.where(r != ord("-"), r, np.where(nongap.any(0), first, 0)).astype(np.uint8)
        h = ref.tree(tree)
        want, want_states = ref.msa_run(h, tree, algo, names, [bytes(r) for r in rows], bytes(cons), reference,
                                        n_threads=1 + trial % 3, want_states=True)
        ref.free(h)
        codes = CODE_OF[rows]
        parent_code = CODE_OF[cons]
        ridx = names.index(reference) if reference else -1
        if algo == 0:
            fr = codes[ridx].astype(np.int8) if reference else None  # refState :1419
            got, states = port.run(tree, 0, codes, parent_code, None, fr, None, 0, n_threads=2, want_states=True)
        else:
            ro = codes[ridx].astype(np.int8) if reference else None  # defaultState :1583-1596
            got, states = port.run(tree, 1, codes, parent_code, ro, None, None, 0, n_threads=2, want_states=True)
        assert got.same_as(want), (algo, trial)
        assert np.array_equal(states, want_states), (algo, trial)


def test_run_merge_msa(port):
    # reference panman.cpp:1445-1466 + panman.hpp:109-151: runs split at length 6, at a position gap, at a type change
    pos = np.asarray([3, 4, 5, 6, 7, 8, 9, 10, 20, 21, 22, 30], np.int32)
    typ = np.asarray([0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 2, 0], np.uint8)
    code = np.asarray([1, 2, 4, 8, 15, 1, 2, 4, 0, 0, 8, 5], np.uint8)
    p, mi, nu = port.merge_msa(pos, (typ << 4) | code)
    assert list(p) == [3, 9, 20, 22, 30]
    assert list(mi) == [(6 << 4) + 0, (2 << 4) + 0, (2 << 4) + 1, (1 << 4) + 2, (1 << 4) + 0]
    assert list(nu) == [0x1248F1, 0x240000, 0x000000, 0x800000, 0x500000]
    p, mi, nu = port.merge_msa(pos[:0], typ[:0])
    assert len(p) == 0


def _random_node_list(rng, n_cols=400):
    """One node's records: stretches of consecutive positions (runs longer than 6 included), type changes inside them."""
    pos = []
    p = int(rng.integers(0, 5))
    while p < n_cols:
        length = int(rng.choice([1, 1, 2, 3, 5, 6, 7, 12, 13, 19]))
        pos.extend(range(p, min(n_cols, p + length)))
        p += length + int(rng.integers(1, 9))
    pos = np.asarray(pos, np.int32)
    typ = np.repeat(rng.integers(0, 3, size=len(pos)), 1).astype(np.uint8)
    same = rng.random(len(pos)) < 0.8
    for i in range(1, len(pos)):
        if same[i]:
            typ[i] = typ[i - 1]
    code = rng.integers(0, 16, size=len(pos)).astype(np.uint8)
    return pos, typ, code


def test_run_merge_msa_vs_reference_struct(port, refnm):
    """a13 pinned to executable reference code: the port's MSA run-merge against the reference's OWN NucMut constructor
    (src/panman.hpp:109-151, compiled from the extracted struct) driven by the loop of src/panman.cpp:1445-1466."""
    rng = np.random.default_rng(1313)
    for trial in range(300):
        pos, typ, code = _random_node_list(rng)
        if trial % 7 == 0:
            pos, typ, code = pos[:1], typ[:1], code[:1]
        shuffle = rng.permutation(len(pos))  # the reference sorts; the port is handed position-sorted lists
        pb_, sb, rp, gap, info, nucs = refnm.merge_msa(pos[shuffle], typ[shuffle], code[shuffle])
        p, mi, nu = port.merge_msa(pos, (typ << 4) | code)
        assert np.array_equal(p, rp) and np.array_equal(mi, info) and np.array_equal(nu, nucs), trial
        assert (pb_ == 0).all() and (sb == -1).all() and (gap == -1).all()  # the constants pmb_merge_runs documents


def test_run_merge_pangraph_vs_reference_struct(port, refnm):
    """The PanGraph merges (src/panman.cpp:1236-1253 non-gap, :1255-1272 gap; 6-tuple ctor src/panman.hpp:154-189)."""
    rng = np.random.default_rng(1414)
    for trial in range(200):
        blocks, poss, gaps, typs, codes = [], [], [], [], []
        for b in sorted(rng.choice(12, size=int(rng.integers(1, 5)), replace=False)):
            pos, typ, code = _random_node_list(rng, n_cols=120)
            if trial % 2:  # gap list: records (pos, gapPos) with runs along gapPos
                gp = pos % 17
                pos = pos // 17
            else:
                gp = np.full(len(pos), -1, np.int32)
            blocks.append(np.full(len(pos), b, np.int32)); poss.append(pos); gaps.append(gp); typs.append(typ); codes.append(code)
        block, pos, gp, typ, code = (np.concatenate(x) for x in (blocks, poss, gaps, typs, codes))
        order = np.lexsort((code, typ, gp, pos, block))  # the order std::sort gives the 6-tuples
        block, pos, gp, typ, code = block[order], pos[order], gp[order], typ[order], code[order]
        if trial % 2:  # (block, pos, gapPos) must be unique per node, as in a real build
            key = np.stack([block, pos, gp], 1)
            keep = np.ones(len(pos), bool)
            keep[1:] = (key[1:] != key[:-1]).any(1)
            block, pos, gp, typ, code = block[keep], pos[keep], gp[keep], typ[keep], code[keep]
        gap = trial % 2
        shuffle = rng.permutation(len(pos))
        rb, rsb, rp, rg, rinfo, rnucs = refnm.merge_pangraph(gap, block[shuffle], pos[shuffle], gp[shuffle], typ[shuffle], code[shuffle])
        ob, op, og, mi, nu = port.merge_pangraph(gap, block, pos, gp, (typ.astype(np.uint8) << 4) | code.astype(np.uint8))
        assert np.array_equal(ob, rb) and np.array_equal(op, rp) and np.array_equal(og, rg), trial
        assert np.array_equal(mi, rinfo) and np.array_equal(nu, rnucs) and (rsb == -1).all(), trial


def test_wire_form_vs_reference_struct(refnm):
    """mutInfo on the wire (writer src/panman.cpp:2876) and back through the reader constructor (src/panman.hpp:191-211):
    the formula pmb_merge_runs' wire output uses, and the round trip it must survive."""
    from oracle.oracle import wire_mut_info

    rng = np.random.default_rng(1515)
    for _ in range(500):
        length = int(rng.integers(1, 7))
        typ = int(rng.integers(0, 3))
        codes = [int(c) for c in rng.integers(0, 16, size=length)]
        nucs = sum(c << (4 * (5 - k)) for k, c in enumerate(codes))
        info = (length << 4) + typ
        w, (bp, bi, bn, bl, bt, bc) = refnm.wire(info, nucs, 77)
        assert w == wire_mut_info(info, nucs)
        assert (bp, bi, bn, bl, bt) == (77, info, nucs, length, typ) and bc[:length] == codes


def test_sankoff_compact_identity(port):
    """the 2-bit excess recurrence used by the CUDA kernels reproduces the literal
    min-plus Sankoff of the port (itself pinned to the verbatim reference above)."""
    rng = np.random.default_rng(5)
    for trial in range(200):
        t = random_tree(int(rng.integers(2, 30)), 900 + trial, ["binary", "polytomy", "unary"][trial % 3], max_arity=6)
        codes = rng.integers(0, rng.integers(1, 17), size=t.n_nodes)
        absent = rng.random(t.n_nodes) < (0.3 if trial % 2 else 0.0)
        lv = np.where(absent, -1, codes).astype(np.int32)
        ps = int(rng.integers(0, 16))
        dflt = int(rng.integers(0, 16)) if trial % 3 == 0 else NO_DEFAULT
        rc, cost, fin, _, _ = port.column(t, 1, lv, -1, ps, dflt)
        if rc != 0:
            continue
        # compact form, bottom-up in reverse creation order (children have larger ids than parents)
        NONE = None
        e = [NONE] * t.n_nodes
        for v in range(t.n_nodes - 1, -1, -1):
            kids = t.child_idx[t.child_off[v]:t.child_off[v + 1]]
            if len(kids) == 0:
                e[v] = NONE if lv[v] < 0 else [0 if i == lv[v] else 2 for i in range(16)]
                continue
            live = [e[c] for c in kids if e[c] is not NONE]
            if not live:
                continue
            r = [sum(1 for x in live if x[i] > 0) for i in range(16)]
            m = min(r)
            e[v] = [min(2, x - m) for x in r]
        F = [-1] * t.n_nodes
        for v in range(t.n_nodes):
            if v == t.root:
                F[v] = dflt if dflt != NO_DEFAULT else e[v].index(0)
                continue
            s = F[t.parent[v]]
            if s == -1 or e[v] is NONE:
                continue
            z = e[v].index(0)
            F[v] = s if e[v][s] == 0 else (min(s, z) if e[v][s] == 1 else z)
        assert F == list(fin), trial
