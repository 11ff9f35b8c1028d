"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the CPU oracles.

Two checkers live here, neither is ever on the product path (panman_b200/ must not import this):

* ``PortOracle``  -- oracle/fs_oracle.c, our array-based C restatement of the reference's
  src/fitchSankoff.cpp + the caller conventions of src/panman.cpp (kind "port").
* ``RefOracle``   -- oracle/_ref/libpanman_ref.so: the reference's src/fitchSankoff.cpp compiled
  verbatim + oracle/ref_driver.cpp (kind "reference"). Built only where /root/reference exists;
  the built .so travels to the GPU box.

Also a pure-Python restatement of the reference's Newick conventions (``parse_newick``,
reference src/panman.cpp:310-450) used to cross-check the C++ host parser.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle_port.so")
REF_SO = os.path.join(HERE, "_ref", "libpanman_ref.so")

NO_DEFAULT = 1 << 28  # reference src/panman.hpp:851


def build(verbose: bool = False) -> None:
    """Compile the port (always) and the verbatim reference build (when /root/reference exists)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", HERE, "all"], stdout=out)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


PGORDER_SO = os.path.join(HERE, "_ref", "libpanman_pgorder.so")


def have_ref_pgorder() -> bool:
    return os.path.exists(PGORDER_SO)


class RefPgOrder:
    """oracle/_ref/libpanman_pgorder.so: the reference's src/chaining.cpp + src/rotation.cpp compiled verbatim, driven by a
    restatement of the block-ordering half of Pangraph::Pangraph (oracle/pangraph_ref_driver.cpp). order(pg) takes the parsed
    PanGraph JSON and returns the block columns (consensus order of chain_align; duplicated blocks repeat) and, per path, which
    columns it owns, on which strand, and as which occurrence ("number")."""
    kind = "reference"

    def __init__(self):
        if not have_ref_pgorder():
            raise FileNotFoundError(PGORDER_SO)
        L = C.CDLL(PGORDER_SO)
        L.refpg_order.restype = C.c_void_p
        L.refpg_n_topo.restype = C.c_int64
        L.refpg_n_topo.argtypes = [C.c_void_p]
        L.refpg_topo_id.restype = C.c_char_p
        L.refpg_topo_id.argtypes = [C.c_void_p, C.c_int64]
        L.refpg_visit.restype = C.c_char_p
        L.refpg_visit.argtypes = [C.c_void_p, C.c_int]
        L.refpg_aligned_walk.restype = C.c_char_p
        L.refpg_aligned_walk.argtypes = [C.c_void_p, C.c_int]
        L.refpg_rotation_index.argtypes = [C.c_void_p, C.c_char_p]
        L.refpg_aligned.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.refpg_free.argtypes = [C.c_void_p]
        self.L = L

    def order(self, pg: dict):
        paths = [p for p in pg["paths"]]
        names = [p["name"].encode() for p in paths]
        off = np.zeros(len(paths) + 1, np.int64)
        ids, strands = [], []
        for k, p in enumerate(paths):
            for b in p["blocks"]:
                ids.append(b["id"].encode())
                strands.append(1 if b.get("strand", True) else 0)
            off[k + 1] = len(ids)
        circ = np.asarray([1 if p.get("circular") else 0 for p in paths], np.int32)
        strands = np.asarray(strands + [0], np.int32)
        blk_ids = [b["id"].encode() for b in pg["blocks"]]
        blk_len = np.asarray([len(b["sequence"]) for b in pg["blocks"]] + [0], np.int32)
        arr = lambda xs: (C.c_char_p * max(1, len(xs)))(*xs)
        h = C.c_void_p(self.L.refpg_order(C.c_int(len(paths)), arr(names), _p(off, C.c_int64), arr(ids), _p(strands, C.c_int32),
                                          _p(circ, C.c_int32), C.c_int(len(blk_ids)), arr(blk_ids), _p(blk_len, C.c_int32)))
        try:
            n = int(self.L.refpg_n_topo(h))
            out = dict(topo_ids=[self.L.refpg_topo_id(h, i).decode() for i in range(n)], aligned={}, strand={}, number={},
                       rotation_index={}, visit=[])
            with_blocks = [p["name"] for p in paths if p["blocks"]]
            out["visit"] = [self.L.refpg_visit(h, k).decode() for k in range(len(set(with_blocks)))]
            out["aligned_walk"] = [self.L.refpg_aligned_walk(h, k).decode() for k in range(len(set(with_blocks)))]
            for name in with_blocks:
                a, s_, nu = (np.empty(max(n, 1), np.int32) for _ in range(3))
                self.L.refpg_aligned(h, name.encode(), a.ctypes.data, s_.ctypes.data, nu.ctypes.data)
                out["aligned"][name], out["strand"][name], out["number"][name] = a[:n].copy(), s_[:n].copy(), nu[:n].copy()
                out["rotation_index"][name] = int(self.L.refpg_rotation_index(h, name.encode()))
            return out
        finally:
            self.L.refpg_free(h)


def wire_mut_info(mut_info, nucs):
    """NucMut.mutInfo as the capnp writer stores it (reference src/panman.cpp:2876): the `length` nucleotide codes,
    right-aligned, above the 8-bit mutInfo. Works on ints and numpy arrays."""
    length = np.asarray(mut_info, np.uint32) >> 4
    return ((np.asarray(nucs, np.uint32) >> (24 - 4 * length)) << 8) + np.asarray(mut_info, np.uint32)


NUCMUT_SO = os.path.join(HERE, "_ref", "libpanman_nucmut.so")


def have_ref_nucmut() -> bool:
    return os.path.exists(NUCMUT_SO)


class RefNucMut:
    """oracle/_ref/libpanman_nucmut.so: the reference's own struct NucMut (src/panman.hpp:75-313, extracted verbatim at
    build time) driven by restatements of the per-node merge loops (oracle/nucmut_driver.cpp). Each call takes ONE node's
    tuples in any order and returns the fields of the NucMut objects the reference appends to Node::nucMutation."""
    kind = "reference"

    def __init__(self):
        if not have_ref_nucmut():
            raise FileNotFoundError(NUCMUT_SO)
        L = C.CDLL(NUCMUT_SO)
        L.refnm_merge_msa.restype = C.c_int64
        L.refnm_merge_pangraph.restype = C.c_int64
        L.refnm_wire.restype = C.c_uint32
        self.L = L
        assert L.refnm_sizeof() == 24  # the 24-byte struct of src/panman.hpp:75-81

    @staticmethod
    def _outs(n):
        return (np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.uint8),
                np.empty(n, np.uint32))

    def merge_msa(self, pos, typ, code):
        """-> (primaryBlockId, secondaryBlockId, nucPosition, nucGapPosition, mutInfo, nucs)"""
        n = len(pos)
        pos, typ, code = (np.ascontiguousarray(pos, np.int32), np.ascontiguousarray(typ, np.int8), np.ascontiguousarray(code, np.int8))
        o = self._outs(max(n, 1))
        k = self.L.refnm_merge_msa(C.c_int64(n), _p(pos, C.c_int32), _p(typ, C.c_int8), _p(code, C.c_int8), _p(o[0], C.c_int32),
                                   _p(o[1], C.c_int32), _p(o[2], C.c_int32), _p(o[3], C.c_int32), _p(o[4], C.c_uint8), _p(o[5], C.c_uint32))
        return tuple(x[:k].copy() for x in o)

    def merge_pangraph(self, gap, block, pos, gap_pos, typ, code):
        n = len(pos)
        a = lambda x: np.ascontiguousarray(x, np.int32)
        block, pos, gap_pos, typ, code = a(block), a(pos), a(gap_pos), a(typ), a(code)
        o = self._outs(max(n, 1))
        k = self.L.refnm_merge_pangraph(C.c_int(gap), C.c_int64(n), _p(block, C.c_int32), _p(pos, C.c_int32), _p(gap_pos, C.c_int32),
                                        _p(typ, C.c_int32), _p(code, C.c_int32), _p(o[0], C.c_int32), _p(o[1], C.c_int32),
                                        _p(o[2], C.c_int32), _p(o[3], C.c_int32), _p(o[4], C.c_uint8), _p(o[5], C.c_uint32))
        return tuple(x[:k].copy() for x in o)

    def wire(self, mut_info: int, nucs: int, nuc_position: int = 0):
        """-> (mutInfo on the wire, src/panman.cpp:2876; and the (nucPosition, mutInfo, nucs, length, type, codes[6]) the
        reader constructor src/panman.hpp:191-211 rebuilds from it)"""
        bp, bl, bt = C.c_int32(), C.c_int32(), C.c_int32()
        bi, bn = C.c_uint8(), C.c_uint32()
        codes = (C.c_int32 * 6)()
        w = self.L.refnm_wire(C.c_uint8(mut_info), C.c_uint32(nucs), C.c_int32(nuc_position), C.byref(bp), C.byref(bi), C.byref(bn),
                              C.byref(bl), C.byref(bt), codes)
        return int(w), (bp.value, bi.value, bn.value, bl.value, bt.value, list(codes))


# --------------------------------------------------------------------------- trees


@dataclass
class FlatTree:
    """Node ids in creation order of the reference parser: an internal node is created when its
    '(' is met, a leaf when its name is met (reference src/panman.cpp:404-437)."""

    names: list
    parent: np.ndarray  # int32, -1 for root
    child_off: np.ndarray  # int32, n+1
    child_idx: np.ndarray  # int32
    root: int = 0
    leaf_row: np.ndarray = field(default=None)  # int32, -1 internal; rows number leaves in id order

    @property
    def n_nodes(self) -> int:
        return len(self.names)

    @property
    def leaves(self) -> np.ndarray:
        return np.nonzero(self.leaf_row >= 0)[0].astype(np.int32)

    @property
    def n_leaves(self) -> int:
        return int((self.leaf_row >= 0).sum())

    @staticmethod
    def from_children(names, children, root=0) -> "FlatTree":
        n = len(names)
        parent = np.full(n, -1, np.int32)
        off = np.zeros(n + 1, np.int32)
        idx = []
        for v in range(n):
            for c in children[v]:
                parent[c] = v
                idx.append(c)
            off[v + 1] = len(idx)
        leaf_row = np.full(n, -1, np.int32)
        r = 0
        for v in range(n):
            if off[v + 1] == off[v]:
                leaf_row[v] = r
                r += 1
        return FlatTree(list(names), parent, off, np.asarray(idx, np.int32), root, leaf_row)

    def has_polytomy(self) -> bool:  # reference src/panman.cpp:621-631
        return bool((np.diff(self.child_off) > 2).any())

    def to_newick(self) -> str:
        out = []
        # iterative to survive caterpillars
        stack = [(self.root, 0)]
        while stack:
            v, k = stack.pop()
            b, e = self.child_off[v], self.child_off[v + 1]
            if b == e:
                out.append(self.names[v])
                continue
            if k == 0:
                out.append("(")
            elif k < e - b:
                out.append(",")
            if k < e - b:
                stack.append((v, k + 1))
                stack.append((int(self.child_idx[b + k]), 0))
            else:
                out.append(")")
        return "".join(out) + ";"


def _split_quote_aware(s: str, delim: str):
    """reference src/panman.cpp:265-295 (stringSplit): a piece with an odd number of apostrophes is glued
    to the following pieces until the count is even."""
    words, start, temp, have_temp = [], 0, 0, False
    while True:
        end = s.find(delim, start)
        if end < 0:
            break
        if not have_temp:
            sub = s[start:end]
            if sub.count("'") % 2 == 1:
                temp, have_temp = start, True
            else:
                words.append(sub)
        else:
            sub = s[temp:end]
            if sub.count("'") % 2 == 0:
                have_temp = False
                words.append(sub)
        start = end + 1
    last = s[start:]
    if last != "":
        words.append(last)
    return words


def parse_newick(newick: str) -> FlatTree:
    """Restates the tree-shape part of createTreeFromNewickString (reference src/panman.cpp:310-450):
    comma-split pieces; per piece count '(' and ')' and collect the leaf name (everything before the first
    ':' or ')' that is not a parenthesis; quoted names keep their inside verbatim, surrounding apostrophes
    stripped); internal nodes are named node_1, node_2, ... in order of their '(' (panman.hpp:793-795);
    children are appended in Newick order (panman.cpp:223-229). Branch lengths are irrelevant to this path."""
    s = newick
    while s and s[-1] == " ":  # stripString, panman.cpp:298-308 (pops two per trailing blank; harmless here)
        s = s[:-2]
    s = s.lstrip(" ")
    names, children, stack = [], [], []
    counter = 0
    for piece in _split_quote_aware(s, ","):
        n_open = n_close = 0
        stop = name_zone = has_apo = False
        leaf = ""
        for ch in piece:
            if name_zone:
                leaf += ch
                if ch == "'":
                    name_zone = False
            elif ch == "'":
                name_zone = has_apo = True
                leaf += ch
            elif ch == ":":
                stop = True
            elif ch == "(":
                n_open += 1
            elif ch == ")":
                stop = True
                n_close += 1
            elif not stop:
                leaf += ch
        if has_apo and leaf[0] == "'" and leaf[-1] == "'":
            leaf = leaf[1:-1]
        for _ in range(n_open):
            counter += 1
            names.append(f"node_{counter}")
            children.append([])
            if stack:
                children[stack[-1]].append(len(names) - 1)
            stack.append(len(names) - 1)
        names.append(leaf)
        children.append([])
        children[stack[-1]].append(len(names) - 1)
        for _ in range(n_close):
            stack.pop()
    if stack:
        raise ValueError("incorrect Newick format")
    return FlatTree.from_children(names, children, 0)


def random_tree(n_leaves: int, seed: int, kind: str = "binary", max_arity: int = 2, prefix: str = "L") -> FlatTree:
    """Test trees. kind: 'binary' = random join of forest roots; 'caterpillar' = one spine;
    'polytomy' = random join of 2..max_arity roots; 'unary' sprinkles single-child nodes."""
    rng = np.random.default_rng(seed)
    kids = {}  # temp id -> children (temp ids)
    roots = list(range(n_leaves))
    nxt = n_leaves
    if kind == "caterpillar":
        order = list(rng.permutation(n_leaves))
        cur = order[0]
        for leaf in order[1:]:
            kids[nxt] = [cur, leaf] if rng.random() < 0.5 else [leaf, cur]
            cur = nxt
            nxt += 1
        roots = [cur]
    else:
        while len(roots) > 1:
            k = 2
            if kind in ("polytomy", "unary"):
                k = int(rng.integers(2, max(2, max_arity) + 1))
            k = min(k, len(roots))
            pick = rng.choice(len(roots), size=k, replace=False)
            ch = [roots[i] for i in pick]
            for i in sorted(pick, reverse=True):
                roots.pop(i)
            kids[nxt] = ch
            roots.append(nxt)
            nxt += 1
            if kind == "unary" and rng.random() < 0.2:
                kids[nxt] = [roots.pop()]
                roots.append(nxt)
                nxt += 1
    if n_leaves == 1:
        kids[nxt] = [0]
        roots = [nxt]
    # renumber in the reference's creation order (pre-order, children left to right)
    names, children, stack = [], [], [(roots[0], -1)]
    counter = 0
    while stack:
        t, par = stack.pop()
        me = len(names)
        if t in kids:
            counter += 1
            names.append(f"node_{counter}")
        else:
            names.append(f"{prefix}{t}")
        children.append([])
        if par >= 0:
            children[par].append(me)
        for c in reversed(kids.get(t, [])):
            stack.append((c, me))
    return FlatTree.from_children(names, children, 0)


# --------------------------------------------------------------------------- results


@dataclass
class MutLists:
    node_offsets: np.ndarray  # int64, n_nodes+1
    pos: np.ndarray  # int32
    type_code: np.ndarray  # uint8, (type<<4)|code

    def of(self, v: int):
        a, b = self.node_offsets[v], self.node_offsets[v + 1]
        return self.pos[a:b], self.type_code[a:b]

    def same_as(self, other: "MutLists") -> bool:
        return (
            np.array_equal(self.node_offsets, other.node_offsets)
            and np.array_equal(self.pos, other.pos)
            and np.array_equal(self.type_code, other.type_code)
        )


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


# --------------------------------------------------------------------------- port


class PortOracle:
    kind = "port"

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build()
        L = C.CDLL(PORT_SO)
        L.orc_run.restype = C.c_int
        L.orc_result_n_mut.restype = C.c_int64
        L.orc_result_offsets.restype = C.POINTER(C.c_int64)
        L.orc_result_pos.restype = C.POINTER(C.c_int32)
        L.orc_result_type_code.restype = C.POINTER(C.c_uint8)
        L.orc_merge_msa.restype = C.c_int64
        L.orc_merge_pangraph.restype = C.c_int64
        for f in (L.orc_result_n_mut, L.orc_result_offsets, L.orc_result_pos, L.orc_result_type_code, L.orc_result_free):
            f.argtypes = [C.c_void_p]
        self.L = L

    def run(self, tree: FlatTree, algo: int, leaf_codes: np.ndarray, parent_code: np.ndarray, root_override=None,
            fwd_root_ref=None, leaf_present=None, block_mode: int = 0, n_threads: int = 1, want_states: bool = False):
        """Same inputs as pmb_run_nuc; leaf_codes is (n_leaves, n_cols) uint8, one code per byte."""
        leaf_codes = np.ascontiguousarray(leaf_codes, np.uint8)
        n_rows, n_cols = leaf_codes.shape
        assert n_rows == tree.n_leaves
        parent_code = np.ascontiguousarray(parent_code, np.uint8)
        ro = None if root_override is None else np.ascontiguousarray(root_override, np.int8)
        fr = None if fwd_root_ref is None else np.ascontiguousarray(fwd_root_ref, np.int8)
        lp = None if leaf_present is None else np.ascontiguousarray(leaf_present, np.uint8)
        states = np.empty((tree.n_nodes, n_cols), np.uint8) if want_states else None
        res = C.c_void_p()
        rc = self.L.orc_run(
            C.c_int(algo), C.c_int(block_mode), C.c_int(tree.n_nodes), C.c_int(tree.root), _p(tree.child_off, C.c_int32),
            _p(tree.child_idx, C.c_int32), _p(tree.leaf_row, C.c_int32), C.c_int64(n_cols), _p(leaf_codes, C.c_uint8),
            _p(lp, C.c_uint8), _p(parent_code, C.c_uint8), _p(ro, C.c_int8), _p(fr, C.c_int8), C.c_int(n_threads),
            _p(states, C.c_uint8), C.byref(res))
        n = self.L.orc_result_n_mut(res)
        off = np.ctypeslib.as_array(self.L.orc_result_offsets(res), (tree.n_nodes + 1,)).copy()
        pos = np.ctypeslib.as_array(self.L.orc_result_pos(res), (max(n, 1),))[:n].copy()
        tc = np.ctypeslib.as_array(self.L.orc_result_type_code(res), (max(n, 1),))[:n].copy()
        self.L.orc_result_free(res)
        if rc != 0:
            raise RuntimeError(f"oracle port: rc={rc} (Sankoff root without finite cost: reference asserts)")
        return MutLists(off, pos, tc), states

    def column(self, tree: FlatTree, mode: int, leaf_val: np.ndarray, fwd_ref: int, parent_state: int,
               default_state: int = NO_DEFAULT):
        """mode 0 nuc Fitch, 1 nuc Sankoff, 2 block Fitch, 3 block Sankoff; leaf_val indexed by node id."""
        n = tree.n_nodes
        lv = np.ascontiguousarray(leaf_val, np.int32)
        fin = np.empty(n, np.int32)
        mt = np.empty(n, np.int32)
        mc = np.empty(n, np.int32)
        ci = C.POINTER(C.c_int)
        if mode in (0, 2):
            fwd = np.empty(n, np.int32)
            rc = self.L.orc_fitch_column(n, tree.root, _p(tree.child_off, C.c_int32), _p(tree.child_idx, C.c_int32),
                                         _p(lv, C.c_int), fwd_ref, parent_state, default_state, int(mode == 2),
                                         _p(fwd, C.c_int), _p(fin, C.c_int), _p(mt, C.c_int), _p(mc, C.c_int))
        else:
            W = 16 if mode == 1 else 3
            fwd = np.empty((n, W), np.int32)
            rc = self.L.orc_sankoff_column(n, tree.root, _p(tree.child_off, C.c_int32), _p(tree.child_idx, C.c_int32),
                                           _p(lv, C.c_int), W, int(mode == 3), parent_state, default_state,
                                           _p(fwd, C.c_int), _p(fin, C.c_int), _p(mt, C.c_int), _p(mc, C.c_int))
        return rc, fwd, fin, mt, mc

    def merge_msa(self, pos: np.ndarray, type_code: np.ndarray):
        n = len(pos)
        pos = np.ascontiguousarray(pos, np.int32)
        tc = np.ascontiguousarray(type_code, np.uint8)
        op = np.empty(max(n, 1), np.int32)
        mi = np.empty(max(n, 1), np.uint8)
        nu = np.empty(max(n, 1), np.uint32)
        k = self.L.orc_merge_msa(C.c_int64(n), _p(pos, C.c_int32), _p(tc, C.c_uint8), _p(op, C.c_int32), _p(mi, C.c_uint8),
                                 _p(nu, C.c_uint32))
        return op[:k].copy(), mi[:k].copy(), nu[:k].copy()


    def merge_pangraph(self, gap: int, block: np.ndarray, pos: np.ndarray, gap_pos: np.ndarray, type_code: np.ndarray):
        """reference src/panman.cpp:1236-1272 on one node's sorted 6-tuples; gap = 0 the non-gap list, 1 the gap list."""
        n = len(pos)
        a = lambda x, t: np.ascontiguousarray(x, t)
        block, pos, gap_pos, tc = a(block, np.int32), a(pos, np.int32), a(gap_pos, np.int32), a(type_code, np.uint8)
        ob, op, og = (np.empty(max(n, 1), np.int32) for _ in range(3))
        mi = np.empty(max(n, 1), np.uint8)
        nu = np.empty(max(n, 1), np.uint32)
        k = self.L.orc_merge_pangraph(C.c_int(gap), C.c_int64(n), _p(block, C.c_int32), _p(pos, C.c_int32), _p(gap_pos, C.c_int32),
                                      _p(tc, C.c_uint8), _p(ob, C.c_int32), _p(op, C.c_int32), _p(og, C.c_int32), _p(mi, C.c_uint8),
                                      _p(nu, C.c_uint32))
        return ob[:k].copy(), op[:k].copy(), og[:k].copy(), mi[:k].copy(), nu[:k].copy()


# --------------------------------------------------------------------------- verbatim reference


class RefOracle:
    kind = "reference"

    def __init__(self):
        if not have_ref():
            raise FileNotFoundError(REF_SO)
        L = C.CDLL(REF_SO)
        L.ref_tree_new.restype = C.c_void_p
        L.ref_tree_free.argtypes = [C.c_void_p]
        L.ref_column.restype = C.c_int
        L.ref_msa_run.restype = C.c_int
        self.L = L

    def tree(self, t: FlatTree):
        arr = (C.c_char_p * t.n_nodes)(*[s.encode() for s in t.names])
        h = self.L.ref_tree_new(C.c_int(t.n_nodes), arr, _p(t.parent, C.c_int), _p(t.child_off, C.c_int),
                                _p(t.child_idx, C.c_int))
        return C.c_void_p(h)

    def free(self, h):
        self.L.ref_tree_free(h)

    def column(self, h, tree: FlatTree, mode: int, leaf_val: np.ndarray, fwd_ref: int, parent_state: int,
               default_state: int = NO_DEFAULT):
        """Returns rc, fwd, final, mut_type, mut_code with mut_code already mapped char -> code for nuc modes
        (the callers store getCodeFromNucleotide(char), reference panman.cpp:1428)."""
        n = tree.n_nodes
        lv = np.ascontiguousarray(leaf_val, np.int32)
        W = {0: 1, 2: 1, 1: 16, 3: 3}[mode]
        fwd = np.empty((n, W), np.int32)
        fin = np.empty(n, np.int32)
        mt = np.empty(n, np.int32)
        ma = np.empty(n, np.int32)
        rc = self.L.ref_column(h, C.c_int(mode), _p(lv, C.c_int), C.c_int(fwd_ref), C.c_int(parent_state),
                               C.c_int(default_state), _p(fwd, C.c_int), _p(fin, C.c_int), _p(mt, C.c_int), _p(ma, C.c_int))
        if mode in (0, 1):
            lut = np.zeros(256, np.int32)
            for ch, code in zip("ACMGRSVTWYHKDBN", range(1, 16)):
                lut[ord(ch)] = code
            ma = np.where(mt >= 0, lut[np.clip(ma, 0, 255)], 0).astype(np.int32)
        if W == 1:
            fwd = fwd[:, 0]
        return rc, fwd, fin, mt, ma

    def msa_run(self, h, tree: FlatTree, algo: int, seq_ids, seqs, consensus: bytes, reference_id: str = "",
                n_threads: int = 1, want_states: bool = False):
        n_cols = len(consensus)
        ids = (C.c_char_p * len(seq_ids))(*[s.encode() for s in seq_ids])
        sq = (C.c_char_p * len(seqs))(*[bytes(s) for s in seqs])
        states = np.empty((tree.n_nodes, n_cols), np.uint8) if want_states else None
        rc = self.L.ref_msa_run(h, C.c_int(algo), C.c_int(len(seq_ids)), ids, sq, C.c_int64(n_cols), C.c_char_p(consensus),
                                C.c_char_p(reference_id.encode()), C.c_int(n_threads), _p(states, C.c_uint8))
        if rc != 0:
            raise RuntimeError("reference driver: reference id not among the sequences")
        counts = np.empty(tree.n_nodes, np.int64)
        self.L.ref_result_counts(h, _p(counts, C.c_int64))
        n = int(counts.sum())
        pos = np.empty(max(n, 1), np.int32)
        ty = np.empty(max(n, 1), np.int8)
        co = np.empty(max(n, 1), np.int8)
        self.L.ref_result_copy(h, _p(pos, C.c_int32), _p(ty, C.c_int8), _p(co, C.c_int8))
        off = np.zeros(tree.n_nodes + 1, np.int64)
        off[1:] = np.cumsum(counts)
        tc = ((ty[:n].astype(np.uint8) << 4) | co[:n].astype(np.uint8)).astype(np.uint8)
        return MutLists(off, pos[:n].copy(), tc), states


CODE_OF = np.zeros(256, np.uint8)
for _ch, _code in zip("ACMGRSVTWYHKDBN", range(1, 16)):
    CODE_OF[ord(_ch)] = _code
CHAR_OF = np.frombuffer(b"-ACMGRSVTWYHKDBN", np.uint8)


def block_mut_from_nuc(mut_type: np.ndarray, mut_code: np.ndarray):
    """Map this repo's nuc-style record of a 3-state block column to the reference's
    (BlockMutationType, inversion) pair: parent absent => (BI, state==reverse); state absent => (BD, False);
    otherwise an inversion, which the reference writes as (BD, True).
    reference src/fitchSankoff.cpp:272-308 / :788-818; BI=1, BD=0 (src/panman.hpp:63-72)."""
    mt = np.asarray(mut_type)
    mc = np.asarray(mut_code)
    btype = np.where(mt == 2, 1, np.where(mt >= 0, 0, -1)).astype(np.int32)
    inv = np.where(mt == 2, (mc == 2), np.where(mt == 0, 1, 0)).astype(np.int32)
    inv = np.where(mt >= 0, inv, 0).astype(np.int32)
    return btype, inv


def ref_run_columns(ref: "RefOracle", tree: FlatTree, algo: int, leaf_codes: np.ndarray, parent_code: np.ndarray,
                    root_override=None, fwd_root_ref=None, leaf_present=None, block_mode: int = 0):
    """Drive the verbatim reference one column at a time with the pmb_run_nuc input convention
    (slow; for fixtures and small parity cases). Returns (MutLists, states[n_nodes, n_cols] uint8, 0xFF = none).
    Block columns are reported nuc-style (see block_mut_from_nuc) so that all engines share one record form."""
    n_rows, n_cols = leaf_codes.shape
    mode = algo + (2 if block_mode else 0)
    h = ref.tree(tree)
    states = np.full((tree.n_nodes, n_cols), 0xFF, np.uint8)
    per_node = [[] for _ in range(tree.n_nodes)]
    leaves = tree.leaves
    rows = tree.leaf_row[leaves]
    try:
        for c in range(n_cols):
            lv = np.full(tree.n_nodes, -1, np.int32)
            codes = leaf_codes[rows, c].astype(np.int32)
            present = np.ones(len(rows), bool) if leaf_present is None else leaf_present[rows].astype(bool)
            if algo == 0:
                lv[leaves] = np.where(present, 1 << codes, -1)
                ps = 1 << int(parent_code[c])
                ov = NO_DEFAULT if root_override is None or root_override[c] < 0 else 1 << int(root_override[c])
                fr = -1 if fwd_root_ref is None or fwd_root_ref[c] < 0 else 1 << int(fwd_root_ref[c])
            else:
                lv[leaves] = np.where(present, codes, -1)
                ps = int(parent_code[c])
                ov = NO_DEFAULT if root_override is None or root_override[c] < 0 else int(root_override[c])
                fr = -1
            rc, _, fin, mt, ma = ref.column(h, tree, mode, lv, fr, ps, ov)
            if rc != 0:
                raise RuntimeError(f"reference would assert at column {c}")
            if algo == 0:
                code = np.where(fin > 0, np.log2(np.maximum(fin, 1)).astype(np.int32), 0xFF)
            else:
                code = np.where(fin >= 0, fin, 0xFF)
            states[:, c] = code.astype(np.uint8)
            for v in np.nonzero(mt >= 0)[0]:
                if block_mode:
                    # reference pair (BI/BD, inversion) -> nuc-style (type, code); the code is the node's state
                    t = 2 if mt[v] == 1 else (0 if ma[v] else 1)
                    per_node[v].append((c, t, 0 if t == 1 else int(code[v])))
                else:
                    per_node[v].append((c, int(mt[v]), int(ma[v])))
    finally:
        ref.free(h)
    off = np.zeros(tree.n_nodes + 1, np.int64)
    pos, tc = [], []
    for v in range(tree.n_nodes):
        per_node[v].sort()
        off[v + 1] = off[v] + len(per_node[v])
        for p, t, k in per_node[v]:
            pos.append(p)
            tc.append((t << 4) | k)
    return MutLists(off, np.asarray(pos, np.int32), np.asarray(tc, np.uint8)), states
