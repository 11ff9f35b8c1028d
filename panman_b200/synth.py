"""Synthetic workloads of BASELINE.json / SURVEY.md 8(d): seeded trees and top-down simulated alignments.

Plumbing only (torch tensors on CPU or CUDA); the product kernels are in csrc/. The generator is stateless per
cell -- every random decision is a 64-bit hash of (seed, stream, node pre-order index, global column) -- so any
column range [c0, c1) can be produced in isolation on any rank and equals the same range of the whole matrix.

Tree kinds: "binary" = seeded random join (pick two roots of the forest uniformly, join), "caterpillar" =
caterpillar-heavy (90 % of the leaves hang one by one off a single spine, the rest form random-join subtrees hung
at uniformly chosen spine positions). Node ids follow the reference parser's creation order (pre-order, children
left to right; src/panman.cpp:404-437) and leaf rows number the leaves in id order, so row order is also the
byte-wise order of the zero-padded leaf names used for the consensus rule (src/panman.cpp:1338-1351).
"""
from dataclasses import dataclass

import numpy as np
import torch

_M64 = (1 << 64) - 1


@dataclass
class Tree:
    parent: np.ndarray  # int32
    child_off: np.ndarray  # int32 n+1
    child_idx: np.ndarray  # int32
    leaf_row: np.ndarray  # int32, -1 internal
    root: int = 0

    @property
    def n_nodes(self):
        return len(self.parent)

    @property
    def n_leaves(self):
        return int((self.leaf_row >= 0).sum())

    @property
    def leaves(self):
        return np.nonzero(self.leaf_row >= 0)[0].astype(np.int32)

    def leaf_names(self):
        w = len(str(max(1, self.n_leaves)))
        return [f"L{r:0{w}d}" for r in range(self.n_leaves)]

    def names(self):
        out, k, ln = [], 0, self.leaf_names()
        for v in range(self.n_nodes):
            if self.leaf_row[v] >= 0:
                out.append(ln[self.leaf_row[v]])
            else:
                k += 1
                out.append(f"node_{k}")
        return out

    def depth(self):
        d = np.zeros(self.n_nodes, np.int32)
        for v in range(1, self.n_nodes):  # parents precede children in creation order
            d[v] = d[self.parent[v]] + 1
        return d


def _finalize(kids, root_tmp):
    """temp ids -> creation-order ids (pre-order, children left to right)."""
    new_children = []
    stack = [(root_tmp, -1)]
    while stack:
        t, par = stack.pop()
        me = len(new_children)
        new_children.append([])
        if par >= 0:
            new_children[par].append(me)
        ch = kids.get(t)
        if ch:
            for c in reversed(ch):
                stack.append((c, me))
    n = len(new_children)
    parent = np.full(n, -1, np.int32)
    off = np.zeros(n + 1, np.int32)
    idx = np.empty(n - 1, np.int32)
    k = 0
    for v in range(n):
        for c in new_children[v]:
            parent[c] = v
            idx[k] = c
            k += 1
        off[v + 1] = k
    leaf_row = np.full(n, -1, np.int32)
    is_leaf = np.diff(off) == 0
    leaf_row[is_leaf] = np.arange(int(is_leaf.sum()), dtype=np.int32)
    return Tree(parent, off, idx, leaf_row, 0)


def _random_join(items, rng, kids, nxt):
    roots = list(items)
    while len(roots) > 1:
        i = int(rng.integers(0, len(roots)))
        a = roots[i]
        roots[i] = roots[-1]
        roots.pop()
        j = int(rng.integers(0, len(roots)))
        b = roots[j]
        kids[nxt] = [a, b]
        roots[j] = nxt
        nxt += 1
    return roots[0], nxt


def make_tree(n_leaves: int, seed: int, kind: str = "binary") -> Tree:
    rng = np.random.default_rng(seed)
    kids = {}
    nxt = n_leaves
    if n_leaves == 1:
        kids[nxt] = [0]
        return _finalize(kids, nxt)
    if kind == "binary":
        root, nxt = _random_join(range(n_leaves), rng, kids, nxt)
        return _finalize(kids, root)
    if kind == "caterpillar":
        n_spine = max(2, int(round(n_leaves * 0.9)))
        rest = list(range(n_spine, n_leaves))
        # the remaining 10 % as random-join subtrees of ~16 leaves hung at uniformly chosen spine positions
        hang = {}
        while rest:
            take = rest[:16]
            rest = rest[16:]
            sub, nxt = _random_join(take, rng, kids, nxt)
            hang.setdefault(int(rng.integers(1, n_spine)), []).append(sub)
        cur = 0
        for i in range(1, n_spine):
            kids[nxt] = [cur, i]
            cur = nxt
            nxt += 1
            for sub in hang.get(i, []):
                kids[nxt] = [cur, sub]
                cur = nxt
                nxt += 1
        return _finalize(kids, cur)
    raise ValueError(kind)


# ---------------------------------------------------------------- stateless per-cell randomness


def _mix_py(x: int) -> int:  # splitmix64 finalizer
    x &= _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def _to_i64(x: int) -> int:
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(x: torch.Tensor, s: int) -> torch.Tensor:
    return (x >> s) & ((1 << (64 - s)) - 1)


def _mix(x: torch.Tensor) -> torch.Tensor:
    x = (x ^ _lsr(x, 30)) * _to_i64(0xBF58476D1CE4E5B9)
    x = (x ^ _lsr(x, 27)) * _to_i64(0x94D049BB133111EB)
    return x ^ _lsr(x, 31)


def _keys(seed: int, stream: int, ids: np.ndarray, device) -> torch.Tensor:
    base = _mix_py(seed * 0x9E3779B97F4A7C15 + stream)
    ks = np.array([_to_i64(_mix_py(base + int(i) * 0xD1342543DE82EF95)) for i in ids], dtype=np.int64)
    return torch.from_numpy(ks).to(device)


_COLK = _to_i64(0xA0761D6478BD642F)


def _u32(keys: torch.Tensor, cols: torch.Tensor) -> torch.Tensor:
    """hash(key_i, col_j) -> 32 uniform bits, int64 tensor of shape (len(keys), len(cols))."""
    return _lsr(_mix(keys[:, None] + cols[None, :] * _COLK), 32)


@dataclass
class MsaSpec:
    seed: int
    p_sub: float
    f_gap: float
    p_N: float


def simulate_msa(tree: Tree, c0: int, c1: int, spec: MsaSpec, device="cpu", cell_budget: int = 1 << 25):
    """Columns [c0, c1) of the simulated alignment as nibble-packed leaf codes.

    Returns (codes4 uint8 [n_leaves, ceil((c1-c0)/2)] (pmb_run_nuc layout), parent_code uint8 [c1-c0]) on `device`.
    Root state uniform over A,C,G,T; on every edge the child copies its parent except with probability p_sub it draws
    a base uniformly; a fraction f_gap of the columns carries one or two clade events (a uniformly chosen non-root
    internal node whose subtree becomes '-', or is the only subtree that is not '-'); leaves become N with probability
    p_N. parent_code = code of the first non-gap leaf in row order (the -M consensus rule), 0 if the column is all gaps.
    """
    dev = torch.device(device)
    n, C = tree.n_nodes, c1 - c0
    cols = torch.arange(c0, c1, dtype=torch.int64, device=dev)
    depth = tree.depth()
    order = np.argsort(depth, kind="stable")
    bounds = np.flatnonzero(np.diff(depth[order])) + 1
    levels = np.split(order, bounds)
    # pre-order interval of every node (ids are pre-order): subtree(v) = [v, v + size[v])
    size = np.ones(n, np.int64)
    for v in range(n - 1, 0, -1):
        size[tree.parent[v]] += size[v]
    internal = np.nonzero(tree.leaf_row < 0)[0]
    non_root_internal = internal[internal != tree.root]

    thr_sub = int(spec.p_sub * (1 << 32))
    thr_gap = int(spec.f_gap * (1 << 32))
    thr_N = int(spec.p_N * (1 << 32))
    BASES = torch.tensor([1, 2, 4, 8], dtype=torch.uint8, device=dev)

    # per-column clade events
    ck = _keys(spec.seed, 7, np.arange(4), dev)
    h = _u32(ck, cols)  # 4 x C
    is_gap_col = h[0] < thr_gap
    ev_lo = torch.zeros((2, C), dtype=torch.int64, device=dev)
    ev_hi = torch.zeros((2, C), dtype=torch.int64, device=dev)
    ev_only = torch.zeros((2, C), dtype=torch.bool, device=dev)
    ev_on = torch.zeros((2, C), dtype=torch.bool, device=dev)
    if len(non_root_internal) > 0:
        nri = torch.from_numpy(non_root_internal.astype(np.int64)).to(dev)
        sz = torch.from_numpy(size).to(dev)
        for e in range(2):
            pick = nri[(h[1 + e] >> 4) % len(non_root_internal)]
            ev_lo[e] = pick
            ev_hi[e] = pick + sz[pick]
            ev_only[e] = (h[1 + e] & 1).bool()
            ev_on[e] = is_gap_col if e == 0 else (is_gap_col & ((h[3] & 1) == 1))

    state = {}  # node -> uint8 tensor [C], only while some child still needs it
    state_rows = torch.empty((n, 0), dtype=torch.uint8)
    codes4 = torch.zeros((tree.n_leaves, (C + 1) // 2), dtype=torch.uint8, device=dev)
    all_states = torch.zeros((n, C), dtype=torch.uint8, device=dev) if n * C <= (1 << 31) else None
    rows_per_chunk = max(1, cell_budget // max(1, C))

    def write_leaves(vs, st):
        """st: uint8 [len(vs), C] simulated bases of leaves vs -> events, N, pack."""
        vt = torch.from_numpy(vs.astype(np.int64)).to(dev)
        for e in range(2):
            inside = (vt[:, None] >= ev_lo[e][None, :]) & (vt[:, None] < ev_hi[e][None, :])
            gap = ev_on[e][None, :] & torch.where(ev_only[e][None, :], ~inside, inside)
            st = torch.where(gap, torch.zeros_like(st), st)
        hn = _u32(_keys(spec.seed, 3, vs, dev), cols)
        st = torch.where(hn < thr_N, torch.full_like(st, 15), st)
        if C % 2:
            st = torch.cat([st, torch.zeros((st.shape[0], 1), dtype=torch.uint8, device=dev)], 1)
        rows = torch.from_numpy(tree.leaf_row[vs].astype(np.int64)).to(dev)
        codes4[rows] = (st[:, 0::2] & 15) | ((st[:, 1::2] & 15) << 4)

    is_leaf = tree.leaf_row >= 0
    use_matrix = all_states is not None
    for lvl in levels:
        for a in range(0, len(lvl), rows_per_chunk):
            vs = lvl[a:a + rows_per_chunk]
            hv = _u32(_keys(spec.seed, 1, vs, dev), cols)
            newbase = BASES[(hv >> 8) & 3]
            if depth[vs[0]] == 0:
                st = newbase
            else:
                ps = tree.parent[vs]
                if use_matrix:
                    par = all_states[torch.from_numpy(ps.astype(np.int64)).to(dev)]
                else:
                    par = torch.stack([state[int(p)] for p in ps])
                st = torch.where(hv < thr_sub, newbase, par)
            lm = is_leaf[vs]
            if lm.any():
                write_leaves(vs[lm], st[torch.from_numpy(lm).to(dev)])
            if use_matrix:
                all_states[torch.from_numpy(vs.astype(np.int64)).to(dev)] = st
            else:
                for i, v in enumerate(vs):
                    if not lm[i]:
                        state[int(v)] = st[i]
        if not use_matrix and depth[lvl[0]] > 0:
            for p in set(int(x) for x in tree.parent[lvl]):
                state.pop(p, None)
    del state_rows

    # consensus: first non-gap leaf in row order
    parent_code = torch.zeros(C, dtype=torch.uint8, device=dev)
    found = torch.zeros(C, dtype=torch.bool, device=dev)
    step = max(1, cell_budget // max(1, C))
    for r0 in range(0, tree.n_leaves, step):
        blk = codes4[r0:r0 + step]
        un = torch.stack([blk & 15, blk >> 4], 2).reshape(blk.shape[0], -1)[:, :C]
        nz = un != 0
        first = torch.argmax(nz.to(torch.uint8), 0)
        anyz = nz.any(0)
        val = un[first, torch.arange(C, device=dev)]
        take = anyz & ~found
        parent_code = torch.where(take, val, parent_code)
        found |= anyz
        if bool(found.all()):
            break
    return codes4, parent_code


def unpack_nibbles(codes4: torch.Tensor, n_cols: int) -> torch.Tensor:
    un = torch.stack([codes4 & 15, codes4 >> 4], 2).reshape(codes4.shape[0], -1)
    return un[:, :n_cols].contiguous()


# the named configurations (SURVEY.md 8d / BASELINE.md section 2)
CONFIGS = {
    "sars20k": dict(n_leaves=20000, n_cols=30000, kind="binary", seed=2, p_sub=3e-5, f_gap=0.01, p_N=1e-3, algos=("fitch",)),
    "indel10k": dict(n_leaves=10000, n_cols=15000, kind="binary", seed=3, p_sub=1e-3, f_gap=0.20, p_N=1e-3,
                     algos=("fitch", "sankoff")),
    "ecoli4k": dict(n_leaves=4000, n_cols=5000000, kind="binary", seed=4, p_sub=1e-4, f_gap=0.05, p_N=1e-4, algos=("fitch",)),
    "caterpillar100k": dict(n_leaves=100000, n_cols=30000, kind="caterpillar", seed=5, p_sub=3e-5, f_gap=0.01, p_N=0.0,
                            algos=("fitch",)),
}
