"""Loader for tests/golden/fitch_sankoff_golden.npz (made by tests/golden/make_golden.py from the verbatim
reference build)."""
import os

import numpy as np

from oracle.oracle import MutLists, parse_newick

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fitch_sankoff_golden.npz")


def load_cases():
    z = np.load(PATH)
    cases = []
    for k in range(int(z["n_cases"][0])):
        p = f"c{k:02d}_"
        tree = parse_newick(bytes(z[p + "newick"]).decode())
        algo, block = (int(x) for x in z[p + "algo_block"])
        cases.append(dict(
            id=k, tree=tree, algo=algo, block=block, codes=z[p + "codes"], parent_code=z[p + "parent_code"],
            root_override=z[p + "root_override"],
            fwd_root_ref=z[p + "fwd_root_ref"] if p + "fwd_root_ref" in z else None,
            leaf_present=z[p + "leaf_present"] if p + "leaf_present" in z else None,
            expect=MutLists(z[p + "node_offsets"], z[p + "pos"], z[p + "type_code"]), states=z[p + "states"]))
    return cases


SARS20_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sars20_pangraph.npz")


def load_sars20():
    """BASELINE.json configs[0] as column batches (tests/golden/make_sars20_golden.py): returns (tree, batches); a batch is
    a dict in the pmb_run_nuc convention with expect[algo] = (MutLists, states) from the verbatim reference. The first
    batch is the block-level pass (3-state, block mode), the others are the nucleotide columns of one block each."""
    z = np.load(SARS20_PATH)
    tree = parse_newick(bytes(z["newick"]).decode())
    nb = int(z["n_blocks"][0])
    batches = [dict(id="blocks", block=1, codes=z["blk_codes"], present=None, parent_code=np.zeros(nb, np.uint8),
                    root_override={0: None, 1: None},
                    expect={a: (MutLists(z[f"blk_a{a}_off"], z[f"blk_a{a}_pos"], z[f"blk_a{a}_tc"]), z[f"blk_a{a}_states"]) for a in (0, 1)})]
    for i in range(nb):
        p = f"b{i}_"
        batches.append(dict(id=f"block{i}", block=0, codes=z[p + "codes"], present=z[p + "present"], parent_code=z[p + "parent_code"],
                            root_override={0: z[p + "root_override"], 1: None}, col_j=z[p + "col_j"], col_k=z[p + "col_k"],
                            expect={a: (MutLists(z[p + f"a{a}_off"], z[p + f"a{a}_pos"], z[p + f"a{a}_tc"]), z[p + f"a{a}_states"])
                                    for a in (0, 1)}))
    return tree, batches


def concat_shards(parts):
    """Checker for the shard merge: parts = [(node_offsets, pos, type_code)] of contiguous column ranges in ascending order;
    a node's merged list is the concatenation of its per-range lists in range order (nothing is sorted)."""
    n_nodes = len(parts[0][0]) - 1
    cnt = np.stack([np.diff(np.asarray(o, np.int64)) for o, _, _ in parts])  # ranges x N
    off = np.zeros(n_nodes + 1, np.int64)
    off[1:] = np.cumsum(cnt.sum(0))
    pos = np.empty(int(off[-1]), np.int32)
    tc = np.empty(int(off[-1]), np.uint8)
    before = np.cumsum(cnt, 0) - cnt
    for k, (o, p, t) in enumerate(parts):
        if len(p) == 0:
            continue
        shift = off[:-1] + before[k] - np.asarray(o, np.int64)[:-1]
        idx = np.repeat(shift, cnt[k]) + np.arange(len(p))
        pos[idx] = p
        tc[idx] = t
    return off, pos, tc


def merge_all_nodes(port, node_offsets, pos, type_code):
    """The oracle's MSA run-merge (reference src/panman.cpp:1445-1466) applied to every node's list: returns
    (offsets int64[N+1], nucPosition, mutInfo, nucs, mutInfo on the wire) in the layout of pmb_merge_runs."""
    from oracle.oracle import wire_mut_info

    n_nodes = len(node_offsets) - 1
    off = np.zeros(n_nodes + 1, np.int64)
    ps, mis, nus = [], [], []
    for v in range(n_nodes):
        a, b = int(node_offsets[v]), int(node_offsets[v + 1])
        if b > a:
            p, mi, nu = port.merge_msa(pos[a:b], type_code[a:b])
            ps.append(p)
            mis.append(mi)
            nus.append(nu)
            off[v + 1] = len(p)
    off = np.cumsum(off)
    cat = lambda xs, t: np.concatenate(xs).astype(t) if xs else np.zeros(0, t)
    mi, nu = cat(mis, np.uint8), cat(nus, np.uint32)
    return off, cat(ps, np.int32), mi, nu, wire_mut_info(mi, nu).astype(np.uint32)


def writer_layout(nucmut, blockmut):
    """TEST INFRASTRUCTURE: restatement of what Tree::getNodesPreorder (reference src/panman.cpp:2854-2929) puts into one node's
    Mutation list. nucmut: [(nucPosition, nucGapPosition, primaryBlockId, secondaryBlockId, mutInfo, nucs)] in Node::nucMutation
    order; blockmut: [(primaryBlockId, secondaryBlockId, blockMutInfo, inversion)]. Returns [(blockId, blockGapExist,
    blockMutExist, blockMutInfo, blockInversion, [(nucPosition, nucGapPosition, nucGapExist, mutInfo)])] in std::map order."""
    groups = {}
    inversion = {}
    for pos, gap, pb, sb, info, nucs in nucmut:
        length = info >> 4
        wire = ((nucs >> (24 - length * 4)) << 8) + info                                     # :2876
        entry = (pos, gap if gap != -1 else 0, 1 if gap != -1 else 0, wire)                   # :2868-2874
        g = groups.setdefault((pb, sb), [[], 0])
        g[0].append(entry)
        g[1] = 2                                                                               # :2878
    for pb, sb, info, inv in blockmut:
        g = groups.setdefault((pb, sb), [[], 0])
        g[1] = int(bool(info))                                                                 # :2883
        inversion[(pb, sb)] = int(bool(inv))
    out = []
    for (pb, sb), (entries, second) in sorted(groups.items()):                                 # std::map order
        exist = int(second != 2)
        inv = inversion.get((pb, sb), 0) if second != 2 else 1                                 # :2893-2897
        block_id = (pb << 32) + sb if sb != -1 else (pb << 32)                                 # :2901-2908
        out.append((block_id, int(sb != -1), exist, int(bool(second)), inv, entries))
    return out
