"""TEST INFRASTRUCTURE: a Python restatement of how the reference's PanGraph path turns the JSON into per-block column
batches (src/panman.cpp:6216-6258 JSON fields; :1006-1045 aligned strings; :873-963 and :1048-1232 column drivers), shared
by the golden-fixture generator (tests/golden/make_sars20_golden.py) and the tests of the C++ adaptor
(panman_b200/host/pangraph.cpp), plus a generator of random PanGraphs in the same JSON layout.

Block columns come from the reference's own compiled ordering code (oracle.RefPgOrder). Which sequence the root is forced to
where several qualify follows the walk order of the reference's maps (panman_b200/host/pangraph.cpp, header comment): the
std::unordered_map walk of the block-level driver is taken from the compiled reference containers (order["aligned_walk"]);
the tbb::concurrent_unordered_map walk of the nucleotide-level driver is restated here (bit-reversed tbb_hasher) and is the
one rule of the flow that no executable reference code pins (TBB is not installed)."""
import json

import numpy as np

from oracle.oracle import CODE_OF


def tbb_walk_key(name: str) -> int:
    """Position of a std::string key in the walk of a tbb::concurrent_unordered_map (split-ordered list: ascending bit-reversed
    hash; tbb_hasher: h = c ^ (h * 0x9E3779B97F4A7C15))."""
    h = 0
    for ch in name.encode():
        c = ch - 256 if ch >= 128 else ch  # char is signed
        h = ((c & 0xFFFFFFFFFFFFFFFF) ^ ((h * 11400714819323198485) & 0xFFFFFFFFFFFFFFFF)) & 0xFFFFFFFFFFFFFFFF
    return int(format(h, "064b")[::-1], 2) | 1


def code_of(ch: str) -> int:
    return int(CODE_OF[ord(ch)])  # anything but the 15 IUPAC letters (incl. '-') -> 0, as getCodeFromNucleotide


def json_order(pg: dict):
    """The block order of a PanGraph WITHOUT duplicated blocks, circular paths or rearrangements, computed without the
    reference: every path is then a subsequence of one common order... which is NOT in general the JSON order. Only used
    where the reference's compiled ordering code is unavailable AND every path lists its blocks in JSON order."""
    ids = [b["id"] for b in pg["blocks"]]
    at = {x: i for i, x in enumerate(ids)}
    out = dict(topo_ids=ids, aligned={}, strand={}, number={}, rotation_index={})
    for p in pg["paths"]:
        if not p["blocks"]:
            continue
        a = np.full(len(ids), -1, np.int32)
        s_ = np.full(len(ids), -1, np.int32)
        nu = np.zeros(len(ids), np.int32)
        for b in p["blocks"]:
            a[at[b["id"]]] = at[b["id"]]
            s_[at[b["id"]]] = 1 if b["strand"] else 0
            nu[at[b["id"]]] = 1
        out["aligned"][p["name"]], out["strand"][p["name"]], out["number"][p["name"]] = a, s_, nu
        out["rotation_index"][p["name"]] = 0
    return out


def build_batches(pg: dict, tree, order: dict, reference: str = ""):
    """Returns (block_states uint8 [n_leaves, n_cols_blocks], [batch per block column]); a batch has codes (uint8 [n_leaves,
    n_cols], one code per byte), present, parent_code, root_override (Fitch stand-in), col_j, col_k. `order` = the block
    columns and per-path ownership (oracle.RefPgOrder.order: the reference's chain_align / rotation code)."""
    row_of_name = {tree.names[v]: int(tree.leaf_row[v]) for v in tree.leaves}
    name_of_row = {r: n for n, r in row_of_name.items()}
    n_leaves = tree.n_leaves
    by_id = {b["id"]: b for b in pg["blocks"]}
    topo = order["topo_ids"]
    bcodes = np.zeros((n_leaves, len(topo)), np.uint8)
    batches = []
    for i, bid in enumerate(topo):
        blk = by_id[bid]
        owners = {}  # name -> occurrence number
        for name, al in order["aligned"].items():
            if name in row_of_name and al[i] != -1:
                owners[name] = int(order["number"][name][i])
                bcodes[row_of_name[name], i] = 1 if order["strand"][name][i] else 2
        cons = blk["sequence"].upper()
        L = len(cons)
        gaps = sorted((int(k), int(v)) for k, v in blk["gaps"].items())
        main = [cons[j] if j < L else "-" for j in range(L + 1)]
        col_j = list(range(L + 1)) + [j for j, g in gaps for _ in range(g)]
        col_k = [-1] * (L + 1) + [k for _, g in gaps for k in range(g)]
        gap_col = {(col_j[c], col_k[c]): c for c in range(L + 1, len(col_j))}
        n_cols = len(col_j)
        codes = np.zeros((n_leaves, n_cols), np.uint8)
        present = np.zeros(n_leaves, np.uint8)

        def per_seq(field, name, number):
            out = []
            for e in blk[field]:
                if e[0]["name"] == name and e[0]["number"] == number:
                    out += e[1]
            return out

        for name, number in owners.items():
            r = row_of_name[name]
            present[r] = 1
            row = [code_of(ch) for ch in main] + [0] * (n_cols - (L + 1))
            for pos, ch in per_seq("mutate", name, number):
                row[pos - 1] = code_of(ch.upper()[0])
            for (pos, off), s_ in per_seq("insert", name, number):
                for t, ch in enumerate(s_.upper()):
                    row[gap_col[(pos, off + t)]] = code_of(ch)
            for pos, ln in per_seq("delete", name, number):
                for j in range(pos, pos + ln):
                    row[j - 1] = 0
            codes[r] = row
        parent_code = np.asarray([code_of(ch) for ch in main] + [0] * (n_cols - (L + 1)), np.uint8)
        # the owner the nucleotide-level driver walks last (individualSequences, a tbb::concurrent_unordered_map)
        root_override = np.full(n_cols, -1, np.int8)
        owners_rows = [int(r) for r in np.nonzero(present)[0]]
        if reference:
            match = [r for r in owners_rows if reference in name_of_row[r]]
            if match:
                root_override[:] = codes[max(match, key=lambda r: tbb_walk_key(name_of_row[r]))]
        elif owners_rows:  # src/panman.cpp:1132: the unguarded find("") of the Fitch main-column branch
            last_row = max(owners_rows, key=lambda r: tbb_walk_key(name_of_row[r]))
            root_override[:L + 1] = codes[last_row, :L + 1]
        batches.append(dict(id=bid, codes=codes, present=present, parent_code=parent_code, root_override=root_override,
                            col_j=np.asarray(col_j, np.int32), col_k=np.asarray(col_k, np.int32)))
    if reference:  # block level (src/panman.cpp:881-897): the last match in the walk of alignedSequences
        block_override = np.full(len(topo), -1, np.int8)
        for name in order["aligned_walk"]:
            if name in row_of_name and reference in name:
                block_override[:] = bcodes[row_of_name[name]]
        return bcodes, batches, block_override
    return bcodes, batches


def random_pangraph(tree, rng, n_blocks=4, max_len=400, duplicates=False, circular=False, shuffle=False) -> str:
    """A random PanGraph in the reference's JSON layout over the leaves of `tree` (every leaf gets a path). duplicates: some
    paths carry a block twice (occurrence numbers 1, 2 with their own mutations); circular: paths are circular and start at
    random blocks; shuffle: some paths list their blocks in another order (rearrangements)."""
    names = [tree.names[v] for v in tree.leaves]
    alphabet = "ACGT"
    blocks, owners = [], []
    dup = {}
    for b in range(n_blocks):
        L = int(rng.integers(5, max_len))
        cons = "".join(rng.choice(list(alphabet), size=L))
        gaps = {}
        for _ in range(int(rng.integers(0, 4))):
            gaps[str(int(rng.integers(0, L + 1)))] = int(rng.integers(1, 9))
        own = [n for n in names if rng.random() < (1.0 if b == 0 else 0.7)] or [names[0]]  # a block nobody owns would trip the reference's assert
        mutate, insert, delete = [], [], []
        for n in own:
            copies = 2 if duplicates and rng.random() < 0.3 else 1
            dup[(b, n)] = copies
            for number in range(1, copies + 1):
                who = {"name": n, "number": number, "strand": True}
                subs = [[int(rng.integers(1, L + 1)), str(rng.choice(list("ACGTNRY")))] for _ in range(int(rng.integers(0, 6)))]
                inss = []
                for pos, g in gaps.items():
                    if rng.random() < 0.5:
                        off = int(rng.integers(0, g))
                        ln = int(rng.integers(1, g - off + 1))
                        inss.append([[int(pos), off], "".join(rng.choice(list(alphabet), size=ln))])
                dl = []
                if rng.random() < 0.4:
                    p = int(rng.integers(1, L + 1))
                    dl.append([p, int(rng.integers(1, min(6, L - p + 1) + 1))])
                mutate.append([who, subs])
                insert.append([who, inss])
                delete.append([who, dl])
        blocks.append({"id": f"BLK{b:03d}", "sequence": cons.lower() if b % 2 else cons, "gaps": gaps, "mutate": mutate,
                       "insert": insert, "delete": delete, "positions": []})
        owners.append(set(own))
    paths = []
    for n in names:
        order = [b for b in range(n_blocks) if n in owners[b]]
        if shuffle and rng.random() < 0.4 and len(order) > 2:
            i, j = sorted(rng.choice(len(order), size=2, replace=False))
            order[i:j + 1] = order[i:j + 1][::-1]
        seq = []
        for b in order:
            seq.append(b)
            if dup[(b, n)] == 2:  # the second copy right behind the first, or at the end of the path
                if rng.random() < 0.5:
                    seq.append(b)
                else:
                    order_tail = True
                    seq.append(-b - 1)
        tail = [-(x + 1) for x in seq if x < 0]
        seq = [x for x in seq if x >= 0] + tail
        seen = {}
        pb = []
        for b in seq:
            seen[b] = seen.get(b, 0) + 1
            pb.append({"id": f"BLK{b:03d}", "name": n, "number": seen[b], "strand": bool(rng.random() < 0.8)})
        circ = bool(circular)
        if circ and len(pb) > 1:
            k = int(rng.integers(0, len(pb)))
            pb = pb[k:] + pb[:k]
            # occurrence numbers count along the path as it is written
            seen = {}
            for e in pb:
                seen[e["id"]] = seen.get(e["id"], 0) + 1
                e["number"] = seen[e["id"]]
        paths.append({"name": n, "offset": 0 if circ else None, "circular": circ, "position": [], "blocks": pb})
    return json.dumps({"paths": paths, "blocks": blocks})
