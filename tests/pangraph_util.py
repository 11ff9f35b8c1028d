"""TEST INFRASTRUCTURE: a Python restatement of how the reference's PanGraph path turns the JSON into per-block column
batches (src/panman.cpp:6216-6258 JSON fields; :1006-1045 aligned strings; :873-963 and :1048-1232 column drivers), shared
by the golden-fixture generator (tests/golden/make_sars20_golden.py) and the tests of the C++ adaptor
(panman_b200/host/pangraph.cpp), plus a generator of random PanGraphs in the same JSON layout.

Stand-ins (documented in both places): block order = JSON order; root override of Fitch main columns without --reference =
the present sequence with the highest leaf row; gap columns / the Sankoff branch have none without --reference."""
import json

import numpy as np

from oracle.oracle import CODE_OF


def code_of(ch: str) -> int:
    return int(CODE_OF[ord(ch)])  # anything but the 15 IUPAC letters (incl. '-') -> 0, as getCodeFromNucleotide


def build_batches(pg: dict, tree):
    """Returns (block_states uint8 [n_leaves, n_blocks], [batch per block]); a batch has codes (uint8 [n_leaves, n_cols], one
    code per byte), present, parent_code, root_override (Fitch stand-in), col_j, col_k."""
    row_of_name = {tree.names[v]: int(tree.leaf_row[v]) for v in tree.leaves}
    n_leaves = tree.n_leaves
    has_block = {}
    for path in pg["paths"]:
        assert path["name"] in row_of_name and not path["circular"]
        for b in path["blocks"]:
            assert b["number"] == 1 and path["name"] not in has_block.setdefault(b["id"], {})
            has_block[b["id"]][path["name"]] = bool(b["strand"])
    bcodes = np.zeros((n_leaves, len(pg["blocks"])), np.uint8)
    batches = []
    for i, blk in enumerate(pg["blocks"]):
        for name, strand in has_block.get(blk["id"], {}).items():
            bcodes[row_of_name[name], i] = 1 if strand else 2
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

        def per_seq(field):
            return {e[0]["name"]: e[1] for e in blk[field] if e[0]["number"] == 1}

        subs, ins, dels = per_seq("mutate"), per_seq("insert"), per_seq("delete")
        for name in has_block.get(blk["id"], {}):
            r = row_of_name[name]
            present[r] = 1
            row = [code_of(ch) for ch in main] + [0] * (n_cols - (L + 1))
            for pos, ch in subs.get(name, []):
                row[pos - 1] = code_of(ch.upper()[0])
            for (pos, off), s in ins.get(name, []):
                for t, ch in enumerate(s.upper()):
                    row[gap_col[(pos, off + t)]] = code_of(ch)
            for pos, ln in dels.get(name, []):
                for j in range(pos, pos + ln):
                    row[j - 1] = 0
            codes[r] = row
        parent_code = np.asarray([code_of(ch) for ch in main] + [0] * (n_cols - (L + 1)), np.uint8)
        root_override = np.full(n_cols, -1, np.int8)
        if present.any():
            last_row = int(np.nonzero(present)[0].max())
            root_override[:L + 1] = codes[last_row, :L + 1]
        batches.append(dict(id=blk["id"], codes=codes, present=present, parent_code=parent_code, root_override=root_override,
                            col_j=np.asarray(col_j, np.int32), col_k=np.asarray(col_k, np.int32)))
    return bcodes, batches


def random_pangraph(tree, rng, n_blocks=4, max_len=400) -> str:
    """A random PanGraph in the reference's JSON layout over the leaves of `tree` (every leaf gets a path)."""
    names = [tree.names[v] for v in tree.leaves]
    alphabet = "ACGT"
    blocks, owners = [], []
    for b in range(n_blocks):
        L = int(rng.integers(5, max_len))
        cons = "".join(rng.choice(list(alphabet), size=L))
        gaps = {}
        for _ in range(int(rng.integers(0, 4))):
            gaps[str(int(rng.integers(0, L + 1)))] = int(rng.integers(1, 9))
        own = [n for n in names if rng.random() < (1.0 if b == 0 else 0.7)] or [names[0]]  # a block nobody owns would trip the reference's assert
        mutate, insert, delete = [], [], []
        for n in own:
            who = {"name": n, "number": 1, "strand": True}
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
        pb = [{"id": f"BLK{b:03d}", "name": n, "number": 1, "strand": bool(rng.random() < 0.8)} for b in range(n_blocks) if n in owners[b]]
        paths.append({"name": n, "offset": None, "circular": False, "position": [], "blocks": pb})
    return json.dumps({"paths": paths, "blocks": blocks})
