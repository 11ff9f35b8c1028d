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
