"""Generates the committed golden vectors from the CPU oracle (run in the build container:
`python tests/golden/make_golden.py`).  The reference itself cannot be executed here (no Julia), so
these freeze the oracle restatement, not outputs of the Julia code - see DESIGN.md "Parity unpinned"."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import mgb_b200  # noqa: E402
from helpers import problem, oracle_eval  # noqa: E402

CASES = [("fem1d", 3, 1.0, False, None, 1.3), ("fem2d", 2, 2.0, False, None, 0.7), ("fem2d", 3, 1.0, False, None, 0.7),
         ("fem2d", 3, 1.5, False, 1, 2.0), ("fem2d", 2, 1.0, True, None, 0.5),
         # 64-node hexahedra (fem3d k=3, operator table u.id u.dx u.dy u.dz s.id): the CSR path, fine and coarse level
         ("fem3d", 2, 1.0, False, None, 0.9), ("fem3d", 2, 1.5, False, 0, 0.9)]


def name(gen, L, p, slack, level, t):
    return f"{gen}_L{L}_p{p}_{'slack' if slack else 'main'}_lev{'F' if level is None else level}.npz"


if __name__ == "__main__":
    for gen, L, p, slack, level, t in CASES:
        if os.path.exists(os.path.join(HERE, name(gen, L, p, slack, level, t))) and "--force" not in sys.argv:
            continue   # committed fixtures stay as they are
        pr = problem(getattr(mgb_b200, gen)(L), p=p, slack=slack, level=level)
        f0, g, H = oracle_eval(pr, t)
        H.sort_indices()
        np.savez_compressed(os.path.join(HERE, name(gen, L, p, slack, level, t)), s=pr["s"], t=t, f0=f0, grad=g,
                            H_data=H.data, H_indices=H.indices, H_indptr=H.indptr, m=H.shape[0])
        print(name(gen, L, p, slack, level, t), H.nnz)
