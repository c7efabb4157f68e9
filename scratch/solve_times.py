"""Whole-solve timing split (GPU assembly vs the host sparse LU that stands in for MUMPS), fem2d p=1, L = 3..Lmax.
The reference publishes only whole-solve seconds on an M4 laptop (docs/src/guide.md:246-253)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"):
    sys.path.insert(0, p)
import numpy as np
import mgb_b200
from mgb_b200 import solver
Lmax = int(sys.argv[1]) if len(sys.argv) > 1 else 6
published = {1: 0.029, 2: 0.039, 3: 0.078, 4: 0.410, 5: 1.771, 6: 68.846, 7: 118.070, 8: 504.672}
out = []
for L in range(3, Lmax + 1):
    geom = mgb_b200.fem2d(L)
    t0 = time.perf_counter()
    sol = solver.amgb(geom, p=1.0)
    wall = time.perf_counter() - t0
    st = sol.stats
    rec = dict(L=L, n=int(geom.x.shape[0]), wall_s=wall, assemble_s=st["assemble_s"], solve_s=st["solve_s"],
               assemblies=st["assemblies"], f0_evals=st["f0_evals"], newton_its=int(np.sum(sol.SOL_main["its"])),
               t_steps=int(len(sol.SOL_main["ts"])), reference_published_M4_s=published.get(L))
    print(json.dumps(rec), flush=True)
    out.append(rec)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "solve_times.json"), "w"), indent=1)
