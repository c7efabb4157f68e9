"""one multigrid level's assembly, a few repetitions (for an ncu launch list): python level_one.py fem2d 8 3"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"):
    sys.path.insert(0, p)
import numpy as np, torch
import mgb_b200
from mgb_b200 import capi
from helpers import problem
gen, L, lev = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ctx = capi.Context(0); dev = torch.device("cuda", 0)
pr = problem(getattr(mgb_b200, gen)(L), level=lev, pert=1e-8 if gen == "fem1d" else 1e-3)
plan = capi.Plan(ctx, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0)
Dz0 = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)
cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
s_d = torch.from_numpy(pr["s"]).to(dev); Dz0_d = cm(Dz0); c_d = cm(pr["c"])
scal = torch.zeros(4, dtype=torch.float64, device=dev); grad = torch.zeros(plan.m, dtype=torch.float64, device=dev)
hval = torch.zeros(max(plan.nnzH, 1), dtype=torch.float64, device=dev)
ms, _, _ = plan.time_assemble(s_d, Dz0_d, c_d, 1.0, 7, scal, grad, hval, 5, 2, split=False)
print(gen, L, lev, plan.info, "ms", ms)
