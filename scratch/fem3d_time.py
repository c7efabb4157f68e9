"""fem3d L=5 assembly (CSR path): for an ncu launch list."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"):
    sys.path.insert(0, p)
import numpy as np, torch
import mgb_b200
from mgb_b200 import capi
from helpers import problem
L = int(sys.argv[1]) if len(sys.argv) > 1 else 5
ctx = capi.Context(0); dev = torch.device("cuda", 0)
geom = mgb_b200.fem3d(L, k=3)
pr = problem(geom)
plan = capi.Plan(ctx, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], pr["p"])
print(plan.info)
Dz0 = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)
cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
s_d = torch.from_numpy(pr["s"]).to(dev); Dz0_d = cm(Dz0); c_d = cm(pr["c"])
scal = torch.zeros(4, dtype=torch.float64, device=dev); grad = torch.zeros(plan.m, dtype=torch.float64, device=dev)
hval = torch.zeros(max(plan.nnzH, 1), dtype=torch.float64, device=dev)
ms, _, _ = plan.time_assemble(s_d, Dz0_d, c_d, 1.0, 7, scal, grad, hval, 5, 2, split=False)
print("ms", ms)
