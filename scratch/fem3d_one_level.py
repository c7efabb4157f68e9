"""fem3d L: one level's dense assembly timed (full / objective only); MGB_B200_LIB selects a variant library"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"):
    sys.path.insert(0, p)
import numpy as np, torch
import mgb_b200
from mgb_b200 import capi
from helpers import problem
L, lev = int(sys.argv[1]), int(sys.argv[2])
ctx = capi.Context(0); dev = torch.device("cuda", 0)
geom = mgb_b200.fem3d(L, k=3)
pr = problem(geom, level=lev)
plan = capi.Plan(ctx, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], pr["p"])
Dz0 = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)
cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
s_d = torch.from_numpy(pr["s"]).to(dev); Dz0_d = cm(Dz0); c_d = cm(pr["c"])
scal = torch.zeros(4, dtype=torch.float64, device=dev); grad = torch.zeros(plan.m, dtype=torch.float64, device=dev)
hval = torch.zeros(max(plan.nnzH, 1), dtype=torch.float64, device=dev)
out = {}
for name, fl in (("full", 7), ("f0", 1), ("f0+grad", 3)):
    plan.time_assemble(s_d, Dz0_d, c_d, 1.0, fl, scal, grad, hval, 2, 2, split=False)
    ms, a, b = plan.time_assemble(s_d, Dz0_d, c_d, 1.0, fl, scal, grad, hval, 5, 2, split=True)
    ms, _, _ = plan.time_assemble(s_d, Dz0_d, c_d, 1.0, fl, scal, grad, hval, 5, 2, split=False)
    out[name] = dict(ms=ms, element_ms=a, gather_ms=b)
print(json.dumps(dict(lib=os.environ.get("MGB_B200_LIB", "default"), L=L, level=lev, **out)))
