"""fem2d L: gather kernel time by what it is asked to produce (flags 7 = all, 5 = Hessian only, 3 = gradient only, 1 = scalars)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"):
    sys.path.insert(0, p)
import numpy as np, torch
from bench import build_problem
from mgb_b200 import capi
L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
ctx = capi.Context(0, stream.cuda_stream)
pr = build_problem(L, 1.0); geom = pr["geom"]
plan = capi.Plan(ctx, pr["D"], pr["R"], geom.x, geom.w, pr["idx"], 1.0)
cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
s_d = torch.from_numpy(pr["s"]).to(dev); Dz0_d = cm(pr["Dz0"]); c_d = cm(pr["c"])
scal = torch.zeros(4, dtype=torch.float64, device=dev); grad = torch.zeros(plan.m, dtype=torch.float64, device=dev)
hval = torch.zeros(plan.nnzH, dtype=torch.float64, device=dev)
out = {}
for fl in (7, 5, 3, 1, 7):
    plan.time_assemble(s_d, Dz0_d, c_d, 1.0, fl, scal, grad, hval, 3, 2, split=True)
    _, a, b = plan.time_assemble(s_d, Dz0_d, c_d, 1.0, fl, scal, grad, hval, 30, 2, split=True)
    t, _, _ = plan.time_assemble(s_d, Dz0_d, c_d, 1.0, fl, scal, grad, hval, 30, 2, split=False)
    out[f"flags{fl}" + ("b" if f"flags{fl}" in out else "")] = dict(total_us=t * 1e3, element_us=a * 1e3, gather_us=b * 1e3)
print(json.dumps(out))
