"""fem2d L=8: element / gather split per level"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"):
    sys.path.insert(0, p)
import numpy as np, torch
import mgb_b200
from mgb_b200 import capi
from helpers import problem
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
ctx = capi.Context(0, stream.cuda_stream)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
geom = mgb_b200.fem2d(L)
for lev in range(L):
    pr = problem(geom, level=lev, pert=1e-3)
    plan = capi.Plan(ctx, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0)
    Dz0 = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)
    cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
    s_d = torch.from_numpy(pr["s"]).to(dev); Dz0_d = cm(Dz0); c_d = cm(pr["c"])
    scal = torch.zeros(4, dtype=torch.float64, device=dev); grad = torch.zeros(plan.m, dtype=torch.float64, device=dev)
    hval = torch.zeros(max(plan.nnzH, 1), dtype=torch.float64, device=dev)
    plan.time_assemble(s_d, Dz0_d, c_d, 1.0, 7, scal, grad, hval, 3, 2, split=False)
    ms, _, _ = plan.time_assemble(s_d, Dz0_d, c_d, 1.0, 7, scal, grad, hval, 20, 2, split=False)
    _, a, b = plan.time_assemble(s_d, Dz0_d, c_d, 1.0, 7, scal, grad, hval, 20, 2, split=True)
    msf, _, _ = plan.time_assemble(s_d, Dz0_d, c_d, 1.0, 1, scal, grad, hval, 20, 2, split=False)
    print(json.dumps(dict(level=lev, m=plan.m, nnzH=plan.nnzH, total_us=ms * 1e3, element_us=a * 1e3, gather_us=b * 1e3, f0_us=msf * 1e3,
                          contribs=plan.info["hess_contribs"], slots=plan.info["slots_per_element"])), flush=True)
    plan.close()
