"""event-timed assembly / objective-only call on fem2d level L (element kernel + gather split)"""
import sys, os, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mgb_b200 import capi
L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
pr = bench.build_problem(L, 1.0)
geom = pr["geom"]; n = geom.x.shape[0]
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
ctx = capi.Context(0, stream.cuda_stream)
plan = capi.Plan(ctx, pr["D"], pr["R"], geom.x, geom.w, pr["idx"], pr["p"])
s_d = torch.from_numpy(pr["s"]).to(dev)
Dz0_d = torch.from_numpy(np.asfortranarray(pr["Dz0"]).T.copy()).to(dev)
c_d = torch.from_numpy(np.asfortranarray(pr["c"]).T.copy()).to(dev)
scal = torch.zeros(4, dtype=torch.float64, device=dev); grad = torch.zeros(plan.m, dtype=torch.float64, device=dev)
hval = torch.zeros(plan.nnzH, dtype=torch.float64, device=dev)
out = {}
for name, flags in (("full", 7), ("f0", 1)):
    plan.time_assemble(s_d, Dz0_d, c_d, 1.0, flags, scal, grad, hval, 5, 2, split=False)
    tot, _, _ = plan.time_assemble(s_d, Dz0_d, c_d, 1.0, flags, scal, grad, hval, reps, 2, split=False)
    _, el, ga = plan.time_assemble(s_d, Dz0_d, c_d, 1.0, flags, scal, grad, hval, reps, 2, split=True)
    tot_nf, _, _ = plan.time_assemble(s_d, Dz0_d, c_d, 1.0, flags, scal, grad, hval, reps, 0, split=False)
    out[name] = dict(total_us=tot * 1e3, element_us=el * 1e3, gather_us=ga * 1e3, total_noflush_us=tot_nf * 1e3)
print(json.dumps(dict(L=L, n=n, info=plan.info, times=out)))
