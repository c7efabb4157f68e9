"""fem3d: one assembly on every level of the hierarchy (dense path on the coarse levels, CSR path on the finest)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"):
    sys.path.insert(0, p)
import numpy as np, torch
import mgb_b200
from mgb_b200 import capi
from helpers import problem
L = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ctx = capi.Context(0); dev = torch.device("cuda", 0)
geom = mgb_b200.fem3d(L, k=3)
res = []
for lev in range(L):
    pr = problem(geom, level=lev)
    t0 = time.time()
    plan = capi.Plan(ctx, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], pr["p"])
    tplan = time.time() - t0
    Dz0 = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)
    cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
    s_d = torch.from_numpy(pr["s"]).to(dev); Dz0_d = cm(Dz0); c_d = cm(pr["c"])
    scal = torch.zeros(4, dtype=torch.float64, device=dev); grad = torch.zeros(plan.m, dtype=torch.float64, device=dev)
    hval = torch.zeros(max(plan.nnzH, 1), dtype=torch.float64, device=dev)
    plan.time_assemble(s_d, Dz0_d, c_d, 1.0, 7, scal, grad, hval, 2, 2, split=False)
    ms, _, _ = plan.time_assemble(s_d, Dz0_d, c_d, 1.0, 7, scal, grad, hval, 5, 2, split=False)
    msf, _, _ = plan.time_assemble(s_d, Dz0_d, c_d, 1.0, 1, scal, grad, hval, 5, 2, split=False)
    rec = dict(mesh=f"fem3d L={L}", n=int(geom.x.shape[0]), level=lev, path="dense" if plan.info["nodes_per_element"] == 64 else ("csr" if plan.info["path"] == 2 else "element"),
               m=plan.m, nnzH=plan.nnzH, chunks=plan.info["elements"], contribs=plan.info["hess_contribs"], plan_MB=plan.info["plan_bytes"] / 1e6,
               plan_s=tplan, assembly_ms=ms, f0_ms=msf)
    print(json.dumps(rec), flush=True)
    res.append(rec)
    plan.close()
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"r2_fem3d_levels_L{L}.json"), "w"), indent=1)
