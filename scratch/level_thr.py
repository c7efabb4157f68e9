"""threshold between the thread-per-entry gather and the lanes-per-output gather (MGB_LONG_AVG): A/B on the levels
whose lists average 5..20 contributions"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"):
    sys.path.insert(0, p)
import numpy as np, torch
import mgb_b200
from mgb_b200 import capi
from helpers import problem
ctx = capi.Context(0); dev = torch.device("cuda", 0)
for gen, L, pert, levels in (("fem2d", 8, 1e-3, (5, 6, 7)), ("fem1d", 16, 1e-8, (11, 12, 13, 14))):
    geom = getattr(mgb_b200, gen)(L)
    for lev in levels:
        pr = problem(geom, level=lev, pert=pert)
        Dz0 = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)
        cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
        s_d = torch.from_numpy(pr["s"]).to(dev); Dz0_d = cm(Dz0); c_d = cm(pr["c"])
        ref = None
        for thr in ("12", "6", "3"):
            os.environ["MGB_LONG_AVG"] = thr
            plan = capi.Plan(ctx, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0)
            scal = torch.zeros(4, dtype=torch.float64, device=dev); grad = torch.zeros(plan.m, dtype=torch.float64, device=dev)
            hval = torch.full((max(plan.nnzH, 1),), float("nan"), dtype=torch.float64, device=dev)
            plan.time_assemble(s_d, Dz0_d, c_d, 1.0, 7, scal, grad, hval, 3, 2, split=False)
            ms, _, _ = plan.time_assemble(s_d, Dz0_d, c_d, 1.0, 7, scal, grad, hval, 20, 2, split=False)
            cur = hval.cpu().numpy()
            ref = cur if ref is None else ref
            print(json.dumps(dict(mesh=f"{gen} L={L}", level=lev, avg=round(plan.info["hess_contribs"] / plan.nnzH, 1), thr=thr, us=round(ms * 1e3, 1),
                                  rel=float(np.abs(cur - ref).max() / np.abs(ref).max()))), flush=True)
            plan.close()
