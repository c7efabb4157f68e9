"""Assembly time on every multigrid level (north_star item 4: the coarse levels run the same numeric-only machinery).
A/B of the chunked coarse-level gather: MGB_GATHER_CHUNK = 0 (one warp per output) vs 128 / 512."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"):
    sys.path.insert(0, p)
import numpy as np, torch
import mgb_b200
from mgb_b200 import capi
from helpers import problem
ctx = capi.Context(0); dev = torch.device("cuda", 0)
out = []
for gen, L, pert in (("fem2d", 8, 1e-3), ("fem1d", 16, 1e-8)):
    geom = getattr(mgb_b200, gen)(L)
    ref = {}
    for chunk in (sys.argv[1:] or ["0", "128"]):
        os.environ["MGB_GATHER_CHUNK"] = chunk
        for lev in range(L):
            pr = problem(geom, level=lev, pert=pert)
            plan = capi.Plan(ctx, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0)
            Dz0 = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)
            cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
            s_d = torch.from_numpy(pr["s"]).to(dev); Dz0_d = cm(Dz0); c_d = cm(pr["c"])
            scal = torch.zeros(4, dtype=torch.float64, device=dev); grad = torch.zeros(plan.m, dtype=torch.float64, device=dev)
            hval = torch.full((max(plan.nnzH, 1),), float("nan"), dtype=torch.float64, device=dev)
            plan.time_assemble(s_d, Dz0_d, c_d, 1.0, 7, scal, grad, hval, 3, 2, split=False)
            ms, _, _ = plan.time_assemble(s_d, Dz0_d, c_d, 1.0, 7, scal, grad, hval, 20, 2, split=False)
            cur = (hval.cpu().numpy(), grad.cpu().numpy())
            key = (gen, L, lev)
            ref.setdefault(key, cur)
            err = max(np.abs(cur[0] - ref[key][0]).max() / np.abs(ref[key][0]).max(), np.abs(cur[1] - ref[key][1]).max() / np.abs(ref[key][1]).max())
            rec = dict(mesh=f"{gen} L={L}", level=lev, m=plan.m, nnzH=plan.nnzH, contribs=plan.info["hess_contribs"], chunk=int(chunk), ms=ms, rel_diff_vs_unchunked=float(err))
            print(json.dumps(rec), flush=True)
            out.append(rec)
            plan.close()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "level_times.json"), "w"), indent=1)
