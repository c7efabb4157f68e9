"""fem3d (CSR path) tuning: SELL sorting window x store mode; results must be bit-identical across variants."""
import os, sys, time, json
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
Dz0 = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)
cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
s_d = torch.from_numpy(pr["s"]).to(dev); Dz0_d = cm(Dz0); c_d = cm(pr["c"])
ref = None
rows = []
variants = [(16384, 16, 1024), (16384, 8, 1024), (16384, 32, 1024), (4096, 16, 1024), (65536, 16, 1024), (16384, 16, 256), (16384, 1000000, 1024)]
for sigma, stage, sg in variants:   # (Hessian sorting window, chunk length, gradient sorting window)
    os.environ["MGB_SELL_SIGMA"] = str(sigma); os.environ["MGB_SELL_CHUNK"] = str(stage); os.environ["MGB_SELL_SIGMA_GRAD"] = str(sg)
    t0 = time.time()
    plan = capi.Plan(ctx, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], pr["p"])
    tp = time.time() - t0
    scal = torch.zeros(4, dtype=torch.float64, device=dev); grad = torch.zeros(plan.m, dtype=torch.float64, device=dev)
    hval = torch.full((max(plan.nnzH, 1),), float("nan"), dtype=torch.float64, device=dev)
    ms, _, _ = plan.time_assemble(s_d, Dz0_d, c_d, 1.0, 7, scal, grad, hval, 20, 2, split=False)
    ms0, _, _ = plan.time_assemble(s_d, Dz0_d, c_d, 1.0, 1, scal, grad, hval, 20, 2, split=False)
    plan.assemble(s_d, Dz0_d, c_d, 1.0, 7, scal, grad, hval); ctx.sync()
    cur = (hval.cpu().numpy(), grad.cpu().numpy(), scal.cpu().numpy())
    if ref is None:
        ref = cur
    same = all(np.allclose(a, b, rtol=1e-13, atol=1e-13 * np.abs(b).max()) for a, b in zip(cur, ref))
    row = dict(L=L, sigma=sigma, chunk=stage, sigma_grad=sg, ms=ms, ms_f0=ms0, plan_s=tp, identical=bool(same), finite=bool(np.isfinite(cur[0]).all()),
               stored=plan.info["hess_stored"], contribs=plan.info["hess_contribs"], alg_bytes=plan.info["alg_bytes"],
               frac=plan.info["alg_bytes"] / (ms * 1e-3) / 1e9 / 6545.6)
    print(json.dumps(row), flush=True)
    rows.append(row)
    plan.close()
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"fem3d_tune_L{L}.json"), "w"), indent=1)
