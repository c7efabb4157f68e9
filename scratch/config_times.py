"""Times one finest-level assembly for every BASELINE.json config (C1..C4) and checks small-size parity."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"):
    sys.path.insert(0, p)
import numpy as np, torch
import mgb_b200
from mgb_b200 import capi
from helpers import problem, check_against_oracle

ctx = capi.Context(0)
dev = torch.device("cuda", 0)
out = []
# parity at small sizes for the 3-D element types (CSR path)
for k in (1, 2, 3):
    plan, _ = check_against_oracle(ctx, mgb_b200.fem3d(2, k=k), 1.0, t=0.9)
    print("fem3d L=2 k=%d parity ok, path=%d m=%d nnzH=%d" % (k, plan.info["path"], plan.m, plan.nnzH), flush=True)
check_against_oracle(ctx, mgb_b200.fem3d(2, k=2), 2.0, t=0.9, level=0)
print("fem3d coarse level parity ok", flush=True)

def time_config(name, geom, p=1.0, reps=10, pert=1e-3):
    t0 = time.perf_counter()
    pr = problem(geom, p=p, pert=pert)
    t_setup = time.perf_counter() - t0
    t0 = time.perf_counter()
    plan = capi.Plan(ctx, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], pr["p"])
    t_plan = time.perf_counter() - t0
    n, nD = geom.x.shape[0], len(pr["D"])
    Dz0 = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)
    cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
    s_d = torch.from_numpy(pr["s"]).to(dev); Dz0_d = cm(Dz0); c_d = cm(pr["c"])
    scal = torch.zeros(4, dtype=torch.float64, device=dev); grad = torch.zeros(plan.m, dtype=torch.float64, device=dev)
    hval = torch.zeros(max(plan.nnzH, 1), dtype=torch.float64, device=dev)
    plan.time_assemble(s_d, Dz0_d, c_d, 1.0, 7, scal, grad, hval, 3, 2, split=False)
    ms, _, _ = plan.time_assemble(s_d, Dz0_d, c_d, 1.0, 7, scal, grad, hval, reps, 2, split=False)
    ms1, _, _ = plan.time_assemble(s_d, Dz0_d, c_d, 1.0, 1, scal, grad, hval, reps, 2, split=False)
    info = plan.info
    rec = dict(config=name, n=n, m=info["m"], nnzH=info["nnzH"], path="element" if info["path"] == 1 else "csr",
               ms_assembly=ms, ms_f0=ms1, alg_bytes=info["alg_bytes"], alg_GBs=info["alg_bytes"] / ms / 1e6,
               frac_of_6545=info["alg_bytes"] / ms / 1e6 / 6545.6, plan_s=t_plan, setup_s=t_setup,
               finite=float(scal.cpu()[1]))
    print(json.dumps(rec), flush=True)
    out.append(rec)
    plan.close()

time_config("C1 fem2d L=3 p=1", mgb_b200.fem2d(3))
time_config("C2 fem1d L=16", mgb_b200.fem1d(16), pert=1e-8)
time_config("C3 fem2d L=8 p=1", mgb_b200.fem2d(8))
time_config("C5-mesh fem2d L=7", mgb_b200.fem2d(7))
for L in (3, 4, 5):
    time_config("C4 fem3d L=%d k=3" % L, mgb_b200.fem3d(L, k=3), reps=5)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "config_times.json"), "w"), indent=1)
