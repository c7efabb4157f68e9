"""time the finest-level assembly at the TRUE iterates of a full solve (first, middle, last Newton step
on the finest level) next to the synthetic iterate bench.py uses.  fem2d L=6 p=1 (a full L=8 solve is >30 min of
host LU); the kernels have no data-dependent control flow, so the times must agree - this run is the evidence."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"):
    sys.path.insert(0, p)
import numpy as np, torch
import mgb_b200
from mgb_b200 import capi, solver
from helpers import problem

L = int(sys.argv[1]) if len(sys.argv) > 1 else 6
geom = mgb_b200.fem2d(L)
rec = []
orig = solver.LevelState.assemble
J_fine = len(geom.refine) if hasattr(geom, "refine") else None

def spy(self, s, Dz0, c, t, flags):
    if flags == 7 or flags == 6:
        rec.append((self.m, s.detach().clone(), Dz0.detach().clone(), float(t)))
    return orig(self, s, Dz0, c, t, flags)

solver.LevelState.assemble = spy
t0 = time.time()
sol = solver.amgb(geom, p=1.0)
print("solve s", time.time() - t0, "Newton its", int(sol.SOL_main["its"].sum()), flush=True)
solver.LevelState.assemble = orig
mfine = max(r[0] for r in rec)
fine = [r for r in rec if r[0] == mfine]
picks = {"first": fine[0], "middle": fine[len(fine) // 2], "last": fine[-1]}
ctx = capi.Context(0); dev = torch.device("cuda", 0)
pr = problem(geom)
plan = capi.Plan(ctx, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0)
cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
c_d = cm(pr["c"])
scal = torch.zeros(4, dtype=torch.float64, device=dev); grad = torch.zeros(plan.m, dtype=torch.float64, device=dev)
hval = torch.zeros(plan.nnzH, dtype=torch.float64, device=dev)
out = {"config": f"fem2d L={L} p=1.0", "n": geom.x.shape[0], "fine_assemblies_in_solve": len(fine), "rows": []}
Dz0s = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)
cases = [("synthetic (bench.py iterate)", torch.from_numpy(pr["s"]).to(dev), cm(Dz0s), 1.0)] + \
        [(k + " Newton step of the solve", v[1], v[2], v[3]) for k, v in picks.items()]
for name, s_d, Dz0_d, t in cases:
    plan.time_assemble(s_d, Dz0_d, c_d, t, 7, scal, grad, hval, 5, 2, split=False)
    ms, _, _ = plan.time_assemble(s_d, Dz0_d, c_d, t, 7, scal, grad, hval, 50, 2, split=False)
    row = dict(iterate=name, t=t, ms_assembly=ms, all_finite=float(scal.cpu()[1]))
    print(json.dumps(row), flush=True)
    out["rows"].append(row)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "iterate_times.json"), "w"), indent=1)
