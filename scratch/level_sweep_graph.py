"""All-level sweep (north_star item 4): one assembly on every level of a hierarchy, (a) as plain calls, (b) recorded
once with mgb_graph_begin/_end and replayed with a single cudaGraphLaunch.  Per-level times with L2 flush, sweep
times without (back-to-back, as inside a multigrid cycle)."""
import json, os, sys, time
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
res = []
for gen, L, pert in (("fem2d", 8, 1e-3), ("fem1d", 16, 1e-8), ("fem2d", 3, 1e-3)):
    geom = getattr(mgb_b200, gen)(L)
    levels = []
    per_level = []
    for lev in range(L):
        pr = problem(geom, level=lev, pert=pert)
        plan = capi.Plan(ctx, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0)
        Dz0 = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)
        cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
        s_d = torch.from_numpy(pr["s"]).to(dev); Dz0_d = cm(Dz0); c_d = cm(pr["c"])
        scal = torch.zeros(4, dtype=torch.float64, device=dev); grad = torch.zeros(plan.m, dtype=torch.float64, device=dev)
        hval = torch.zeros(max(plan.nnzH, 1), dtype=torch.float64, device=dev)
        plan.time_assemble(s_d, Dz0_d, c_d, 1.0, 7, scal, grad, hval, 3, 2, split=False)
        ms, _, _ = plan.time_assemble(s_d, Dz0_d, c_d, 1.0, 7, scal, grad, hval, 20, 2, split=False)
        per_level.append(ms * 1e3)
        levels.append((plan, s_d, Dz0_d, c_d, scal, grad, hval))
    def sweep():
        for plan, s_d, Dz0_d, c_d, scal, grad, hval in levels:
            plan.assemble(s_d, Dz0_d, c_d, 1.0, 7, scal, grad, hval)
    def timed(fn, reps=30):
        for _ in range(5): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e3
    us_calls = timed(sweep)             # each call replays its own cached graph
    os.environ["MGB_GRAPH"] = "0"
    ctx.graph_begin(); sweep(); g = ctx.graph_end()
    us_graph = timed(g.launch)
    rec = dict(mesh=f"{gen} L={L}", levels=L, per_level_us_flushed=per_level, sum_per_level_us=sum(per_level),
               sweep_us_calls_back_to_back=us_calls, sweep_us_one_graph=us_graph)
    print(json.dumps(rec), flush=True)
    res.append(rec)
    g.close()
    for lv in levels: lv[0].close()
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "r2_level_sweep.json"), "w"), indent=1)
