"""small invocation of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
element + gather (fine, coarse, slack, two-cone), CSR path (fem3d), sharded plans on virtual ranks incl. the
peer-memory scalar exchange and the all-gather of a distributed unknown, reductions, spmat products."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import mgb_b200
from mgb_b200 import capi
from mgb_b200.hpc import uniform_partition
from helpers import problem, cuda_eval
import mgb_oracle as O

ctx = capi.Context(0)
dev = torch.device("cuda", 0)
for gen, L, kw in (("fem2d", 3, {}), ("fem2d", 3, dict(level=0)), ("fem2d", 2, dict(slack=True)), ("fem1d", 4, {}),
                   ("fem1d", 4, dict(level=1)), ("fem3d", 2, {})):
    pr = problem(getattr(mgb_b200, gen)(L), **kw)
    plan, out, H = cuda_eval(ctx, pr, 0.7)
    assert out["scal"][1] == 1.0
    plan.close()
# two cones
geom = mgb_b200.fem2d(2)
Dt, idxA, idxB = O.parabolic_tables(2)
M = O.amg_helper(geom, O.PARABOLIC_STATE, Dt)
u = np.sin(geom.x[:, 0]) + geom.x[:, 1] ** 2
z0 = O.parabolic_feasible_start(M, u, 2, 1.0)
R = M.R_fine[-1]
plan = capi.Plan(ctx, M.D, R, geom.x, geom.w, idxB, 1.0, idx2=idxA, p2=2.0)
Dz0 = np.stack([Dk @ z0 for Dk in M.D], axis=1)
o = plan.assemble_host(np.zeros(R.shape[1]), Dz0, np.ones((geom.x.shape[0], len(Dt))), 0.6, 7)
assert o["scal"][1] == 1.0
plan.close()
# sharded plans, 3 virtual ranks, split mode + all-gather of the unknown
geom = mgb_b200.fem2d(3)
pr = problem(geom)
n, m = geom.x.shape[0], pr["R"].shape[1]
N = 3
rp, op = uniform_partition(n, N, geom.block) - 1, uniform_partition(m, N) - 1
plans = [capi.DistPlan(ctx, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0, r, N, rp, op) for r in range(N)]
wins = [pl.window()[0] for pl in plans]
for pl in plans:
    pl.attach_local(wins)
Dz0 = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)
cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
s_d = torch.from_numpy(pr["s"]).to(dev)
ins = [(cm(Dz0[pl.rows]), cm(pr["c"][pl.rows])) for pl in plans]
for step in range(3):
    for pl in plans:
        pl.s_publish(s_d[pl.dinfo["own0"]: pl.dinfo["own1"]].clone())
    sin = [pl.s_wait() for pl in plans]
    for r, pl in enumerate(plans):
        pl.begin(sin[r], ins[r][0], ins[r][1], 0.8, 7)
    for pl in plans:
        hp, gp, sp_ = pl.end(0.8, 7)
        assert ctx.to_host(sp_, 4)[1] == 1.0
for pl in plans:
    pl.close()
# reductions / spmat / isfinite
v = torch.from_numpy(np.random.default_rng(0).normal(size=5000)).to(dev)
ctx.reduce("dot", v, 5000, v); ctx.reduce("maxabs", v, 5000); assert ctx.all_isfinite(v, 5000)
A = capi.SpMat(ctx, pr["R"])
y = torch.zeros(pr["R"].shape[0], dtype=torch.float64, device=dev)
A.mv(s_d, y); A.mv(y, s_d.clone(), trans=True)
ctx.sync()
print("SANITIZE_CASE_OK launches", capi.launch_count())
