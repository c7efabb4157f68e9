"""two-cone (parabolic) assembly time on fem2d L=7 (config C5 mesh): element path vs CSR path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"):
    sys.path.insert(0, p)
import numpy as np, torch
import mgb_b200
from mgb_b200 import capi
import mgb_oracle as O
L = int(sys.argv[1]) if len(sys.argv) > 1 else 7
ctx = capi.Context(0); dev = torch.device("cuda", 0)
geom = mgb_b200.fem2d(L)
Dt, idxA, idxB = O.parabolic_tables(2)
M = O.amg_helper(geom, O.PARABOLIC_STATE, Dt)
n = geom.x.shape[0]
u = np.sin(geom.x[:, 0]) + geom.x[:, 1] ** 2
z0 = O.parabolic_feasible_start(M, u, 2, 1.0)
R = M.R_fine[-1]
rng = np.random.default_rng(5)
s = 1e-3 * rng.uniform(-1, 1, size=R.shape[1]); c = rng.normal(size=(n, len(Dt)))
Dz0 = np.stack([Dk @ z0 for Dk in M.D], axis=1)
cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
for path in (capi.PATH_ELEMENT, capi.PATH_CSR):
    plan = capi.Plan(ctx, M.D, R, geom.x, geom.w, idxB, 1.0, idx2=idxA, p2=2.0, force_path=path)
    s_d = torch.from_numpy(s).to(dev); Dz0_d = cm(Dz0); c_d = cm(c)
    scal = torch.zeros(4, dtype=torch.float64, device=dev); grad = torch.zeros(plan.m, dtype=torch.float64, device=dev)
    hval = torch.zeros(plan.nnzH, dtype=torch.float64, device=dev)
    plan.time_assemble(s_d, Dz0_d, c_d, 1.0, 7, scal, grad, hval, 3, 2, split=False)
    ms, _, _ = plan.time_assemble(s_d, Dz0_d, c_d, 1.0, 7, scal, grad, hval, 10, 2, split=False)
    print("parabolic fem2d L=%d n=%d m=%d nnzH=%d path=%s ms=%.4f alg_bytes=%d hess_contribs=%d" % (L, n, plan.m, plan.nnzH, "element" if path == 1 else "csr", ms, plan.info["alg_bytes"], plan.info["hess_contribs"]), flush=True)
    plan.close()
