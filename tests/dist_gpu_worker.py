import os
import sys

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import mgb_b200  # noqa: E402
from mgb_b200 import capi, dist as mdist  # noqa: E402
import mgb_oracle as O  # noqa: E402
from helpers import problem  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
for gen, L, level in (("fem2d", 4, None), ("fem2d", 3, 1), ("fem1d", 5, None)):
    geom = getattr(mgb_b200, gen)(L)
    pr = problem(geom, level=level)
    n, m = geom.x.shape[0], pr["R"].shape[1]
    rows = mdist.element_rows(n, geom.block, rank, world)
    ctx = capi.Context(lr)
    plan = capi.Plan(ctx, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0, rows=rows)
    gplan = capi.Plan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0)
    grp, gci = gplan.pattern()
    lrp, lci = plan.pattern()
    ex = mdist.build_exchange(rank, world, m, grp.astype(np.int64), gci.astype(np.int64), lrp.astype(np.int64),
                              lci.astype(np.int64), dev)
    exch = mdist.Exchanger(ex, dev, ctx=ctx, n_loc_h=plan.nnzH, m=m)
    t = 0.8
    Dz0 = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)[rows[0]:rows[1]]
    cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
    s_d = torch.from_numpy(pr["s"]).to(dev)
    hval, grad, scal = exch.views()
    plan.assemble(s_d, cm(Dz0), cm(pr["c"][rows[0]:rows[1]]), t, 7, scal, grad, hval)
    h_own, g_own, scal = exch.exchange()
    Q = O.EuclidianPower(idx=pr["idx"], p=1.0)
    argsg = (pr["s"], pr["x"], pr["w"], t * pr["c"], pr["R"], pr["D"], pr["z0"], Q)
    Hg, gg, f0g = O.f2(*argsg).tocsr(), O.f1(*argsg), O.f0(*argsg)
    lo, hi = int(ex.m_part[rank] - 1), int(ex.m_part[rank + 1] - 1)
    Hown = sp.csr_matrix((h_own.cpu().numpy(), ex.own_colidx, ex.own_rowptr), shape=(hi - lo, m))
    errH = abs(Hown - Hg[lo:hi]).max() / abs(Hg).max()
    errg = np.abs(g_own.cpu().numpy() - gg[lo:hi]).max() / np.abs(gg).max()
    errf = abs(float(scal[0].cpu()) - f0g) / abs(f0g)
    assert errH < 1e-12 and errg < 1e-12 and errf < 1e-12, (gen, L, level, errH, errg, errf)
    assert float(scal[1].cpu()) == 1.0
dist.barrier()
if rank == 0:
    print("DIST_OK")
dist.destroy_process_group()
