"""torchrun worker: sharded (owner-computes) assembly, fused call - the gather kernel of every rank publishes its
objective partials into the peers' windows (CUDA IPC over NVLink) and waits for theirs - checked against the oracle
on every rank.  Launched by tests/test_dist_peer_gpu.py with >= 2 GPUs."""
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
ctx = capi.Context(lr, torch.cuda.current_stream(dev).cuda_stream)
for gen, L, colocate in (("fem2d", 4, True), ("fem2d", 6, True), ("fem2d", 5, False), ("fem1d", 7, True)):
    geom = getattr(mgb_b200, gen)(L)
    pr = problem(geom)
    n, m = geom.x.shape[0], pr["R"].shape[1]
    plan = mdist.create_peer_plan(ctx, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0, geom.block, rank, world,
                                  colocate=colocate)
    # colocate: the unknowns are renumbered rank-major (plan.perm, new -> old); the oracle keeps the reference's
    # numbering and its outputs are renumbered for the comparison
    perm = plan.perm if plan.perm is not None else np.arange(m)
    assert (plan.perm is not None) == colocate
    d = plan.dinfo
    Dz0 = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)[plan.rows]
    cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
    Dz0_d, c_d = cm(Dz0), cm(pr["c"][plan.rows])
    Q = O.EuclidianPower(idx=pr["idx"], p=1.0)
    rng = np.random.default_rng(11)
    for step in range(4):
        t = 0.8 + 0.05 * step
        if step:
            pr["s"] = pr["s"] + 1e-4 * rng.uniform(-1, 1, size=pr["s"].shape)   # same seed on every rank
        s_d = torch.from_numpy(pr["s"][perm]).to(dev)
        if step % 2 == 1:   # row-distributed unknown: all-gather over NVLink peer memory inside the library
            s_own = s_d[d["own0"]: d["own1"]].clone()
            hp, gp, sp_ = plan.dist_assemble_s(s_own, Dz0_d, c_d, t, 7)
        else:
            hp, gp, sp_ = plan.dist_assemble(s_d, Dz0_d, c_d, t, 7)
        h_own, g_own, scal = ctx.to_host(hp, d["n_own_h"]), ctx.to_host(gp, d["n_own_g"]), ctx.to_host(sp_, 4)
        argsg = (pr["s"], pr["x"], pr["w"], t * pr["c"], pr["R"], pr["D"], pr["z0"], Q)
        Hg, gg, f0g = O.f2(*argsg).tocsr()[perm][:, perm].tocsr(), O.f1(*argsg)[perm], O.f0(*argsg)
        lo, hi = d["own0"], d["own1"]
        orp, oci = plan.own_pattern()
        Hown = sp.csr_matrix((h_own, oci.astype(np.int64), orp.astype(np.int64)), shape=(hi - lo, m))
        errH = abs(Hown - Hg[lo:hi]).max() / abs(Hg).max()
        errg = np.abs(g_own - gg[lo:hi]).max() / np.abs(gg).max()
        errf = abs(scal[0] - f0g) / abs(f0g)
        assert errH < 1e-12 and errg < 1e-12 and errf < 1e-12, (gen, L, step, errH, errg, errf)
        assert scal[1] == 1.0 and plan.dist_info()["err"] == 0
    # back-to-back epochs without host synchronisation (two window parities, epoch tags)
    for _ in range(50):
        hp, gp, sp_ = plan.dist_assemble(s_d, Dz0_d, c_d, t, 7)
    h2 = ctx.to_host(hp, d["n_own_h"])
    assert np.array_equal(h2, h_own), "epoch pipelining changed the result"
    mdist.destroy_peer_plan(plan)
dist.barrier()
if rank == 0:
    print("PEER_OK")
dist.destroy_process_group()
