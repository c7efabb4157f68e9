"""world_size-2 gloo test of the multi-GPU host logic (partition, interface exchange maps, owner-side
summation).  Local per-rank values come from the oracle restricted to the rank's quadrature rows, so
no GPU is needed; the CUDA kernels are covered by the gpu tests."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, gen, L, level, q):
    try:
        for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
            sys.path.insert(0, p)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import mgb_b200
        from mgb_b200 import capi
        from mgb_b200 import dist as mdist
        import mgb_oracle as O
        from helpers import problem
        geom = getattr(mgb_b200, gen)(L)
        pr = problem(geom, level=level)
        n, m = geom.x.shape[0], pr["R"].shape[1]
        row0, row1 = mdist.element_rows(n, geom.block, rank, world)
        gplan = capi.Plan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0)
        lplan = capi.Plan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0, rows=(row0, row1))
        grp, gci = gplan.pattern()
        lrp, lci = lplan.pattern()
        ex = mdist.build_exchange(rank, world, m, grp.astype(np.int64), gci.astype(np.int64), lrp.astype(np.int64),
                                  lci.astype(np.int64), torch.device("cpu"))
        # local contributions from the oracle restricted to this rank's quadrature rows
        Q = O.EuclidianPower(idx=pr["idx"], p=1.0)
        t = 0.8
        Dl = [d[row0:row1] for d in pr["D"]]
        args = (pr["s"], pr["x"][row0:row1], pr["w"][row0:row1], t * pr["c"][row0:row1], pr["R"], Dl, pr["z0"], Q)
        Hl = O.f2(*args).tocsr()
        gl = O.f1(*args)
        f0l = O.f0(*args)
        rows = np.repeat(np.arange(m), np.diff(lrp))
        exch = mdist.Exchanger(ex, torch.device("cpu"), n_loc_h=lplan.nnzH, m=m)
        hv, gv, sv = exch.views()
        hv.copy_(torch.from_numpy(np.asarray(Hl[rows, lci]).ravel().copy()))
        gv.copy_(torch.from_numpy(gl.copy()))
        sv.copy_(torch.tensor([f0l, 1.0, 0.0, 0.0], dtype=torch.float64))
        h_own, g_own, scal = exch.exchange()
        # global oracle
        argsg = (pr["s"], pr["x"], pr["w"], t * pr["c"], pr["R"], pr["D"], pr["z0"], Q)
        Hg = O.f2(*argsg).tocsr()
        gg = O.f1(*argsg)
        lo, hi = int(ex.m_part[rank] - 1), int(ex.m_part[rank + 1] - 1)
        Hown = sp.csr_matrix((h_own.numpy(), ex.own_colidx, ex.own_rowptr), shape=(hi - lo, m))
        errH = abs(Hown - Hg[lo:hi]).max() / abs(Hg).max()
        errg = np.abs(g_own.numpy() - gg[lo:hi]).max() / np.abs(gg).max()
        errf = abs(float(scal[0]) - O.f0(*argsg)) / abs(O.f0(*argsg))
        cover = int(sum(ex.h_recv_splits))
        assert float(scal[1]) == 1.0
        q.put((rank, float(errH), float(errg), float(errf), cover, (row0, row1), (lo, hi)))
        dist.destroy_process_group()
    except Exception as exc:  # pragma: no cover
        import traceback
        q.put((rank, "error", traceback.format_exc()))


@pytest.mark.parametrize("gen,L,level", [("fem1d", 4, None), ("fem2d", 3, None), ("fem2d", 3, 1)])
def test_two_rank_exchange_matches_global_assembly(gen, L, level):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + {"fem1d": 0, "fem2d": 1}[gen] + (7 if level else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, gen, L, level, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[1] != "error", r[2]
        rank, errH, errg, errf, cover, rows, own = r
        assert errH < 1e-12 and errg < 1e-12 and errf < 1e-12, r
        assert cover > 0
    rows = sorted(r[5] for r in res)
    assert rows[0][1] == rows[1][0] and rows[0][0] == 0      # contiguous row blocks
    owns = sorted(r[6] for r in res)
    assert owns[0][1] == owns[1][0] and owns[0][0] == 0


def test_partition_rule():
    from mgb_b200.hpc import uniform_partition
    assert np.array_equal(uniform_partition(10, 1), [1, 11])            # [1, n+1] for one rank (tools/profile_solve.jl:24)
    assert np.array_equal(uniform_partition(10, 3), [1, 5, 8, 11])      # first n mod P ranks get the extra row
    assert np.array_equal(uniform_partition(28, 2, block=7), [1, 15, 29])
    assert np.array_equal(uniform_partition(21, 2, block=7), [1, 15, 22])
    p = uniform_partition(229376, 8, block=7)
    assert p[0] == 1 and p[-1] == 229377 and np.all(np.diff(p) % 7 == 0)
