"""CPU tests of the host logic of the fused peer-memory exchange (mgb_dist_*): partitions,
destination maps, staging layout, owner-side sums.  The local per-rank values come from the oracle
restricted to the rank's quadrature rows; the stores the push kernel would do over NVLink are replayed
with numpy (in-process, any number of ranks) or travel through a gloo all_to_all (two processes).  The
CUDA kernels themselves are covered by tests/test_dist_peer_gpu.py."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _partitions(n, m, block, nranks):
    from mgb_b200.hpc import uniform_partition
    return uniform_partition(n, nranks, block) - 1, uniform_partition(m, nranks) - 1


def _local_values(pr, plan, row0, row1, t):
    """oracle on this rank's rows, sampled on the rank's local pattern"""
    import mgb_oracle as O
    Q = O.EuclidianPower(idx=pr["idx"], p=1.0)
    Dl = [d[row0:row1] for d in pr["D"]]
    args = (pr["s"], pr["x"][row0:row1], pr["w"][row0:row1], t * pr["c"][row0:row1], pr["R"], Dl, pr["z0"], Q)
    Hl = O.f2(*args).tocsr()
    lrp, lci = plan.pattern()
    rows = np.repeat(np.arange(plan.m), np.diff(lrp))
    hv = np.asarray(Hl[rows, lci]).ravel()
    # <c,Dz>_w partial: f0 = sum w F + t <c,Dz>_w ; keep the split the kernels use
    Dz = O.apply_D(Dl, pr["z0"] + pr["R"] @ pr["s"])
    cdot = float(np.sum(pr["w"][row0:row1, None] * pr["c"][row0:row1] * Dz))
    f0 = O.f0(*args)
    return hv, O.f1(*args), np.array([f0 - t * cdot, cdot, 0.0, 0.0])


def _pushes(plan, hv, gv, sv):
    """(dest rank, window offset, value) of everything this rank stores"""
    from mgb_b200 import capi
    mp = plan.maps()
    d = [mp["h_dest"].astype(np.int64)]
    v = [hv]
    touched = np.flatnonzero(mp["g_dest"] >= 0)
    d.append(mp["g_dest"][touched].astype(np.int64))
    v.append(gv[touched])
    d, v = np.concatenate(d), np.concatenate(v)
    rk, off = d >> capi.DIST_RANK_SHIFT, d & capi.DIST_OFF_MASK
    for p in range(plan.nranks):  # staged scalars
        lay = plan.layout(p)
        rk = np.concatenate([rk, np.full(4, p)])
        off = np.concatenate([off, lay["off_stg_scal"] + 4 * plan.rank + np.arange(4)])
        v = np.concatenate([v, sv])
    return rk, off, v


def _finish(plan, win, t):
    """numpy twin of finish_kernel"""
    mp, lay = plan.maps(), plan.layout(plan.rank)
    for pos, ptr, off, stg in ((mp["fh_pos"], mp["fh_ptr"], lay["off_h"], lay["off_stg_h"]),
                               (mp["fg_pos"], mp["fg_ptr"], lay["off_g"], lay["off_stg_g"])):
        for j in range(pos.size):
            win[off + pos[j]] = sum(win[stg + r] for r in range(ptr[j], ptr[j + 1]))   # source-rank order
    S = win[lay["off_stg_scal"]: lay["off_stg_scal"] + 4 * plan.nranks].reshape(plan.nranks, 4).sum(axis=0)
    f0 = S[0] + t * S[1]
    return (win[lay["off_h"]: lay["off_h"] + lay["n_own_h"]], win[lay["off_g"]: lay["off_g"] + lay["n_own_g"]], f0)


def _check_owned(plan, pr, h_own, g_own, f0, t):
    import mgb_oracle as O
    Q = O.EuclidianPower(idx=pr["idx"], p=1.0)
    argsg = (pr["s"], pr["x"], pr["w"], t * pr["c"], pr["R"], pr["D"], pr["z0"], Q)
    Hg, gg, f0g = O.f2(*argsg).tocsr(), O.f1(*argsg), O.f0(*argsg)
    d = plan.dinfo
    lo, hi = d["own0"], d["own1"]
    orp, oci = plan.own_pattern()
    Hown = sp.csr_matrix((h_own, oci.astype(np.int64), orp.astype(np.int64)), shape=(hi - lo, plan.m))
    assert abs(Hown - Hg[lo:hi]).max() <= 1e-12 * abs(Hg).max()
    assert np.abs(g_own - gg[lo:hi]).max() <= 1e-12 * np.abs(gg).max()
    assert abs(f0 - f0g) <= 1e-12 * abs(f0g)


@pytest.mark.parametrize("gen,L,nranks", [("fem2d", 3, 2), ("fem2d", 3, 3), ("fem2d", 2, 4), ("fem1d", 4, 2), ("fem1d", 5, 5)])
def test_peer_maps_reproduce_global_assembly(gen, L, nranks):
    import mgb_b200
    from mgb_b200 import capi
    from helpers import problem
    geom = getattr(mgb_b200, gen)(L)
    pr = problem(geom)
    n, m = geom.x.shape[0], pr["R"].shape[1]
    row_part, out_part = _partitions(n, m, geom.block, nranks)
    t = 0.8
    plans = [capi.DistPlan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0, r, nranks, row_part, out_part)
             for r in range(nranks)]
    wins = [np.full(pl.layout(r)["size"], np.nan) for r, pl in enumerate(plans)]
    nstaged = 0
    for r, pl in enumerate(plans):
        assert pl.dinfo["row0"] == row_part[r] and pl.dinfo["own1"] == out_part[r + 1]
        hv, gv, sv = _local_values(pr, pl, int(row_part[r]), int(row_part[r + 1]), t)
        rk, off, v = _pushes(pl, hv, gv, sv)
        for p in range(nranks):
            sel = rk == p
            assert np.unique(off[sel]).size == sel.sum(), "two stores of one rank hit the same window slot"
            assert np.all(np.isnan(wins[p][off[sel]])), "stores of different ranks collide"
            wins[p][off[sel]] = v[sel]
        nstaged += pl.dinfo["n_fh"]
    assert nstaged > 0, "no interface entries: the test would not exercise the staging path"
    for r, pl in enumerate(plans):
        h_own, g_own, f0 = _finish(pl, wins[r], t)
        assert not np.isnan(h_own).any() and not np.isnan(g_own).any(), "an owned entry was never written"
        _check_owned(pl, pr, h_own, g_own, f0, t)
    # owned blocks tile the global pattern exactly
    gplan = capi.Plan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0)
    grp, gci = gplan.pattern()
    assert sum(pl.dinfo["n_own_h"] for pl in plans) == gplan.nnzH
    assert np.array_equal(np.concatenate([pl.own_pattern()[1] for pl in plans]), gci)


def test_peer_plan_rejects_bad_partitions():
    import mgb_b200
    from mgb_b200 import capi
    from helpers import problem
    geom = mgb_b200.fem2d(2)
    pr = problem(geom)
    n, m = geom.x.shape[0], pr["R"].shape[1]
    rp, op = _partitions(n, m, geom.block, 2)
    bad = rp.copy(); bad[1] += 1   # splits a 7-node element
    with pytest.raises(capi.MgbError):
        capi.DistPlan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0, 0, 2, bad, op)
    with pytest.raises(capi.MgbError):
        capi.DistPlan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0, 2, 2, rp, op)
    # no GPU context: every numeric call fails loudly
    pl = capi.DistPlan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0, 0, 2, rp, op)
    with pytest.raises(capi.MgbError):
        pl.window()
    with pytest.raises(capi.MgbError):
        pl.begin(0, 0, 0, 1.0, 7)


def _gloo_worker(rank, world, port, q):
    try:
        for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
            sys.path.insert(0, p)
        import torch
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import mgb_b200
        from mgb_b200 import capi
        from helpers import problem
        geom = mgb_b200.fem2d(3)
        pr = problem(geom)
        n, m = geom.x.shape[0], pr["R"].shape[1]
        row_part, out_part = _partitions(n, m, geom.block, world)
        t = 0.8
        pl = capi.DistPlan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0, rank, world, row_part, out_part)
        hv, gv, sv = _local_values(pr, pl, int(row_part[rank]), int(row_part[rank + 1]), t)
        rk, off, v = _pushes(pl, hv, gv, sv)
        # the NVLink stores, as messages: (offset, value) pairs grouped by destination rank
        order = np.argsort(rk, kind="stable")
        counts = [int((rk == p).sum()) for p in range(world)]
        send = torch.from_numpy(np.stack([off[order].astype(np.float64), v[order]], axis=1).copy())
        rc = torch.zeros(world, dtype=torch.int64)
        dist.all_to_all_single(rc, torch.tensor(counts, dtype=torch.int64))
        recv = torch.empty((int(rc.sum()), 2), dtype=torch.float64)
        dist.all_to_all_single(recv, send, output_split_sizes=[int(c) for c in rc], input_split_sizes=counts)
        win = np.full(pl.layout(rank)["size"], np.nan)
        roff = recv[:, 0].numpy().astype(np.int64)
        assert np.unique(roff).size == roff.size
        win[roff] = recv[:, 1].numpy()
        h_own, g_own, f0 = _finish(pl, win, t)
        assert not np.isnan(h_own).any() and not np.isnan(g_own).any()
        _check_owned(pl, pr, h_own, g_own, f0, t)
        q.put((rank, "ok", pl.dinfo["n_fh"]))
        dist.destroy_process_group()
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, "error", traceback.format_exc()))


def test_peer_maps_two_processes_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1500)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[1] == "ok", r[2]
