"""GPU parity of the sharded (owner-computes) assembly: element_kernel -> gather_kernel on every rank's elements,
objective scalars summed across ranks through peer-memory words.

Single-GPU box: the ranks are *virtual* - N distributed plans on one device and one stream, windows attached by
raw pointer (mgb_dist_attach_local), split mode: every rank's mgb_dist_begin (which publishes its partial sums)
runs before any mgb_dist_end (which collects them), so no kernel ever waits for another one on the same GPU.
With >= 2 GPUs the torchrun workers run the fused call (mgb_dist_assemble: the gather kernel itself waits for the
peers' words) over CUDA IPC + NVLink (tests/dist_peer_worker.py, tests/dist_solve_worker.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _virtual_ranks(ctx, gen, L, nranks, p=1.0, slack=False, steps=3):
    import torch
    import mgb_b200
    from mgb_b200 import capi
    from mgb_b200.hpc import uniform_partition
    from helpers import problem, oracle_eval
    geom = getattr(mgb_b200, gen)(L)
    pr = problem(geom, p=p, slack=slack)
    n, m = geom.x.shape[0], pr["R"].shape[1]
    row_part = uniform_partition(n, nranks, geom.block) - 1
    out_part = uniform_partition(m, nranks) - 1
    plans = [capi.DistPlan(ctx, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], p, r, nranks, row_part, out_part, slack=slack)
             for r in range(nranks)]
    wins = [pl.window()[0] for pl in plans]
    for pl in plans:
        pl.attach_local(wins)
    # every quadrature row is "primary" (counted in the scalars) on exactly one rank
    prim = np.concatenate([pl.rows[: pl.dinfo["n_primary"]] for pl in plans])
    assert np.array_equal(np.sort(prim), np.arange(n))
    dev = torch.device("cuda", ctx.device)
    Dz0 = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)
    cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
    ins = [(cm(Dz0[pl.rows]), cm(pr["c"][pl.rows])) for pl in plans]
    flags = capi.WANT_F0 | capi.WANT_GRAD | capi.WANT_HESS
    rng = np.random.default_rng(5)
    for step in range(steps):   # several epochs: exercises both window parities and the epoch tags
        t = 0.8 + 0.1 * step
        if step:
            pr["s"] = pr["s"] + 1e-4 * rng.uniform(-1, 1, size=pr["s"].shape)
        s_d = torch.from_numpy(pr["s"]).to(dev)
        if step % 2 == 1:
            # row-distributed unknown (the reference's HPCVector): every rank publishes its block, the library
            # all-gathers it into each rank's window; the assembly then reads the gathered copy
            for pl in plans:
                pl.s_publish(s_d[pl.dinfo["own0"]: pl.dinfo["own1"]].clone())
            s_in = [pl.s_wait() for pl in plans]
            for pl, ptr in zip(plans, s_in):
                assert np.array_equal(ctx.to_host(ptr, m), pr["s"])
        else:
            s_in = [s_d] * nranks
        for r, pl in enumerate(plans):
            pl.begin(s_in[r], ins[r][0], ins[r][1], t, flags)
        ptrs = [pl.end(t, flags) for pl in plans]
        f0_o, g_o, H_o = oracle_eval(pr, t)
        scals = []
        for r, pl in enumerate(plans):
            d = pl.dinfo
            hp, gp, sp_ = ptrs[r]
            h_own = ctx.to_host(hp, d["n_own_h"])
            g_own = ctx.to_host(gp, d["n_own_g"])
            scal = ctx.to_host(sp_, 4)
            scals.append(scal)
            orp, oci = pl.own_pattern()
            lo, hi = d["own0"], d["own1"]
            Hown = sp.csr_matrix((h_own, oci.astype(np.int64), orp.astype(np.int64)), shape=(hi - lo, m))
            assert abs(Hown - H_o[lo:hi]).max() <= 1e-12 * abs(H_o).max(), (gen, L, nranks, step, r)
            assert np.abs(g_own - g_o[lo:hi]).max() <= 1e-12 * np.abs(g_o).max()
            assert abs(scal[0] - f0_o) <= 1e-12 * abs(f0_o) and scal[1] == 1.0
            assert pl.dist_info()["err"] == 0
        assert all(np.array_equal(scals[0], sc) for sc in scals), "rank-ordered sums must give identical bits on every rank"
    # f0-only call (line search): the scalars still cross the ranks
    for r, pl in enumerate(plans):
        pl.begin(s_d, ins[r][0], ins[r][1], t, capi.WANT_F0)
    for pl in plans:
        _, _, sp_ = pl.end(t, capi.WANT_F0)
        assert abs(ctx.to_host(sp_, 4)[0] - f0_o) <= 1e-12 * abs(f0_o)
    for pl in plans:
        pl.close()


@pytest.mark.parametrize("gen,L,nranks", [("fem2d", 4, 2), ("fem2d", 4, 3), ("fem2d", 3, 4), ("fem2d", 5, 8),
                                          ("fem1d", 6, 2), ("fem1d", 5, 4)])
def test_virtual_ranks_match_oracle(gpu_ctx, gen, L, nranks):
    _virtual_ranks(gpu_ctx, gen, L, nranks)


def test_virtual_ranks_p_slack_and_single_rank(gpu_ctx):
    _virtual_ranks(gpu_ctx, "fem2d", 3, 2, p=1.5)
    _virtual_ranks(gpu_ctx, "fem2d", 3, 2, slack=True)     # three state variables (feasibility phase)
    _virtual_ranks(gpu_ctx, "fem2d", 3, 1)                 # nranks = 1


def test_sharded_two_cone_plan(gpu_ctx):
    """the parabolic barrier (two cones, three state variables) shards like the others"""
    import torch
    import mgb_b200
    import mgb_oracle as O
    from mgb_b200 import capi
    from mgb_b200.hpc import uniform_partition
    geom = mgb_b200.fem2d(3)
    dim, p, t, N = 2, 1.0, 0.6, 3
    Dt, idxA, idxB = O.parabolic_tables(dim)
    M = O.amg_helper(geom, O.PARABOLIC_STATE, Dt)
    n = geom.x.shape[0]
    rng = np.random.default_rng(5)
    u = np.sin(geom.x[:, 0]) + geom.x[:, 1] ** 2
    z0 = O.parabolic_feasible_start(M, u, dim, p)
    R = M.R_fine[-1]
    m = R.shape[1]
    s = 1e-3 * rng.uniform(-1, 1, size=m)
    c = rng.normal(size=(n, len(Dt)))
    Q = O.Intersection([O.EuclidianPower(idx=idxA, p=2.0), O.EuclidianPower(idx=idxB, p=p)])
    args = (s, geom.x, geom.w, t * c, R, M.D, z0, Q)
    f0, g, H = O.f0(*args), O.f1(*args), O.f2(*args).tocsr()
    rp, op = uniform_partition(n, N, geom.block) - 1, uniform_partition(m, N) - 1
    plans = [capi.DistPlan(gpu_ctx, M.D, R, geom.x, geom.w, idxB, p, r, N, rp, op, idx2=idxA, p2=2.0) for r in range(N)]
    wins = [pl.window()[0] for pl in plans]
    for pl in plans:
        pl.attach_local(wins)
    dev = torch.device("cuda", gpu_ctx.device)
    Dz0 = np.stack([Dk @ z0 for Dk in M.D], axis=1)
    cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
    s_d = torch.from_numpy(s).to(dev)
    ins = [(cm(Dz0[pl.rows]), cm(c[pl.rows])) for pl in plans]
    for r, pl in enumerate(plans):
        pl.begin(s_d, ins[r][0], ins[r][1], t, 7)
    for pl in plans:
        hp, gp, sp_ = pl.end(t, 7)
        d = pl.dinfo
        orp, oci = pl.own_pattern()
        lo, hi = d["own0"], d["own1"]
        Hown = sp.csr_matrix((gpu_ctx.to_host(hp, d["n_own_h"]), oci.astype(np.int64), orp.astype(np.int64)), shape=(hi - lo, m))
        assert abs(Hown - H[lo:hi]).max() <= 1e-12 * abs(H).max()
        assert np.abs(gpu_ctx.to_host(gp, d["n_own_g"]) - g[lo:hi]).max() <= 1e-12 * np.abs(g).max()
        assert abs(gpu_ctx.to_host(sp_, 4)[0] - f0) <= 1e-12 * abs(f0)
    for pl in plans:
        pl.close()


def test_missing_peer_is_reported_not_summed(gpu_ctx):
    """a rank whose peer never publishes its scalars must not return a partial sum as the objective: after the
    time-out the scalars are NaN with all_finite = 0 and the plan's error flag is set (the GPU does not hang)"""
    import torch
    import mgb_b200
    from mgb_b200 import capi
    from mgb_b200.hpc import uniform_partition
    from helpers import problem
    os.environ["MGB_DIST_TIMEOUT_S"] = "0.2"
    try:
        geom = mgb_b200.fem2d(3)
        pr = problem(geom)
        n, m = geom.x.shape[0], pr["R"].shape[1]
        rp, op = uniform_partition(n, 2, geom.block) - 1, uniform_partition(m, 2) - 1
        plans = [capi.DistPlan(gpu_ctx, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0, r, 2, rp, op) for r in range(2)]
    finally:
        del os.environ["MGB_DIST_TIMEOUT_S"]
    wins = [pl.window()[0] for pl in plans]
    for pl in plans:
        pl.attach_local(wins)
    dev = torch.device("cuda", gpu_ctx.device)
    Dz0 = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)
    cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
    s_d = torch.from_numpy(pr["s"]).to(dev)
    pl = plans[0]
    pl.begin(s_d, cm(Dz0[pl.rows]), cm(pr["c"][pl.rows]), 1.0, 7)   # rank 1 never runs
    _, _, sp_ = pl.end(1.0, 7)
    scal = gpu_ctx.to_host(sp_, 4)
    assert pl.dist_info()["err"] == 1
    assert np.isnan(scal[0]) and scal[1] == 0.0
    for pl in plans:
        pl.close()


def test_two_gpu_sharded_assembly_matches_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29641", os.path.join(ROOT, "tests", "dist_peer_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "PEER_OK" in res.stdout


def test_two_gpu_sharded_solve_matches_oracle():
    """configs[0] of BASELINE.json: fem2d_mpi_solve(L=3, p=1.0) on 2 ranks, one GPU each (+ a 1-D, a p=1.5 and a
    fem3d case, which falls back to redundant assembly)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29643", os.path.join(ROOT, "tests", "dist_solve_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "SOLVE_OK" in res.stdout
