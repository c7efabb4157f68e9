"""GPU parity of the fused peer-memory exchange (element_kernel -> push_kernel -> finish_kernel).

Single-GPU box: the ranks are *virtual* - N distributed plans on one device, windows attached by raw
pointer (mgb_dist_attach_local), all pushes launched before any finish (same stream), so the flag
protocol, destination maps, staging sums and epoch double-buffering run exactly as on N GPUs.
With >= 2 GPUs the torchrun worker runs the real thing over CUDA IPC + NVLink (tests/dist_peer_worker.py)."""
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
    dev = torch.device("cuda", ctx.device)
    Dz0 = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)
    cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
    ins = [(cm(Dz0[row_part[r]:row_part[r + 1]]), cm(pr["c"][row_part[r]:row_part[r + 1]])) for r in range(nranks)]
    flags = capi.WANT_F0 | capi.WANT_GRAD | capi.WANT_HESS
    rng = np.random.default_rng(5)
    for step in range(steps):   # several epochs: exercises both window parities and the flag counters
        t = 0.8 + 0.1 * step
        if step:
            pr["s"] = pr["s"] + 1e-4 * rng.uniform(-1, 1, size=pr["s"].shape)
        s_d = torch.from_numpy(pr["s"]).to(dev)
        for r, pl in enumerate(plans):
            pl.begin(s_d, ins[r][0], ins[r][1], t, flags)
        ptrs = [pl.end(t, flags) for pl in plans]
        f0_o, g_o, H_o = oracle_eval(pr, t)
        for r, pl in enumerate(plans):
            d = pl.dinfo
            hp, gp, sp_ = ptrs[r]
            h_own = ctx.to_host(hp, d["n_own_h"])
            g_own = ctx.to_host(gp, d["n_own_g"])
            scal = ctx.to_host(sp_, 4)
            orp, oci = pl.own_pattern()
            lo, hi = d["own0"], d["own1"]
            Hown = sp.csr_matrix((h_own, oci.astype(np.int64), orp.astype(np.int64)), shape=(hi - lo, m))
            assert abs(Hown - H_o[lo:hi]).max() <= 1e-12 * abs(H_o).max(), (gen, L, nranks, step, r)
            assert np.abs(g_own - g_o[lo:hi]).max() <= 1e-12 * np.abs(g_o).max()
            assert abs(scal[0] - f0_o) <= 1e-12 * abs(f0_o) and scal[1] == 1.0
            assert pl.dist_info()["err"] == 0
    # f0-only call (line search): no Hessian/gradient pushes, scalars still cross ranks
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


def test_virtual_ranks_p_and_single_rank(gpu_ctx):
    _virtual_ranks(gpu_ctx, "fem2d", 3, 2, p=1.5)
    _virtual_ranks(gpu_ctx, "fem2d", 3, 1)      # nranks = 1: every entry is single-source


def test_missing_peer_times_out_without_hanging(gpu_ctx):
    """a rank whose peer never publishes its flag reports err=1 after the timeout instead of hanging the GPU"""
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
    plans[0].begin(s_d, cm(Dz0[rp[0]:rp[1]]), cm(pr["c"][rp[0]:rp[1]]), 1.0, 7)   # rank 1 never runs
    plans[0].end(1.0, 7)
    gpu_ctx.sync()
    assert plans[0].dist_info()["err"] == 1
    for pl in plans:
        pl.close()


def test_fused_finish_two_streams(gpu_ctx):
    """mgb_dist_assemble (finish fused into the push kernel's last CTA): two ranks on two streams of one GPU,
    each waiting in-kernel for the other's epoch flag."""
    import torch
    import mgb_b200
    from mgb_b200 import capi
    from mgb_b200.hpc import uniform_partition
    from helpers import problem, oracle_eval
    geom = mgb_b200.fem2d(4)
    pr = problem(geom)
    n, m = geom.x.shape[0], pr["R"].shape[1]
    rp, op = uniform_partition(n, 2, geom.block) - 1, uniform_partition(m, 2) - 1
    dev = torch.device("cuda", gpu_ctx.device)
    streams = [torch.cuda.Stream(dev) for _ in range(2)]
    ctxs = [capi.Context(gpu_ctx.device, st.cuda_stream) for st in streams]
    plans = [capi.DistPlan(ctxs[r], pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0, r, 2, rp, op) for r in range(2)]
    wins = [pl.window()[0] for pl in plans]
    for pl in plans:
        pl.attach_local(wins)
    Dz0 = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)
    cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
    ins = [(cm(Dz0[rp[r]:rp[r + 1]]), cm(pr["c"][rp[r]:rp[r + 1]])) for r in range(2)]
    s_d = torch.from_numpy(pr["s"]).to(dev)
    torch.cuda.synchronize(dev)
    t = 0.9
    f0_o, g_o, H_o = oracle_eval(pr, t)
    for rep in range(6):   # repeated epochs, alternating which rank launches first
        order = (0, 1) if rep % 2 == 0 else (1, 0)
        ptrs = {}
        for r in order:
            ptrs[r] = plans[r].dist_assemble(s_d, ins[r][0], ins[r][1], t, 7)
        for r in range(2):
            d = plans[r].dinfo
            hp, gp, sp_ = ptrs[r]
            h_own, g_own, scal = ctxs[r].to_host(hp, d["n_own_h"]), ctxs[r].to_host(gp, d["n_own_g"]), ctxs[r].to_host(sp_, 4)
            orp, oci = plans[r].own_pattern()
            lo, hi = d["own0"], d["own1"]
            Hown = sp.csr_matrix((h_own, oci.astype(np.int64), orp.astype(np.int64)), shape=(hi - lo, m))
            assert abs(Hown - H_o[lo:hi]).max() <= 1e-12 * abs(H_o).max()
            assert np.abs(g_own - g_o[lo:hi]).max() <= 1e-12 * np.abs(g_o).max()
            assert abs(scal[0] - f0_o) <= 1e-12 * abs(f0_o) and scal[1] == 1.0
            assert plans[r].dist_info()["err"] == 0
    for pl in plans:
        pl.close()
    for c in ctxs:
        c.close()


def test_two_gpu_peer_exchange_matches_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29641", os.path.join(ROOT, "tests", "dist_peer_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "PEER_OK" in res.stdout


def test_two_gpu_sharded_solve_matches_oracle():
    """configs[0] of BASELINE.json: fem2d_mpi_solve(L=3, p=1.0) on 2 ranks, one GPU each (+ a 1-D and a p=1.5 case)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29643", os.path.join(ROOT, "tests", "dist_solve_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "SOLVE_OK" in res.stdout
