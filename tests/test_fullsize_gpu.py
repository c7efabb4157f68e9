"""Full-size checks at BASELINE.json's sizes (fem2d L=8: n=229,376; fem1d L=16: n=131,072; fem3d L=5: n=262,144), where the oracle is
too slow to run inside the suite: size-independent properties of the assembled objects.

  * R'HR is symmetric on its (symmetric) frozen pattern;
  * the gradient is the derivative of the objective        (central differences along a random direction);
  * the Hessian is the derivative of the gradient          (H v vs central difference of gradients);
  * a sharded assembly (4 virtual ranks, owner-computes) reproduces the single-plan values;
  * non-finite / infeasible iterates are reported as data (all_finite = 0), never as an error.
"""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def _setup(gpu_ctx, gen, L, p=1.0):
    import mgb_b200
    from mgb_b200 import capi
    from helpers import problem
    geom = getattr(mgb_b200, gen)(L)
    pr = problem(geom, p=p, pert=1e-8 if gen == "fem1d" else 1e-3)
    plan = capi.Plan(gpu_ctx, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], p)
    Dz0 = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)
    return geom, pr, plan, Dz0


@pytest.mark.parametrize("gen,L", [("fem2d", 8), ("fem1d", 16), ("fem3d", 5)])
def test_derivative_identities_at_full_size(gpu_ctx, gen, L):
    from mgb_b200 import capi
    geom, pr, plan, Dz0 = _setup(gpu_ctx, gen, L)
    t = 0.7
    s = pr["s"]
    out = plan.assemble_host(s, Dz0, pr["c"], t, 7)
    assert out["scal"][1] == 1.0
    rp, ci = plan.pattern()
    H = sp.csr_matrix((out["hval"], ci.astype(np.int64), rp.astype(np.int64)), shape=(plan.m, plan.m))
    asym = abs(H - H.T).max()
    assert asym <= 1e-13 * abs(H).max(), asym
    # objective -> gradient, along the gradient itself (a random direction in 2e5 dimensions has a directional
    # derivative below the rounding noise of the objective)
    g = out["grad"]
    v = g / np.linalg.norm(g)
    eps = 1e-8 if gen == "fem1d" else 1e-4   # the 1-D feasible set is thin at L=16 (element size 2^-16)
    op = plan.assemble_host(s + eps * v, None, None, t, 1, upload_inputs=False)
    om = plan.assemble_host(s - eps * v, None, None, t, 1, upload_inputs=False)
    assert op["scal"][1] == 1.0 and om["scal"][1] == 1.0
    dfd = (op["scal"][0] - om["scal"][0]) / (2 * eps)
    assert abs(dfd - np.linalg.norm(g)) <= 1e-4 * np.linalg.norm(g), (dfd, np.linalg.norm(g))
    # gradient -> Hessian, along a random direction
    rng = np.random.default_rng(3)
    v = rng.standard_normal(plan.m)
    v /= np.linalg.norm(v)
    eps = 1e-9 if gen == "fem1d" else 1e-6
    op = plan.assemble_host(s + eps * v, None, None, t, 3, upload_inputs=False)
    om = plan.assemble_host(s - eps * v, None, None, t, 3, upload_inputs=False)
    hv_fd = (op["grad"] - om["grad"]) / (2 * eps)
    hv = H @ v
    assert np.linalg.norm(hv_fd - hv) <= 1e-4 * np.linalg.norm(hv), np.linalg.norm(hv_fd - hv) / np.linalg.norm(hv)

def test_sharded_equals_single_at_full_size(gpu_ctx):
    """fem2d L=8 on 4 virtual ranks (split mode, one GPU): owned blocks tile the single-plan result"""
    import torch
    from mgb_b200 import capi
    from mgb_b200 import dist as mdist
    geom, pr, plan, Dz0 = _setup(gpu_ctx, "fem2d", 8)
    t = 0.7
    ref = plan.assemble_host(pr["s"], Dz0, pr["c"], t, 7)
    rp, ci = plan.pattern()
    n, m, N = geom.x.shape[0], plan.m, 4
    row_part, out_part = mdist.peer_partitions(n, m, geom.block, N)
    plans = [capi.DistPlan(gpu_ctx, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0, r, N, row_part, out_part) for r in range(N)]
    wins = [p.window()[0] for p in plans]
    for p in plans:
        p.attach_local(wins)
    dev = torch.device("cuda", gpu_ctx.device)
    cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
    s_d = torch.from_numpy(pr["s"]).to(dev)
    ins = [(cm(Dz0[p.rows]), cm(pr["c"][p.rows])) for p in plans]
    for r, p in enumerate(plans):
        p.begin(s_d, ins[r][0], ins[r][1], t, 7)
    ptrs = [p.end(t, 7) for p in plans]
    hs, gs = [], []
    for r, p in enumerate(plans):
        d = p.dinfo
        hs.append(gpu_ctx.to_host(ptrs[r][0], d["n_own_h"]))
        gs.append(gpu_ctx.to_host(ptrs[r][1], d["n_own_g"]))
        scal = gpu_ctx.to_host(ptrs[r][2], 4)
        assert abs(scal[0] - ref["scal"][0]) <= 1e-13 * abs(ref["scal"][0]) and scal[1] == 1.0
        assert p.dist_info()["err"] == 0
        assert np.array_equal(p.own_pattern()[1], ci[rp[d["own0"]]:rp[d["own1"]]])
    h_all, g_all = np.concatenate(hs), np.concatenate(gs)
    assert h_all.size == plan.nnzH
    # (a rank lists its primary elements before its halo elements, so an entry's two contributions may be added in the
    # other order than in the single plan: agreement to rounding, not bit for bit)
    assert np.abs(h_all - ref["hval"]).max() <= 1e-13 * np.abs(ref["hval"]).max()
    assert np.abs(g_all - ref["grad"]).max() <= 1e-13 * np.abs(ref["grad"]).max()
    for p in plans:
        p.close()


@pytest.mark.parametrize("gen,L", [("fem2d", 3), ("fem3d", 2)])
def test_nonfinite_and_infeasible_iterates_are_data(gpu_ctx, gen, L):
    """amgb_all_isfinite semantics (reference src/MultiGridBarrierMPI.jl:121-133): rc = 0, all_finite = 0
    (element path and CSR path)"""
    from mgb_b200 import capi
    geom, pr, plan, Dz0 = _setup(gpu_ctx, gen, L)
    s = pr["s"].copy()
    s[7] = np.nan
    out = plan.assemble_host(s, Dz0, pr["c"], 1.0, 7)
    assert out["scal"][1] == 0.0
    # push the slack variable far below the cone: s^(2/p) - |q|^2 < 0 at many points
    s = pr["s"].copy()
    s[plan.m // 2:] -= 1e6
    out = plan.assemble_host(s, Dz0, pr["c"], 1.0, 1)
    assert out["scal"][1] == 0.0 and out["scal"][3] > 0 and not np.isfinite(out["scal"][0])
    # and a feasible call on the same plan afterwards is unaffected
    out = plan.assemble_host(pr["s"], Dz0, pr["c"], 1.0, 1)
    assert out["scal"][1] == 1.0 and np.isfinite(out["scal"][0])
