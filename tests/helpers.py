"""Shared test helpers: seeded feasible iterates and oracle/CUDA comparison."""
import numpy as np
import scipy.sparse as sp

import mgb_b200
from mgb_b200 import amg as amg_mod


def problem(geom, p=1.0, slack=False, seed=20261018, level=None, pert=1e-3):
    """Seeded strictly feasible iterate for the default p-Laplace problem on ``geom``."""
    dim = geom.dim
    M_main, M_feas = amg_mod.amg(geom)
    M = M_feas if slack else M_main
    n = geom.x.shape[0]
    rng = np.random.default_rng(seed)
    g = amg_mod.DEFAULT_G[dim]
    f = amg_mod.DEFAULT_F[dim]
    z0 = np.array([g(geom.x[i]) for i in range(n)], dtype=float)
    c = np.array([f(geom.x[i]) for i in range(n)], dtype=float)
    if slack:
        z0 = np.hstack([z0, np.full((n, 1), 0.5)])
        c = np.hstack([c, np.ones((n, 1))])
    z0 = z0.reshape(-1, order="F")
    J = (len(M.R_fine) - 1) if level is None else level
    R = M.R_fine[J]
    s = pert * rng.uniform(-1.0, 1.0, size=R.shape[1])
    idx = list(range(1, dim + 2))
    return dict(M=M, R=R, D=M.D, z0=z0, s=s, c=c, idx=idx, p=float(p), slack=slack, x=geom.x, w=geom.w, level=J)


def oracle_eval(pr, t):
    import mgb_oracle as O
    Q = O.EuclidianPower(idx=pr["idx"], p=pr["p"], slack=pr["slack"])
    ct = t * pr["c"]
    args = (pr["s"], pr["x"], pr["w"], ct, pr["R"], pr["D"], pr["z0"], Q)
    return O.f0(*args), O.f1(*args), O.f2(*args).tocsr()


def cuda_eval(ctx, pr, t, force_path=0, host=True):
    from mgb_b200 import capi
    plan = capi.Plan(ctx, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], pr["p"], slack=pr["slack"],
                     force_path=force_path)
    Dz0 = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)
    out = plan.assemble_host(pr["s"], Dz0, pr["c"], t, capi.WANT_F0 | capi.WANT_GRAD | capi.WANT_HESS | capi.STORE_DZ)
    rp, ci = plan.pattern()
    H = sp.csr_matrix((out["hval"], ci.astype(np.int64), rp.astype(np.int64)), shape=(plan.m, plan.m))
    return plan, out, H


def rel(a, b):
    na = np.linalg.norm(np.asarray(b).ravel())
    return np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel()) / (na if na > 0 else 1.0)


def check_against_oracle(ctx, geom, p, t, slack=False, level=None, force_path=0, tol=1e-12, pert=1e-3):
    """north_star tolerance: gradient and Hessian within 1e-12 relative; pattern: oracle's
    (cancellation-dependent) pattern is contained in the plan's structural pattern and every extra
    entry is numerically zero."""
    pr = problem(geom, p=p, slack=slack, level=level, pert=pert)
    f0_o, g_o, H_o = oracle_eval(pr, t)
    plan, out, H_c = cuda_eval(ctx, pr, t, force_path=force_path)
    assert out["scal"][1] == 1.0, "iterate reported non-finite"
    assert abs(out["scal"][0] - f0_o) <= tol * max(1.0, abs(f0_o)), (out["scal"][0], f0_o)
    assert rel(out["grad"], g_o) <= tol, rel(out["grad"], g_o)
    diff = (H_c - H_o).tocsr()
    hn = abs(H_o).max()
    assert abs(diff).max() <= tol * hn, (abs(diff).max(), hn)
    assert sp.linalg.norm(diff) <= tol * sp.linalg.norm(H_o)
    # pattern containment
    Ho = H_o.copy(); Ho.eliminate_zeros()
    Pc = sp.csr_matrix((np.ones(H_c.nnz), H_c.indices, H_c.indptr), shape=H_c.shape)
    Po = sp.csr_matrix((np.ones(Ho.nnz), Ho.indices, Ho.indptr), shape=Ho.shape)
    assert (Po - Po.multiply(Pc)).nnz == 0, "oracle pattern not contained in plan pattern"
    # Dz seam
    import mgb_oracle as O
    Dz_o = O.apply_D(pr["D"], pr["z0"] + pr["R"] @ pr["s"])
    assert rel(out["Dz"], Dz_o) <= 1e-14
    return plan, out
