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


# ---------------------------------------------------------------------------------------------------------------
# Extended-precision (80-bit long double) evaluation of the same assembly, for meshes where the float64 result is
# conditioning-limited: on fem1d L=16 the derivative operator has entries of 2^16, so the reference's own association
# D*(z0 + R*s) (test/test_apply_d.jl:44) carries 4e-12 absolute rounding into Dz and 4e-9 relative into the gradient -
# two correct float64 implementations cannot agree to 1e-12 there.  One cone (q, s) = Dz[:, idx], no slack.
# ---------------------------------------------------------------------------------------------------------------
LD = np.longdouble


def _rows_padded(A):
    """CSR -> (cols [n, r], vals [n, r] long double, mask) with r = longest row"""
    A = sp.csr_matrix(A)
    n = A.shape[0]
    cnt = np.diff(A.indptr)
    r = int(cnt.max()) if n else 0
    cols = np.zeros((n, r), dtype=np.int64)
    vals = np.zeros((n, r), dtype=LD)
    pos = np.arange(A.nnz) - np.repeat(A.indptr[:-1], cnt)
    row = np.repeat(np.arange(n), cnt)
    cols[row, pos] = A.indices
    vals[row, pos] = A.data.astype(LD)
    return cols, vals


def _spmv_ld(A, x):
    A = sp.coo_matrix(A)
    out = np.zeros(A.shape[0], dtype=LD)
    np.add.at(out, A.row, A.data.astype(LD) * x[A.col])
    return out


def oracle_eval_longdouble(pr, t):
    """f0, gradient, Hessian (scipy CSR of float64-rounded long double sums) in long double arithmetic, following
    the reference's formulas literally: Dz = D*(z0 + R*s), g = R'*sum_k D_k'*(w.*(F1_k + t c_k)),
    H = R'*(sum_jk D_j' diag(w.*F2_jk) D_k)*R (test/test_map_rows_compare.jl:102-123,165-171)."""
    assert not pr["slack"]
    idx, p = pr["idx"], LD(pr["p"])
    D, R = pr["D"], sp.csr_matrix(pr["R"])
    nD, m = len(D), R.shape[1]
    w = pr["w"].astype(LD)
    z = pr["z0"].astype(LD) + _spmv_ld(R, pr["s"].astype(LD))
    Dz = np.stack([_spmv_ld(Dk, z) for Dk in D], axis=1)
    q, s = Dz[:, idx[:-1]], Dz[:, idx[-1]]
    a = LD(2) / p
    mu = LD(0) if pr["p"] in (1.0, 2.0) else (LD(1) if pr["p"] < 2 else LD(2))
    phi = s ** a - (q * q).sum(axis=1)
    assert (phi > 0).all() and (s > 0).all()
    c = LD(t) * pr["c"].astype(LD)
    f0 = (w * (-np.log(phi) - mu * np.log(s))).sum() + sum((w * c[:, k] * Dz[:, k]).sum() for k in range(nD))
    ds = a * s ** (a - 1)
    y1 = np.zeros_like(Dz)
    y2 = np.zeros((Dz.shape[0], nD, nD), dtype=LD)
    for mq, k in enumerate(idx[:-1]):
        y1[:, k] = 2 * q[:, mq] / phi
        for mq2, k2 in enumerate(idx[:-1]):
            y2[:, k, k2] = 4 * q[:, mq] * q[:, mq2] / (phi * phi) + (2 / phi if mq == mq2 else 0)
        y2[:, k, idx[-1]] = y2[:, idx[-1], k] = -2 * q[:, mq] * ds / (phi * phi)
    y1[:, idx[-1]] = -ds / phi - mu / s
    y2[:, idx[-1], idx[-1]] = -a * (a - 1) * s ** (a - 2) / phi + ds * ds / (phi * phi) + mu / (s * s)
    E = [_rows_padded(sp.csr_matrix(Dk) @ R) for Dk in D]   # structure only matters; values recomputed below
    # E_k = D_k R in long double: rows of D_k R through the padded rows of D_k and R
    Rc, Rv = _rows_padded(R)
    Eld = []
    for Dk in D:
        dc, dv = _rows_padded(Dk)
        cols = Rc[dc]                                   # [n, rD, rR]
        vals = dv[:, :, None] * Rv[dc]
        Eld.append((cols.reshape(cols.shape[0], -1), vals.reshape(vals.shape[0], -1)))
    g = np.zeros(m, dtype=LD)
    for k in range(nD):
        cols, vals = Eld[k]
        np.add.at(g, cols.ravel(), (vals * (w * (y1[:, k] + c[:, k]))[:, None]).ravel())
    keys, prods = [], []
    for j in range(nD):
        for k in range(nD):
            if not np.any(y2[:, j, k] != 0):
                continue
            cj, vj = Eld[j]
            ck, vk = Eld[k]
            pv = (vj * (w * y2[:, j, k])[:, None])[:, :, None] * vk[:, None, :]
            kk = cj[:, :, None] * m + ck[:, None, :]
            nz = pv != 0
            keys.append(kk[nz]); prods.append(pv[nz])
    keys = np.concatenate(keys); prods = np.concatenate(prods)
    order = np.argsort(keys, kind="stable")
    keys, prods = keys[order], prods[order]
    uk, first = np.unique(keys, return_index=True)
    vals = np.add.reduceat(prods, first)
    H = sp.csr_matrix((vals.astype(np.float64), (uk // m, uk % m)), shape=(m, m))
    return float(f0), g, H
