"""Pins the CPU oracle against every known-answer vector / structural size the reference holds for
the hot path, and checks its derivatives by finite differences."""
import numpy as np
import pytest
import scipy.sparse as sp

import mgb_b200
import mgb_oracle as O


def test_map_rows_kats():
    # reference test/test_helpers.jl:123-167
    x = np.array([[1.0, 2], [3, 4], [5, 6]])
    assert np.array_equal(O.map_rows(lambda r: r.sum(), x), [3.0, 7.0, 11.0])
    x = np.array([[1.0, 2], [3, 4]])
    assert np.array_equal(O.map_rows(lambda r: np.array([r.sum(), r.prod()]), x), [[3.0, 2.0], [7.0, 12.0]])
    y = np.array([10.0, 20.0])
    assert np.array_equal(O.map_rows(lambda rx, ry: rx.sum() + ry[0], x, y), [13.0, 27.0])


def test_amgb_diag_kat():
    # reference test/test_diag.jl:28-46: diag of 1:10
    d = O.amgb_diag(np.arange(1.0, 11.0))
    assert d.shape == (10, 10) and d.nnz == 10
    assert np.array_equal(d.diagonal(), np.arange(1.0, 11.0))


def test_structural_sizes():
    # fem2d n = 14*4^(L-1): reference docs/src/guide.md:246-253
    for L, n in [(1, 14), (2, 56), (3, 224), (4, 896), (5, 3584)]:
        assert mgb_b200.fem2d(L).x.shape == (n, 2)
    # fem1d n = 2^(L+1); finest Dirichlet subspace 16x7 at L=3: reference test/test_nonsquare.jl:28
    g = mgb_b200.fem1d(3)
    assert g.x.shape == (16, 1) and g.subspaces["dirichlet"][-1].shape == (16, 7)
    assert mgb_b200.fem1d(2).x.shape == (8, 1)  # test/test_partition_debug.jl:34
    # D0 is n x (nu*n): test/test_partition_debug.jl:34
    M = O.amg_helper(mgb_b200.fem1d(2), (("u", "dirichlet"), ("s", "full")), O.DEFAULT_D[1])
    assert all(d.shape == (8, 16) for d in M.D)


@pytest.mark.parametrize("p", [1.0, 1.5, 2.0, 3.0])
@pytest.mark.parametrize("slack", [False, True])
def test_barrier_derivatives_fd(p, slack):
    rng = np.random.default_rng(1)
    nD = 5 if slack else 4
    Q = O.EuclidianPower(idx=[1, 2, 3], p=p, slack=slack)
    y = rng.normal(size=(6, nD)) * 0.3
    y[:, 3] = 3.0 + rng.uniform(size=6)
    if slack:
        y[:, 4] = 0.5
    g = Q.F1(None, y)
    H = Q.F2(None, y)
    h = 1e-6
    for k in range(nD):
        e = np.zeros(nD); e[k] = h
        gfd = (Q.F(None, y + e) - Q.F(None, y - e)) / (2 * h)
        assert np.allclose(g[:, k], gfd, rtol=1e-6, atol=1e-8)
        Hfd = (Q.F1(None, y + e) - Q.F1(None, y - e)) / (2 * h)
        assert np.allclose(H[:, :, k], Hfd, rtol=1e-5, atol=1e-7)
    assert np.allclose(H, np.swapaxes(H, 1, 2))


def test_f1_f2_are_derivatives_of_f0():
    geom = mgb_b200.fem2d(2)
    M = O.amg_helper(geom, (("u", "dirichlet"), ("s", "full")), O.DEFAULT_D[2])
    n = geom.x.shape[0]
    z0 = np.array([O.DEFAULT_G[2](geom.x[i]) for i in range(n)]).reshape(-1, order="F")
    c = 0.3 * np.array([O.DEFAULT_F[2](geom.x[i]) for i in range(n)])
    Q = O.EuclidianPower(idx=[1, 2, 3], p=1.0)
    R = M.R_fine[-1]
    rng = np.random.default_rng(0)
    s = 1e-2 * rng.normal(size=R.shape[1])
    args = (geom.x, geom.w, c, R, M.D, z0, Q)
    g = O.f1(s, *args)
    H = O.f2(s, *args)
    d = rng.normal(size=s.size)
    h = 1e-5
    assert np.isclose(g @ d, (O.f0(s + h * d, *args) - O.f0(s - h * d, *args)) / (2 * h), rtol=1e-6)
    Hd = (O.f1(s + h * d, *args) - O.f1(s - h * d, *args)) / (2 * h)
    assert np.allclose(H @ d, Hd, rtol=1e-5, atol=1e-8)
    assert abs(H - H.T).max() < 1e-10


def test_hessian_loop_matches_dense_formula():
    # the literal-weight case of reference test/test_matrix_addition.jl:38-80 / test_d0_construction.jl:108-135
    g = mgb_b200.fem1d(2)
    n = g.x.shape[0]
    D = [g.operators["dx"], g.operators["id"]]
    y = np.zeros((n, 4)); y[:, 0] = 0.5; y[:, 1] = 0.1; y[:, 2] = 0.1; y[:, 3] = 0.3
    H = O.hessian_fine(y, g.w, D)
    Dd = [d.toarray() for d in D]
    W = lambda v: np.diag(g.w * v)
    ref = Dd[0].T @ W(y[:, 0]) @ Dd[0] + Dd[1].T @ W(y[:, 3]) @ Dd[1] + Dd[1].T @ W(y[:, 2]) @ Dd[0] + Dd[0].T @ W(y[:, 2]) @ Dd[1]
    assert np.abs(H.toarray() - ref).max() < 1e-12
    R = g.subspaces["dirichlet"][-1]
    RHR = (R.T @ H @ R).toarray()
    assert np.abs(RHR - R.toarray().T @ ref @ R.toarray()).max() < 1e-12


def test_oracle_solves_quick_cases():
    # reference test/test_quick.jl:108-140 (1D L=3 p=1) and test/test_2d.jl (2D L=2 p=2): the solve converges
    sol = O.amgb(mgb_b200.fem1d(3), p=1.0)
    assert sol.z.shape == (16, 2) and np.all(np.isfinite(sol.z))
    cd = sol.SOL_main["c_dot_Dz"]
    assert abs(cd[-1] - cd[-2]) < 1e-5
    sol = O.amgb(mgb_b200.fem2d(2), p=2.0)
    assert sol.z.shape == (56, 2)
    # feasibility of the final iterate: s >= |grad u|^p
    M = O.amg_helper(mgb_b200.fem2d(2), (("u", "dirichlet"), ("s", "full")), O.DEFAULT_D[2])
    Dz = O.apply_D(M.D, sol.z.reshape(-1, order="F"))
    assert np.all(Dz[:, 3] >= (Dz[:, 1] ** 2 + Dz[:, 2] ** 2) ** (2.0 / 2.0) - 1e-12)


def test_sparse_product_and_solve_kats():
    # reference test/test_basic_ops.jl:27-110: literal A (3x2), B (2x3); A*B, A'A and (A'A + 0.01 I) \ ones
    import scipy.sparse as sp
    A = sp.csc_matrix(np.array([[1.0, 0.0], [2.0, 3.0], [0.0, 4.0]]))
    B = sp.csc_matrix(np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]))
    assert np.array_equal((A @ B).toarray(), [[1, 2, 3], [14, 19, 24], [16, 20, 24]])
    AtA = (A.T @ A)
    assert np.array_equal(AtA.toarray(), [[5, 6], [6, 25]])
    x = O.solve(AtA + 0.01 * sp.identity(2, format="csc"), np.ones(2))
    xe = np.linalg.solve(np.array([[5.01, 6.0], [6.0, 25.01]]), np.ones(2))
    assert np.linalg.norm(x - xe) < 1e-10          # the reference's own tolerance (test_basic_ops.jl:93)
    # the product's solve seam stand-in gives the same answer
    from mgb_b200 import solver
    assert np.linalg.norm(solver.solve(AtA + 0.01 * sp.identity(2, format="csc"), np.ones(2)) - xe) < 1e-10


@pytest.mark.parametrize("p", [1.0, 1.3, 2.0, 2.7])
@pytest.mark.parametrize("variant", ["main3d", "main2d", "slack", "two_cones"])
def test_barrier_derivatives_match_automatic_differentiation(p, variant):
    """Upstream obtains F1 / F2 by automatic differentiation of the pointwise F (ForwardDiff); the oracle and the
    CUDA kernels use hand-derived formulas.  Pin them to autodiff of the same F (torch.autograd, float64) to
    rounding - a check by the reference's own method rather than by finite differences."""
    import torch
    rng = np.random.default_rng(7)
    if variant == "main3d":
        nD, sets = 5, [dict(idx=[1, 2, 3, 4], p=p, slack=False)]
    elif variant == "main2d":
        nD, sets = 4, [dict(idx=[1, 2, 3], p=p, slack=False)]
    elif variant == "slack":
        nD, sets = 5, [dict(idx=[1, 2, 3], p=p, slack=True)]
    else:  # parabolic: {s1 >= u^2} (p = 2) and {s2 >= |grad u|^p} on [u.id u.dx u.dy s1.id s2.id]
        nD, sets = 5, [dict(idx=[1, 2, 4], p=p, slack=False), dict(idx=[0, 3], p=2.0, slack=False)]
    Q = O.Intersection([O.EuclidianPower(**kw) for kw in sets])
    y = rng.normal(size=(8, nD)) * 0.3
    for kw in sets:
        y[:, kw["idx"][-1]] = 2.5 + rng.uniform(size=8)         # s well inside the cone
    if variant == "slack":
        y[:, -1] = rng.uniform(-0.5, 0.5, size=8)

    def F_torch(row):
        tot = row.new_zeros(())
        for kw in sets:
            q = row[kw["idx"][:-1]]
            s = row[kw["idx"][-1]] + (row[-1] if kw["slack"] else 0.0)
            a = 2.0 / kw["p"]
            tot = tot - torch.log(s ** a - (q * q).sum()) - O._mu(kw["p"]) * torch.log(s)
            if kw["slack"]:
                tot = tot - torch.log(1.0 + row[-1])
        return tot

    g, H = Q.F1(None, y), Q.F2(None, y)
    for r in range(y.shape[0]):
        row = torch.tensor(y[r], dtype=torch.float64, requires_grad=True)
        f_ad = float(F_torch(row).detach())
        assert abs(f_ad - float(Q.F(None, y[r:r + 1])[0])) <= 1e-13 * max(1.0, abs(f_ad))
        g_ad = torch.autograd.functional.jacobian(F_torch, row).numpy()
        H_ad = torch.autograd.functional.hessian(F_torch, row).numpy()
        assert np.abs(g[r] - g_ad).max() <= 1e-12 * max(1.0, np.abs(g_ad).max())
        assert np.abs(H[r] - H_ad).max() <= 1e-11 * max(1.0, np.abs(H_ad).max())


@pytest.mark.parametrize("gen,L,p,level", [("fem2d", 2, 1.0, None), ("fem2d", 2, 1.5, 0), ("fem1d", 3, 2.0, None), ("fem3d", 1, 1.0, None)])
def test_assembly_is_the_autodiff_gradient_and_hessian_of_the_objective(gen, L, p, level):
    """The whole assembly (apply_D, barrier map, w-scaling, the Hessian double loop, R'.R) against automatic
    differentiation of the objective s -> f0(s) written densely in torch float64: g = grad f0 and R'HR = hess f0 to
    rounding.  Exact (no finite-difference step), and independent of the sparse algebra the oracle uses."""
    import torch
    from helpers import problem
    geom = getattr(mgb_b200, gen)(L)
    pr = problem(geom, p=p, level=level, pert=1e-2)
    Q = O.EuclidianPower(idx=pr["idx"], p=pr["p"])
    t = 0.7
    args = (pr["x"], pr["w"], t * pr["c"], pr["R"], pr["D"], pr["z0"], Q)
    g, H = O.f1(pr["s"], *args), O.f2(pr["s"], *args).toarray()
    T = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
    Dd = [T(Dk.toarray()) for Dk in pr["D"]]
    Rd, z0, w, c = T(pr["R"].toarray()), T(pr["z0"]), T(pr["w"]), T(t * pr["c"])
    idx, a, mu = pr["idx"], 2.0 / p, O._mu(p)

    def f0_torch(s):
        z = z0 + Rd @ s
        Dz = torch.stack([Dk @ z for Dk in Dd], dim=1)
        q, ss = Dz[:, idx[:-1]], Dz[:, idx[-1]]
        F = -torch.log(ss ** a - (q * q).sum(dim=1)) - mu * torch.log(ss)
        return (w * F).sum() + (w[:, None] * c * Dz).sum()

    s = T(pr["s"]).requires_grad_(True)
    assert abs(float(f0_torch(s).detach()) - O.f0(pr["s"], *args)) <= 1e-12 * abs(O.f0(pr["s"], *args))
    g_ad = torch.autograd.functional.jacobian(f0_torch, s).numpy()
    H_ad = torch.autograd.functional.hessian(f0_torch, s).numpy()
    assert np.abs(g - g_ad).max() <= 1e-11 * np.abs(g_ad).max()
    assert np.abs(H - H_ad).max() <= 1e-11 * np.abs(H_ad).max()
