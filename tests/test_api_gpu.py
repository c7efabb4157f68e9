"""The reference's unit-level checks of its overloads (test/test_helpers.jl:47-167) against the
Python mirror of the same names, plus the integration pattern of test/test_quick.jl."""
import numpy as np
import pytest
import scipy.sparse as sp

import mgb_b200
import mgb_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from mgb_b200 import api as a
    return a


@pytest.fixture(scope="module")
def be(api):
    return api.backend_cuda()


def test_amgb_zeros_diag_blockdiag(api, be):
    from mgb_b200.hpc import HPCSparseMatrix, HPCMatrix, HPCVector
    A = HPCSparseMatrix(sp.csr_matrix((10, 10)), be)
    Z = api.amgb_zeros(A, 5, 5)
    assert isinstance(Z, HPCSparseMatrix) and Z.shape == (5, 5)
    Zd = api.amgb_zeros(HPCMatrix(np.zeros((10, 10)), be), 4, 6)
    assert isinstance(Zd, HPCMatrix) and Zd.shape == (4, 6)
    v = HPCVector(np.array([1.0, 2, 3]), be)
    Dm = api.amgb_diag(A, v)
    assert isinstance(Dm, HPCSparseMatrix) and Dm.shape == (3, 3)
    assert np.array_equal(api.to_host(Dm).diagonal(), [1, 2, 3])
    D4 = api.amgb_diag(A, np.array([1.0, 2, 3, 4]))
    assert D4.shape == (4, 4)
    # reference test/test_diag.jl:28-46
    assert np.array_equal(api.to_host(api.amgb_diag(A, np.arange(1.0, 11.0))).diagonal(), np.arange(1.0, 11.0))
    C = api.amgb_blockdiag(HPCSparseMatrix(sp.identity(2, format="csr"), be), HPCSparseMatrix(sp.identity(3, format="csr"), be))
    assert C.shape == (5, 5)


def test_amgb_all_isfinite(api, be):
    from mgb_b200.hpc import HPCVector, HPCMatrix
    assert api.amgb_all_isfinite(HPCVector(np.array([1.0, 2, 3]), be)) is True
    assert api.amgb_all_isfinite(HPCVector(np.array([1.0, np.inf, 3]), be)) is False
    assert api.amgb_all_isfinite(HPCMatrix(np.array([[1.0, np.nan], [0, 1]]), be)) is False
    assert api.amgb_all_isfinite(HPCVector(np.zeros(0), be)) is True   # empty


def test_vector_reductions(api, be, gpu_ctx):
    """dot / sum / norm of HPCVectors (reference tools/profile_scaling.jl:89-109): deterministic device reductions"""
    from mgb_b200.hpc import HPCVector
    rng = np.random.default_rng(5)
    for n in (1, 7, 256, 1000, 300001):
        a, b = rng.standard_normal(n), rng.standard_normal(n)
        x, y = HPCVector(a, be), HPCVector(b, be)
        assert abs(api.dot(x, y) - a @ b) <= 1e-13 * (np.abs(a) @ np.abs(b))
        assert abs(api.vsum(x) - a.sum()) <= 1e-13 * np.abs(a).sum()
        assert abs(api.norm(x) - np.linalg.norm(a)) <= 1e-14 * np.linalg.norm(a)
        assert api.norm(x, np.inf) == np.abs(a).max()
        assert api.dot(x, y) == api.dot(x, y)            # same bits on every call
    a = np.array([1.0, np.nan, 2.0])
    assert np.isnan(api.vsum(HPCVector(a, be)))          # non-finite entries are data, not errors
    assert gpu_ctx.reduce("sum", None, 0) == 0.0         # empty local block (more ranks than rows)


def test_map_rows_kats(api, be):
    from mgb_b200.hpc import HPCVector, HPCMatrix
    x = HPCMatrix(np.array([[1.0, 2], [3, 4], [5, 6]]), be)
    r = api.map_rows(lambda row: sum(row), x)
    assert isinstance(r, HPCVector) and np.array_equal(api.to_host(r), [3, 7, 11])
    x = HPCMatrix(np.array([[1.0, 2], [3, 4]]), be)
    r = api.map_rows(lambda row: (row.sum(), row.prod()), x)
    assert isinstance(r, HPCMatrix) and np.array_equal(api.to_host(r), [[3, 2], [7, 12]])
    y = HPCVector(np.array([10.0, 20.0]), be)
    r = api.map_rows(lambda rx, ry: sum(rx) + ry[0], x, y)
    assert np.array_equal(api.to_host(r), [13, 27])
    r = api.map_rows_gpu(lambda rx, ry: (rx[0] ** 2, ry[0] * rx[1]), x, y)
    assert np.array_equal(api.to_host(r), [[1, 20], [9, 80]])


@pytest.mark.parametrize("p,slack", [(1.0, False), (1.5, False), (2.0, True)])
def test_map_rows_barrier_matches_oracle(api, be, p, slack):
    from mgb_b200.hpc import HPCMatrix
    rng = np.random.default_rng(3)
    n, nD = 1000, 5 if slack else 4
    Dz = rng.normal(size=(n, nD)) * 0.4
    Dz[:, 3] = 2.5 + rng.uniform(size=n)
    if slack:
        Dz[:, 4] = rng.uniform(-0.5, 0.5, size=n)
    x = HPCMatrix(rng.normal(size=(n, 2)), be)
    Dzm = HPCMatrix(Dz, be)
    Q = api.convex_Euclidian_power([1, 2, 3], p)
    Q.slack = slack
    Qo = O.EuclidianPower(idx=[1, 2, 3], p=p, slack=slack)
    F = api.to_host(api.map_rows_gpu(Q.F, x, Dzm))
    F1 = api.to_host(api.map_rows_gpu(Q.F1, x, Dzm))
    F2 = api.to_host(api.map_rows_gpu(Q.F2, x, Dzm))
    assert np.allclose(F, Qo.F(None, Dz), rtol=1e-13, atol=1e-13)
    assert np.allclose(F1, Qo.F1(None, Dz), rtol=1e-13, atol=1e-13)
    assert np.allclose(F2, Qo.F2(None, Dz).reshape(n, nD * nD), rtol=1e-13, atol=1e-13)


def test_geometry_roundtrip_and_solve(api, be):
    # reference test/test_quick.jl:52-140
    from mgb_b200.hpc import HPCMatrix, HPCVector, HPCSparseMatrix
    g = api.fem1d_mpi(L=3, backend=be)
    assert isinstance(g.x, HPCMatrix) and isinstance(g.w, HPCVector) and isinstance(g.operators["id"], HPCSparseMatrix)
    gn = api.mpi_to_native(g)
    ref = mgb_b200.fem1d(3)
    assert np.allclose(gn.x, ref.x) and np.allclose(gn.w, ref.w)
    assert (gn.operators["dx"] - ref.operators["dx"]).nnz == 0
    assert g.operators["id"].rowptr.dtype == np.int32 and g.operators["id"].rowptr[0] == 1
    sol = api.amgb(g, p=1.0, verbose=False, tol=1e-10)
    soln = api.mpi_to_native(sol)
    assert soln.z.shape == (16, 2)
    sol_ref = O.amgb(ref, p=1.0, tol=1e-10)
    assert np.linalg.norm(soln.z - sol_ref.z) < 1e-10 * 1000   # reference tolerance: TOL*1000, test_quick.jl:140


def test_fem2d_mpi_solve_wrapper(api):
    sol = api.fem2d_mpi_solve(L=2, p=2.0, verbose=False)       # reference test/test_2d.jl L=2 p=2
    soln = api.mpi_to_native(sol)
    sol_ref = O.amgb(mgb_b200.fem2d(2), p=2.0)
    assert np.linalg.norm(soln.z - sol_ref.z) / np.linalg.norm(sol_ref.z) < 1e-9
    assert np.array_equal(soln.SOL_main["its"], sol_ref.SOL_main["its"])


def test_spmat_products_both_orientations(gpu_ctx):
    """HPCSparseMatrix * HPCVector and A' * v (reference test/test_nonsquare.jl:45-72) through mgb_spmat_mv:
    short rows (thread per row) and the restriction of a coarse level, whose transposed rows hold thousands of
    entries (chunked two-stage path); alpha / beta / in-place update as the Newton driver uses them."""
    import torch
    from mgb_b200 import capi
    from helpers import problem
    dev = torch.device("cuda", gpu_ctx.device)
    rng = np.random.default_rng(9)
    mats = [sp.random(300, 120, density=0.05, random_state=3, format="csr"),
            problem(mgb_b200.fem2d(5), level=0)["R"], problem(mgb_b200.fem1d(12), level=1)["R"]]
    for A in mats:
        A = sp.csr_matrix(A)
        M = capi.SpMat(gpu_ctx, A)
        assert max(np.diff(A.tocsc().indptr)) >= 1024 or A.shape[0] == 300
        for trans in (False, True):
            op = A.T if trans else A
            x = rng.standard_normal(op.shape[1]); y0 = rng.standard_normal(op.shape[0])
            x_d, y0_d = torch.from_numpy(x).to(dev), torch.from_numpy(y0).to(dev)
            y_d = torch.empty(op.shape[0], dtype=torch.float64, device=dev)
            M.mv(x_d, y_d, trans=trans)
            ref = op @ x
            assert np.abs(y_d.cpu().numpy() - ref).max() <= 1e-13 * (abs(op) @ np.abs(x)).max()
            M.mv(x_d, y0_d, trans=trans, alpha=-0.5, beta=2.0, y0_dev=y0_d)      # in place: y0 <- 2 y0 - A x / 2
            ref2 = 2.0 * y0 - 0.5 * ref
            assert np.abs(y0_d.cpu().numpy() - ref2).max() <= 1e-13 * ((abs(op) @ np.abs(x)).max() + np.abs(y0).max())
        M.close()


def test_diag_scale_numeric(gpu_ctx):
    """mgb_diag_scale = the diagonal amgb_diag wraps (reference src:137-147, test/test_diag.jl:28-46: spdiagm of 1:10):
    out = w .* y[:, col] on the device"""
    import torch
    dev = torch.device("cuda", gpu_ctx.device)
    n, k = 1000, 5
    rng = np.random.default_rng(1)
    w, y = rng.normal(size=n), rng.normal(size=(n, k))
    w_d = torch.from_numpy(w).to(dev)
    y_d = torch.from_numpy(np.ascontiguousarray(y.T)).to(dev)      # column-major n x k
    out = torch.zeros(n, dtype=torch.float64, device=dev)
    for col in (0, 3, 4):
        gpu_ctx.diag_scale(w_d, y_d, n, n, col, out)
        assert np.array_equal(out.cpu().numpy(), w * y[:, col])
    # the reference's literal: amgb_diag(1:10) has the diagonal 1..10
    one = torch.ones(10, dtype=torch.float64, device=dev)
    z = torch.arange(1, 11, dtype=torch.float64, device=dev)
    o10 = torch.zeros(10, dtype=torch.float64, device=dev)
    gpu_ctx.diag_scale(one, z, 10, 10, 0, o10)
    assert o10.cpu().tolist() == [float(v) for v in range(1, 11)]
