"""Two-cone barrier (intersection of convex sets) and parabolic_solve, after the reference's
test/test_parabolic.jl:36-104 (1D L=2, h=0.5, t1=1.0, p=2.0: ts equal, every snapshot < 1e-10)."""
import numpy as np
import pytest
import scipy.sparse as sp

import mgb_b200
from mgb_b200 import capi, solver
import mgb_oracle as O

from helpers import rel

pytestmark = pytest.mark.gpu


def _two_cone_problem(geom, p, seed=5, level=-1):
    dim = geom.dim
    Dt, idxA, idxB = O.parabolic_tables(dim)
    M = O.amg_helper(geom, O.PARABOLIC_STATE, Dt)
    n = geom.x.shape[0]
    rng = np.random.default_rng(seed)
    u = np.sin(geom.x[:, 0]) + (geom.x[:, 1] ** 2 if dim > 1 else 0.0)
    z0 = O.parabolic_feasible_start(M, u, dim, p)
    R = M.R_fine[level]
    s = 1e-3 * rng.uniform(-1, 1, size=R.shape[1])
    c = rng.normal(size=(n, len(Dt)))
    return M, R, z0, s, c, idxA, idxB


@pytest.mark.parametrize("path", [capi.PATH_ELEMENT, capi.PATH_CSR])
@pytest.mark.parametrize("gen,L,p", [("fem1d", 3, 2.0), ("fem2d", 2, 1.0), ("fem2d", 3, 1.5), ("fem2d", 4, 1.0)])
def test_two_cone_assembly_matches_oracle(gpu_ctx, gen, L, p, path):
    """both numeric paths: the fused element kernels (MODE 2) and the general CSR kernels"""
    geom = getattr(mgb_b200, gen)(L)
    M, R, z0, s, c, idxA, idxB = _two_cone_problem(geom, p)
    Q = O.Intersection([O.EuclidianPower(idx=idxA, p=2.0), O.EuclidianPower(idx=idxB, p=p)])
    t = 0.6
    args = (s, geom.x, geom.w, t * c, R, M.D, z0, Q)
    f0, g, H = O.f0(*args), O.f1(*args), O.f2(*args).tocsr()
    plan = capi.Plan(gpu_ctx, M.D, R, geom.x, geom.w, idxB, p, idx2=idxA, p2=2.0, force_path=path)
    assert plan.info["path"] == path
    assert capi.Plan(None, M.D, R, geom.x, geom.w, idxB, p, idx2=idxA, p2=2.0).info["path"] == capi.PATH_ELEMENT  # default
    Dz0 = np.stack([Dk @ z0 for Dk in M.D], axis=1)
    out = plan.assemble_host(s, Dz0, c, t, 7)
    rp, ci = plan.pattern()
    Hc = sp.csr_matrix((out["hval"], ci, rp), shape=(plan.m, plan.m))
    assert out["scal"][1] == 1.0
    assert abs(out["scal"][0] - f0) <= 1e-12 * abs(f0)
    assert rel(out["grad"], g) <= 1e-12
    assert abs(Hc - H).max() <= 1e-12 * abs(H).max()


@pytest.mark.parametrize("path", [capi.PATH_ELEMENT, capi.PATH_CSR])
@pytest.mark.parametrize("gen,L,level,p", [("fem2d", 3, 1, 1.0), ("fem2d", 3, 0, 2.0), ("fem1d", 4, 1, 1.5)])
def test_two_cone_coarse_levels(gpu_ctx, gen, L, level, p, path):
    """coarse multigrid levels: the id-like operators become dense element rows (the !FINE kernel instances)"""
    geom = getattr(mgb_b200, gen)(L)
    M, R, z0, s, c, idxA, idxB = _two_cone_problem(geom, p, level=level)
    Q = O.Intersection([O.EuclidianPower(idx=idxA, p=2.0), O.EuclidianPower(idx=idxB, p=p)])
    t = 0.6
    args = (s, geom.x, geom.w, t * c, R, M.D, z0, Q)
    f0, g, H = O.f0(*args), O.f1(*args), O.f2(*args).tocsr()
    plan = capi.Plan(gpu_ctx, M.D, R, geom.x, geom.w, idxB, p, idx2=idxA, p2=2.0, force_path=path)
    assert plan.info["path"] == path
    Dz0 = np.stack([Dk @ z0 for Dk in M.D], axis=1)
    out = plan.assemble_host(s, Dz0, c, t, 7)
    rp, ci = plan.pattern()
    Hc = sp.csr_matrix((out["hval"], ci, rp), shape=(plan.m, plan.m))
    assert out["scal"][1] == 1.0
    assert abs(out["scal"][0] - f0) <= 1e-12 * abs(f0)
    assert rel(out["grad"], g) <= 1e-12
    assert abs(Hc - H).max() <= 1e-12 * abs(H).max()
    Ho = H.copy(); Ho.eliminate_zeros()
    Pc = sp.csr_matrix((np.ones(Hc.nnz), Hc.indices, Hc.indptr), shape=Hc.shape)
    Po = sp.csr_matrix((np.ones(Ho.nnz), Ho.indices, Ho.indptr), shape=Ho.shape)
    assert (Po - Po.multiply(Pc)).nnz == 0, "oracle pattern not contained in plan pattern"


def test_parabolic_1d_reference_case():
    geom = mgb_b200.fem1d(2)
    sol = solver.parabolic_solve(geom, h=0.5, t1=1.0, p=2.0)
    ref = O.parabolic_solve(geom, h=0.5, t1=1.0, p=2.0)
    assert len(sol.ts) >= 2 and len(sol.u) == len(sol.ts)
    assert np.array_equal(sol.ts, ref.ts)
    for a, b in zip(sol.u, ref.u):
        assert np.linalg.norm(a - b) < 1e-10 * max(1.0, np.linalg.norm(b))


def test_parabolic_2d_and_api_wrapper():
    from mgb_b200 import api
    g = api.fem2d_mpi(L=2)
    sol = api.parabolic_solve(g, h=0.25, t1=0.5, p=1.0)
    soln = api.mpi_to_native(sol)
    ref = O.parabolic_solve(mgb_b200.fem2d(2), h=0.25, t1=0.5, p=1.0)
    assert np.array_equal(soln.ts, ref.ts) and len(soln.u) == 3
    for a, b in zip(soln.u, ref.u):
        assert a.shape == (56, 3)
        assert np.linalg.norm(a - b) < 1e-9 * np.linalg.norm(b)
