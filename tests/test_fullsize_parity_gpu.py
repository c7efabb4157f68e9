"""CUDA path vs the CPU oracle AT the sizes BASELINE.json names (north_star tolerance: gradient / Hessian values
1e-12 relative, objective 1e-12, the oracle's pattern contained in the plan's frozen pattern) - the same comparison
the small-mesh parity tests make (helpers.check_against_oracle), after the reference's own H_mpi vs H_native loop
(test/test_map_rows_compare.jl:102-123,165-171; `norm(H_mpi - H_native) < 1e-12`, test/test_matrix_addition.jl:84-95).

  C3  fem2d L=8  p=1.0   n = 229,376   (headline; element path)
  C2  fem1d L=16         n = 131,072   (element path, 2-node elements)
  C4  fem3d L=5  k=3     n = 262,144   (CSR path, SELL/chunk replay)
  C5  fem2d L=7 two-cone n = 57,344    (parabolic barrier, element path MODE 2)

The oracle needs 0.3 - 5 s per case at these sizes.
"""
import numpy as np
import pytest
import scipy.sparse as sp

import mgb_b200
from mgb_b200 import capi
import mgb_oracle as O

from helpers import check_against_oracle, rel

pytestmark = pytest.mark.gpu


def test_C3_fem2d_L8_matches_oracle(gpu_ctx):
    geom = mgb_b200.fem2d(8)
    assert geom.x.shape[0] == 229376                       # docs/src/guide.md:253
    plan, out = check_against_oracle(gpu_ctx, geom, 1.0, t=1.0)
    assert plan.info["path"] == capi.PATH_ELEMENT and plan.m == 196610


def test_C3_fem2d_L8_coarse_level_matches_oracle(gpu_ctx):
    """one coarse level of the L=8 hierarchy at full n (north_star item 4: Galerkin levels use the same machinery)"""
    check_against_oracle(gpu_ctx, mgb_b200.fem2d(8), 1.0, t=1.0, level=5)


def test_C2_fem1d_L16_matches_oracle(gpu_ctx):
    """fem1d L=16 is conditioning-limited in float64: the derivative operator has entries of 2^16, so the reference's
    own association D*(z0 + R*s) (test/test_apply_d.jl:44) - which the oracle follows - puts 4e-12 of absolute rounding
    into Dz, 4e-9 relative into the gradient and 3e-12 into the Hessian.  The arbiter here is the same assembly in
    80-bit long double (helpers.oracle_eval_longdouble): the CUDA path must match it to the north_star tolerance
    (1e-12; 1e-11 for the gradient, whose entries are sums of cancelling +-2^16-weighted terms), and must sit at
    least as close to it as the float64 oracle does; CUDA vs the float64 oracle is bounded by the oracle's own
    distance to the long-double result."""
    from helpers import problem, oracle_eval, oracle_eval_longdouble, cuda_eval
    geom = mgb_b200.fem1d(16)
    assert geom.x.shape[0] == 2 ** 17
    # the 1-D feasible set is thin at L=16 (element size 2^-16): small perturbation of the lifted start
    pr = problem(geom, p=1.0, pert=1e-8)
    t = 1.0
    f0_o, g_o, H_o = oracle_eval(pr, t)
    f0_l, g_l, H_l = oracle_eval_longdouble(pr, t)
    plan, out, H_c = cuda_eval(gpu_ctx, pr, t)
    assert plan.info["path"] == capi.PATH_ELEMENT and out["scal"][1] == 1.0
    g_l64 = g_l.astype(np.float64)
    err_g_cuda, err_g_orc = rel(out["grad"], g_l64), rel(g_o, g_l64)
    hn = abs(H_l).max()
    err_h_cuda, err_h_orc = abs(H_c - H_l).max() / hn, abs(H_o - H_l).max() / hn
    assert abs(out["scal"][0] - f0_l) <= 1e-12 * abs(f0_l)
    assert err_g_cuda <= 1e-11, err_g_cuda
    assert err_h_cuda <= 1e-12, err_h_cuda
    assert err_g_cuda <= err_g_orc and err_h_cuda <= max(err_h_orc, 1e-14), (err_g_cuda, err_g_orc, err_h_cuda, err_h_orc)
    # against the float64 oracle: within the oracle's own rounding distance to the long-double result
    assert rel(out["grad"], g_o) <= 2.0 * err_g_orc + 1e-12
    assert abs(H_c - H_o).max() / hn <= 2.0 * err_h_orc + 1e-12
    # pattern: the oracle's (cancellation-dependent) pattern is contained in the plan's
    Ho = H_o.copy(); Ho.eliminate_zeros()
    Pc = sp.csr_matrix((np.ones(H_c.nnz), H_c.indices, H_c.indptr), shape=H_c.shape)
    Po = sp.csr_matrix((np.ones(Ho.nnz), Ho.indices, Ho.indptr), shape=Ho.shape)
    assert (Po - Po.multiply(Pc)).nnz == 0


def test_C4_fem3d_L5_matches_oracle(gpu_ctx):
    geom = mgb_b200.fem3d(5)
    assert geom.x.shape[0] == 262144
    plan, out = check_against_oracle(gpu_ctx, geom, 1.0, t=1.0)
    assert plan.info["path"] == capi.PATH_CSR


def test_C5_two_cone_fem2d_L7_matches_oracle(gpu_ctx):
    """the barrier upstream parabolic_solve uses (reference test/test_parabolic.jl:48, docs/src/guide.md:360-367):
    {s1 >= u^2} with {s2 >= |grad u|^p} on [u.id; u.dx; u.dy; s1.id; s2.id]"""
    geom = mgb_b200.fem2d(7)
    assert geom.x.shape[0] == 57344                        # docs/src/guide.md:252
    dim, p, t = 2, 1.0, 0.6
    Dt, idxA, idxB = O.parabolic_tables(dim)
    M = O.amg_helper(geom, O.PARABOLIC_STATE, Dt)
    n = geom.x.shape[0]
    rng = np.random.default_rng(5)
    u = np.sin(geom.x[:, 0]) + geom.x[:, 1] ** 2
    z0 = O.parabolic_feasible_start(M, u, dim, p)
    R = M.R_fine[-1]
    s = 1e-4 * rng.uniform(-1, 1, size=R.shape[1])
    c = rng.normal(size=(n, len(Dt)))
    Q = O.Intersection([O.EuclidianPower(idx=idxA, p=2.0), O.EuclidianPower(idx=idxB, p=p)])
    args = (s, geom.x, geom.w, t * c, R, M.D, z0, Q)
    f0, g, H = O.f0(*args), O.f1(*args), O.f2(*args).tocsr()
    assert np.isfinite(f0)
    plan = capi.Plan(gpu_ctx, M.D, R, geom.x, geom.w, idxB, p, idx2=idxA, p2=2.0)
    assert plan.info["path"] == capi.PATH_ELEMENT
    Dz0 = np.stack([Dk @ z0 for Dk in M.D], axis=1)
    out = plan.assemble_host(s, Dz0, c, t, 7)
    rp, ci = plan.pattern()
    Hc = sp.csr_matrix((out["hval"], ci, rp), shape=(plan.m, plan.m))
    assert out["scal"][1] == 1.0
    assert abs(out["scal"][0] - f0) <= 1e-12 * abs(f0)
    assert rel(out["grad"], g) <= 1e-12
    assert abs(Hc - H).max() <= 1e-12 * abs(H).max()
    assert sp.linalg.norm(Hc - H) <= 1e-12 * sp.linalg.norm(H)
