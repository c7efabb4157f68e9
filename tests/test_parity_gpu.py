"""GPU parity: CUDA assembly (through the C ABI, host buffers) vs the CPU oracle on seeded inputs."""
import numpy as np
import pytest

import mgb_b200
from mgb_b200 import capi

from helpers import check_against_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("L,p", [(1, 1.0), (2, 2.0), (3, 1.0), (3, 1.5), (4, 1.0)])
def test_fem2d_fine_level(gpu_ctx, L, p):
    plan, _ = check_against_oracle(gpu_ctx, mgb_b200.fem2d(L), p, t=0.7)
    assert plan.info["path"] == capi.PATH_ELEMENT
    assert plan.info["nodes_per_element"] == 7


@pytest.mark.parametrize("L,p", [(1, 1.0), (3, 1.0), (4, 2.0), (6, 1.5)])
def test_fem1d_fine_level(gpu_ctx, L, p):
    plan, _ = check_against_oracle(gpu_ctx, mgb_b200.fem1d(L), p, t=1.3)
    assert plan.info["path"] == capi.PATH_ELEMENT


@pytest.mark.parametrize("level", [0, 1, 2])
def test_fem2d_coarse_levels(gpu_ctx, level):
    check_against_oracle(gpu_ctx, mgb_b200.fem2d(4), 1.0, t=2.0, level=level)


@pytest.mark.parametrize("level", [0, 2])
def test_fem1d_coarse_levels(gpu_ctx, level):
    check_against_oracle(gpu_ctx, mgb_b200.fem1d(5), 2.0, t=2.0, level=level)


@pytest.mark.parametrize("gen,L", [("fem1d", 3), ("fem2d", 2), ("fem2d", 3)])
def test_feasibility_slack_variant(gpu_ctx, gen, L):
    check_against_oracle(gpu_ctx, getattr(mgb_b200, gen)(L), 1.0, t=0.5, slack=True)
    check_against_oracle(gpu_ctx, getattr(mgb_b200, gen)(L), 2.0, t=0.5, slack=True, level=0)


@pytest.mark.parametrize("gen,L,level,slack", [("fem1d", 4, None, False), ("fem2d", 4, None, False), ("fem2d", 4, 1, False),
                                                  ("fem2d", 3, None, True), ("fem2d", 3, 0, True)])
def test_two_stage_element_path(gpu_ctx, gen, L, level, slack):
    """MGB_PLAN_TWO_STAGE is accepted for compatibility (the two-stage element path is the only one)"""
    plan, _ = check_against_oracle(gpu_ctx, getattr(mgb_b200, gen)(L), 1.0, t=0.9, level=level, slack=slack,
                                   force_path=capi.PLAN_TWO_STAGE)
    assert plan.info["path"] == capi.PATH_ELEMENT


def test_bitwise_reproducible(gpu_ctx):
    """fixed summation order: two runs give identical bits"""
    from helpers import problem, cuda_eval
    pr = problem(mgb_b200.fem2d(5))
    _, o1, H1 = cuda_eval(gpu_ctx, pr, 0.7)
    _, o2, H2 = cuda_eval(gpu_ctx, pr, 0.7)
    assert np.array_equal(H1.data, H2.data) and np.array_equal(o1["grad"], o2["grad"]) and o1["scal"][0] == o2["scal"][0]


@pytest.mark.parametrize("gen,L,level", [("fem1d", 4, None), ("fem2d", 3, None), ("fem2d", 3, 1), ("fem3d", 2, None)])
def test_csr_path(gpu_ctx, gen, L, level):
    geom = mgb_b200.fem3d(L, k=1) if gen == "fem3d" else getattr(mgb_b200, gen)(L)
    plan, _ = check_against_oracle(gpu_ctx, geom, 1.0, t=0.9, level=level, force_path=capi.PATH_CSR)
    assert plan.info["path"] == capi.PATH_CSR


def test_fem3d_q3_default_table(gpu_ctx):
    """config C4 at a size the oracle finishes in seconds: fem3d k=3 (64-node elements, nD=5), reference
    problem data src/MultiGridBarrierMPI.jl:736-738"""
    check_against_oracle(gpu_ctx, mgb_b200.fem3d(2), 1.0, t=0.9)
    check_against_oracle(gpu_ctx, mgb_b200.fem3d(2), 1.5, t=0.9, level=0)


@pytest.mark.parametrize("gen,L,nranks", [("fem2d", 3, 2), ("fem1d", 5, 3)])
def test_plans_from_hpc_local_blocks_sum_to_the_oracle(gpu_ctx, gen, L, nranks):
    """every rank builds its plan from its own HPCSparseMatrix storage (mgb_plan_create_local: compressed column
    ids + col_indices, 1-based); the rank contributions add up to the oracle's gradient and Hessian"""
    import scipy.sparse as sp
    from helpers import problem, oracle_eval, rel
    from mgb_b200.hpc import Backend, HPCSparseMatrix, uniform_partition
    geom = getattr(mgb_b200, gen)(L)
    pr = problem(geom)
    n = pr["x"].shape[0]
    part = uniform_partition(n, nranks, block=geom.block)
    f0_o, g_o, H_o = oracle_eval(pr, 0.8)
    Dz0 = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)
    f0 = 0.0; g = 0.0; H = None
    for rank in range(nranks):
        be = Backend(device="cpu", rank=rank, nranks=nranks)
        lo, hi = int(part[rank] - 1), int(part[rank + 1] - 1)
        blocks = [HPCSparseMatrix(Dk, be, row_partition=part).local_storage() for Dk in pr["D"]]
        plan = capi.Plan.from_local_blocks(gpu_ctx, blocks, pr["R"], n, pr["x"][lo:hi], pr["w"][lo:hi], pr["idx"], 1.0)
        out = plan.assemble_host(pr["s"], Dz0[lo:hi], pr["c"][lo:hi], 0.8, capi.WANT_F0 | capi.WANT_GRAD | capi.WANT_HESS)
        rp, ci = plan.pattern()   # the rows this rank's quadrature points touch: differs per rank
        Hr = sp.csr_matrix((out["hval"], ci.astype(np.int64), rp.astype(np.int64)), shape=(plan.m, plan.m))
        f0 += out["scal"][0]; g = g + out["grad"]; H = Hr if H is None else H + Hr
    assert abs(f0 - f0_o) <= 1e-12 * max(1.0, abs(f0_o))
    assert rel(g, g_o) <= 1e-12
    assert abs(H - H_o).max() <= 1e-12 * abs(H_o).max()


@pytest.mark.parametrize("chunk,sigma", [("2", "64"), ("5", "1"), ("1000000", "16384")])
def test_csr_replay_chunking_and_windows(gpu_ctx, chunk, sigma, monkeypatch):
    """CSR path replay lists: chunk = 2 cuts nearly every list of apply_D, gradient and Hessian into runs whose
    partial sums a second kernel adds (all three combine kernels run); chunk = 10^6 never cuts.  The sorting
    window changes which lanes sit in a slice, never a value."""
    monkeypatch.setenv("MGB_SELL_CHUNK", chunk)
    monkeypatch.setenv("MGB_SELL_SIGMA", sigma)
    monkeypatch.setenv("MGB_SELL_SIGMA_GRAD", sigma)
    check_against_oracle(gpu_ctx, mgb_b200.fem3d(2), 1.0, t=0.9)
    check_against_oracle(gpu_ctx, mgb_b200.fem2d(3), 1.5, t=0.9, level=1, force_path=capi.PATH_CSR)
    check_against_oracle(gpu_ctx, mgb_b200.fem2d(2), 1.0, t=0.5, slack=True, force_path=capi.PATH_CSR)
    check_against_oracle(gpu_ctx, mgb_b200.fem1d(5), 2.0, t=0.9, level=0, force_path=capi.PATH_CSR)


@pytest.mark.parametrize("gen,L,level", [("fem2d", 5, None), ("fem2d", 4, 1), ("fem1d", 9, None), ("fem1d", 6, 0)])
def test_objective_only_call_has_the_bits_of_the_full_assembly(gpu_ctx, gen, L, level):
    """line-search points (flags = WANT_F0) run other kernel instances than a full assembly; the objective they
    return must be the same number bit for bit (the Newton driver compares one against the other)"""
    from helpers import problem
    pr = problem(getattr(mgb_b200, gen)(L), level=level)
    plan = capi.Plan(gpu_ctx, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0)
    Dz0 = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)
    full = plan.assemble_host(pr["s"], Dz0, pr["c"], 0.7, 7)["scal"].copy()
    for rep in range(3):
        one = plan.assemble_host(pr["s"], Dz0, pr["c"], 0.7, 1)["scal"].copy()
        assert np.array_equal(one, full), (one, full)


@pytest.mark.parametrize("L,level,p", [(2, 0, 1.0), (3, 0, 1.0), (3, 1, 1.0), (3, 1, 1.5), (4, 2, 1.0)])
def test_fem3d_coarse_levels_dense_path(gpu_ctx, L, level, p):
    """coarse levels of fem3d's Q3 hexahedra: every fine point touches up to 64 unknowns per variable - the dense
    contraction kernel (kernels_dense.cuh: per chunk of a coarse element full uu / us / ss blocks, then the ordinary
    gather) instead of product lists (reference problem data src/MultiGridBarrierMPI.jl:735-745)"""
    plan, _ = check_against_oracle(gpu_ctx, mgb_b200.fem3d(L), p, t=0.9, level=level)
    assert plan.info["path"] == capi.PATH_ELEMENT and plan.info["nodes_per_element"] == 64


@pytest.mark.parametrize("level,p", [(0, 1.0), (1, 1.5), (2, 1.0)])
def test_dense_path_partial_tiles(gpu_ctx, monkeypatch, level, p):
    """Q2 hexahedra (27 points per element) forced onto the dense path (it would not pay there): groups of two or three
    fine elements make chunks of 54 / 81 points, so the 16-point tiles of the TMA ring end in partial tiles - the
    zero-filled tail must not contribute - and at the finest level a group's dofs fill the 64-wide block only partly."""
    monkeypatch.setenv("MGB_DENSE_MIN_PRODUCTS", "0")
    plan, _ = check_against_oracle(gpu_ctx, mgb_b200.fem3d(3, k=2), p, t=0.9, level=level)
    assert plan.info["path"] == capi.PATH_ELEMENT and plan.info["nodes_per_element"] == 64
