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
    """the element_kernel + gather_kernel pair kept for A/B comparison against the patch-fused kernel"""
    plan, _ = check_against_oracle(gpu_ctx, getattr(mgb_b200, gen)(L), 1.0, t=0.9, level=level, slack=slack,
                                   force_path=capi.PLAN_TWO_STAGE)
    assert plan.info["path"] == capi.PATH_ELEMENT


@pytest.mark.parametrize("patch", ["16", "32", "64"])
def test_patch_fused_path(gpu_ctx, patch, monkeypatch):
    """opt-in patch-fused kernel (MGB_PATCH): CTA-local records in shared memory + interface partials"""
    monkeypatch.setenv("MGB_PATCH", patch)
    check_against_oracle(gpu_ctx, mgb_b200.fem2d(4), 1.0, t=0.9)
    check_against_oracle(gpu_ctx, mgb_b200.fem2d(4), 1.5, t=0.9, level=1)
    check_against_oracle(gpu_ctx, mgb_b200.fem2d(3), 1.0, t=0.9, slack=True)
    check_against_oracle(gpu_ctx, mgb_b200.fem1d(6), 1.0, t=0.9)


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
