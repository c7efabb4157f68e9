"""Host logic of the sharded (owner-computes) plans on CPU: symbolic-only DistPlans (ctx=None).

  * in-process: for 1..5 ranks the owned patterns tile the global pattern bit-exactly, every quadrature row is
    primary on exactly one rank, and a rank's rows contain every element that touches its output rows;
  * two processes over gloo (world_size 2): each rank builds only its own plan, the owned blocks are gathered with
    torch.distributed and must reassemble the global pattern (the N>1 host path without a GPU)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

import mgb_b200
from mgb_b200 import capi
from mgb_b200.hpc import uniform_partition

from helpers import problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("gen,L,slack", [("fem2d", 3, False), ("fem2d", 4, False), ("fem1d", 5, False), ("fem2d", 2, True)])
@pytest.mark.parametrize("nranks", [1, 2, 3, 5])
def test_owned_patterns_tile_the_global_pattern(gen, L, slack, nranks):
    geom = getattr(mgb_b200, gen)(L)
    pr = problem(geom, slack=slack)
    n, m = geom.x.shape[0], pr["R"].shape[1]
    glob = capi.Plan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0, slack=slack)
    grp, gci = glob.pattern()
    row_part, out_part = uniform_partition(n, nranks, geom.block) - 1, uniform_partition(m, nranks) - 1
    E = [(Dk @ pr["R"]).tocsr() for Dk in pr["D"]]
    touched = sp.csr_matrix(sum(abs(Ek) for Ek in E))          # n x m: dofs every quadrature row touches
    prim_all = []
    for r in range(nranks):
        pl = capi.DistPlan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0, r, nranks, row_part, out_part, slack=slack)
        d = pl.dinfo
        lo, hi = d["own0"], d["own1"]
        assert (lo, hi) == (out_part[r], out_part[r + 1])
        rp, ci = pl.own_pattern()
        assert np.array_equal(rp, grp[lo:hi + 1] - grp[lo]) and np.array_equal(ci, gci[grp[lo]:grp[hi]])
        rows = pl.rows
        assert len(np.unique(rows)) == len(rows) and len(rows) % geom.block == 0
        prim_all.append(rows[: d["n_primary"]])
        # completeness: every quadrature row with a dof in [lo, hi) is evaluated by this rank
        need = np.flatnonzero(np.diff(touched[:, lo:hi].tocsr().indptr) > 0)
        assert np.isin(need, rows).all()
    assert np.array_equal(np.sort(np.concatenate(prim_all)), np.arange(n))


def test_coarse_levels_and_unstructured_operators_refuse_to_shard():
    """the refusal every rank sees alike (solver.LevelState then assembles the level redundantly)"""
    geom = mgb_b200.fem2d(4)
    pr = problem(geom, level=0)
    n, m = geom.x.shape[0], pr["R"].shape[1]
    rp, op = uniform_partition(n, 2, geom.block) - 1, uniform_partition(m, 2) - 1
    with pytest.raises(capi.MgbError, match="sharded plans need the element path"):
        capi.DistPlan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0, 0, 2, rp, op)
    g3 = mgb_b200.fem3d(1)
    p3 = problem(g3)
    n, m = g3.x.shape[0], p3["R"].shape[1]
    rp, op = uniform_partition(n, 2, g3.block) - 1, uniform_partition(m, 2) - 1
    with pytest.raises(capi.MgbError, match="sharded plans need the element path"):
        capi.DistPlan(None, p3["D"], p3["R"], p3["x"], p3["w"], p3["idx"], 1.0, 0, 2, rp, op)


def test_two_process_gloo_plans_reassemble_the_global_pattern():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29647", os.path.join(ROOT, "tests", "dist_cpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "GLOO_OK" in res.stdout
