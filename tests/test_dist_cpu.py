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


@pytest.mark.parametrize("gen,L,slack", [("fem2d", 4, False), ("fem1d", 6, False), ("fem2d", 3, True)])
@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_colocated_ownership_is_a_renumbering_and_halves_the_element_work(gen, L, slack, nranks):
    """dist.colocated_partition: rank r owns block r of EVERY variable (unknowns renumbered rank-major = a column
    permutation of R).  The owned patterns tile the permuted global pattern, and a rank evaluates about E/P elements
    instead of the 2E/P of contiguous blocks of the stacked unknowns."""
    from mgb_b200 import dist as mdist
    geom = getattr(mgb_b200, gen)(L)
    pr = problem(geom, slack=slack)
    n, m = geom.x.shape[0], pr["R"].shape[1]
    sizes = mdist.variable_blocks(pr["R"], n)
    assert sizes.sum() == m and len(sizes) == (3 if slack else 2)
    perm, out_part = mdist.colocated_partition(sizes, nranks)
    assert np.array_equal(np.sort(perm), np.arange(m)) and out_part[0] == 0 and out_part[-1] == m
    offs = np.concatenate([[0], np.cumsum(sizes)])
    for r in range(nranks):   # the owned block holds the rank's uniform block of every variable, back to back
        own = perm[out_part[r]:out_part[r + 1]]
        for k, sz in enumerate(sizes):
            part = uniform_partition(int(sz), nranks) - 1
            assert np.array_equal(own[(own >= offs[k]) & (own < offs[k + 1])], np.arange(part[r], part[r + 1]) + offs[k])
    Rp = pr["R"].tocsr()[:, perm].tocsr()
    grp, gci = capi.Plan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0, slack=slack).pattern()
    H = sp.csr_matrix((np.ones(len(gci)), gci.astype(np.int64), grp.astype(np.int64)), shape=(m, m))
    Hp = H[perm][:, perm].tocsr()
    Hp.sort_indices()
    row_part = uniform_partition(n, nranks, geom.block) - 1
    cont_part = uniform_partition(m, nranks) - 1
    prim_all, el_col, el_cont = [], [], []
    for r in range(nranks):
        pl = capi.DistPlan(None, pr["D"], Rp, pr["x"], pr["w"], pr["idx"], 1.0, r, nranks, row_part, out_part, slack=slack)
        lo, hi = pl.dinfo["own0"], pl.dinfo["own1"]
        rp, ci = pl.own_pattern()
        assert np.array_equal(rp, Hp.indptr[lo:hi + 1] - Hp.indptr[lo]) and np.array_equal(ci, Hp.indices[Hp.indptr[lo]:Hp.indptr[hi]])
        prim_all.append(pl.rows[: pl.dinfo["n_primary"]])
        el_col.append(pl.dinfo["elements"])
        el_cont.append(capi.DistPlan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0, r, nranks, row_part, cont_part,
                                     slack=slack).dinfo["elements"])
    assert np.array_equal(np.sort(np.concatenate(prim_all)), np.arange(n))
    assert max(el_col) < 0.8 * max(el_cont), (el_col, el_cont)


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
