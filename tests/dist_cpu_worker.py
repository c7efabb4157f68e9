"""torchrun worker (gloo, CPU): every rank builds only its own symbolic DistPlan; the owned pattern blocks, gathered
over torch.distributed, must reassemble the global pattern, and the primary rows must partition the quadrature rows."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import mgb_b200  # noqa: E402
from mgb_b200 import capi, dist as mdist  # noqa: E402
from helpers import problem  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
for gen, L in (("fem2d", 4), ("fem1d", 6)):
    geom = getattr(mgb_b200, gen)(L)
    pr = problem(geom)
    n, m = geom.x.shape[0], pr["R"].shape[1]
    row_part, out_part = mdist.peer_partitions(n, m, geom.block, world)
    pl = capi.DistPlan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0, rank, world, row_part, out_part)
    rp, ci = pl.own_pattern()
    blocks = [None] * world
    dist.all_gather_object(blocks, (pl.dinfo["own0"], rp, ci, pl.rows[: pl.dinfo["n_primary"]]))
    grp, gci = capi.Plan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0).pattern()
    rowptr = np.concatenate([[0]] + [np.diff(b[1]) for b in sorted(blocks, key=lambda b: b[0])]).cumsum()
    colidx = np.concatenate([b[2] for b in sorted(blocks, key=lambda b: b[0])])
    assert np.array_equal(rowptr, grp) and np.array_equal(colidx, gci)
    assert np.array_equal(np.sort(np.concatenate([b[3] for b in blocks])), np.arange(n))
# the same with the rank-major numbering of the unknowns (dist.colocated_partition): every rank permutes the columns of
# R alike, the owned blocks tile the permuted pattern, and a rank evaluates fewer quadrature rows than with contiguous
# blocks of the stacked unknowns
import scipy.sparse as sp  # noqa: E402
geom = mgb_b200.fem2d(4)
pr = problem(geom)
n, m = geom.x.shape[0], pr["R"].shape[1]
perm, out_part = mdist.colocated_partition(mdist.variable_blocks(pr["R"], n), world)
row_part, cont_part = mdist.peer_partitions(n, m, geom.block, world)
Rp = pr["R"].tocsr()[:, perm].tocsr()
pl = capi.DistPlan(None, pr["D"], Rp, pr["x"], pr["w"], pr["idx"], 1.0, rank, world, row_part, out_part)
pc = capi.DistPlan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0, rank, world, row_part, cont_part)
assert pl.dinfo["n_rows"] < pc.dinfo["n_rows"], (pl.dinfo["n_rows"], pc.dinfo["n_rows"])
rp, ci = pl.own_pattern()
blocks = [None] * world
dist.all_gather_object(blocks, (pl.dinfo["own0"], rp, ci))
grp, gci = capi.Plan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0).pattern()
H = sp.csr_matrix((np.ones(len(gci)), gci.astype(np.int64), grp.astype(np.int64)), shape=(m, m))
Hp = H[perm][:, perm].tocsr()
Hp.sort_indices()
rowptr = np.concatenate([[0]] + [np.diff(b[1]) for b in sorted(blocks, key=lambda b: b[0])]).cumsum()
colidx = np.concatenate([b[2] for b in sorted(blocks, key=lambda b: b[0])])
assert np.array_equal(rowptr, Hp.indptr) and np.array_equal(colidx, Hp.indices)
# HPC-typed geometry: every rank keeps only its row block of each operator; the whole matrix comes back with one
# collective gather (what the symbolic phase of a level needs of R, once - reference src/MultiGridBarrierMPI.jl:357-371)
from mgb_b200 import hpc  # noqa: E402
geom = mgb_b200.fem2d(3)
be = hpc.Backend(device="cpu", rank=rank, nranks=world)
for name in sorted(geom.operators):
    A = geom.operators[name].tocsr()
    H = hpc.HPCSparseMatrix(A, be)
    lo, hi = H.row_partition[rank] - 1, H.row_partition[rank + 1] - 1
    assert H.local.shape[0] == hi - lo and H.nnz == A[lo:hi].nnz and H.nnz < A.nnz
    W = H.gather()
    assert (abs(W - A)).nnz == 0 and W.shape == A.shape
assert hpc.GATHER_STATS["calls"] == len(geom.operators) and hpc.GATHER_STATS["bytes"] > 0
dist.barrier()
if rank == 0:
    print("GLOO_OK")
dist.destroy_process_group()
