"""Sharding of the assembly over the GPUs of one box (one process per GPU): owner-computes.

Partitioning follows the reference's HPCSparseArrays layout: the outputs - gradient entries
and rows of R'HR - are split in contiguous blocks of the m unknowns, the quadrature rows in contiguous blocks of
whole broken elements.  A rank evaluates every element that touches one of ITS output rows and completes its rows of
R'HR and its block of the gradient locally: nothing but the three objective scalars crosses NVLink, and those cross
as self-validating peer-memory words written by the gather kernel itself (csrc/kernels.cuh dist_publish /
dist_collect).

Because R = blockdiag(R_u, R_s) numbers every u unknown before every s unknown, a contiguous block of the stacked
unknowns is uncorrelated with the element partition: each element is then evaluated by about two ranks (the owner of
its u rows and the owner of its s rows; at P = 2 every rank evaluates EVERY element).  ``colocated_partition``
renumbers the unknowns rank-major instead - rank r owns block r of every variable, stored back to back - which is
a column permutation of R at the boundary and nothing else (the library only ever sees "a contiguous block of the
unknowns"); a rank then evaluates E/P elements plus a thin halo.  ``create_peer_plan(colocate=True)`` (the default)
applies it and leaves the permutation in ``plan.perm``.

torch.distributed is used once per plan, to pass the 64-byte CUDA IPC handles of the scalar windows around.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch.distributed as dist

from .hpc import uniform_partition


def owner_of(part: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """rank owning each 0-based index under the 1-based offset vector ``part``."""
    return np.searchsorted(part[1:] - 1, idx, side="right")


def element_rows(n: int, block: int, rank: int, nranks: int):
    """[row0,row1) of this rank in the reference's row partition of x / w: whole elements, first ranks take the
    remainder (uniform_partition)."""
    part = uniform_partition(n, nranks, block)
    return int(part[rank] - 1), int(part[rank + 1] - 1)


def peer_partitions(n: int, m: int, block: int, nranks: int):
    """0-based offsets (length nranks+1): quadrature rows in whole elements, unknowns uniformly
    (uniform_partition: the HPCSparseArrays-style split, first ``units mod P`` ranks take one extra unit)."""
    return uniform_partition(n, nranks, block) - 1, uniform_partition(m, nranks) - 1


def variable_blocks(R, n: int) -> np.ndarray:
    """sizes of the column blocks of R = blockdiag(R_1, ..., R_nu) (rows [k n, (k+1) n) belong to variable k)"""
    Rc = R.tocsc()
    Rc.sort_indices()
    m = Rc.shape[1]
    nz = np.diff(Rc.indptr) > 0
    var = np.zeros(m, dtype=np.int64)
    var[nz] = Rc.indices[Rc.indptr[:-1][nz]] // n
    # an empty column (eliminated unknown) stays with its left neighbour's variable
    for j in np.nonzero(~nz)[0]:
        var[j] = var[j - 1] if j else 0
    if np.any(np.diff(var) < 0):
        raise ValueError("R is not block diagonal with the variables in order")
    return np.bincount(var, minlength=int(R.shape[0] // n))


def colocated_partition(sizes: Sequence[int], nranks: int):
    """Rank-major renumbering of unknowns stacked by variable: -> (perm, out_part).  ``perm[new] = old``; in the new
    numbering rank r owns the contiguous block [out_part[r], out_part[r+1]) = its uniform block of variable 0, then
    of variable 1, ...  Use as  R[:, perm], s[perm];  results come back as grad[perm] and H[perm][:, perm]."""
    sizes = [int(v) for v in sizes]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    parts = [uniform_partition(sz, nranks) - 1 for sz in sizes]
    perm = np.concatenate([np.arange(parts[k][r], parts[k][r + 1], dtype=np.int64) + offs[k]
                           for r in range(nranks) for k in range(len(sizes))])
    out_part = np.concatenate([[0], np.cumsum([sum(int(parts[k][r + 1] - parts[k][r]) for k in range(len(sizes)))
                                               for r in range(nranks)])]).astype(np.int64)
    return perm, out_part


def create_peer_plan(ctx, D, R, x, w, idx, p: float, block: int, rank: int, nranks: int, group=None, slack: bool = False,
                     idx2=None, p2: float = 2.0, colocate: bool = True):
    """Collective: every rank builds its DistPlan, the scalar windows are cross-mapped over CUDA IPC, and a
    barrier makes sure every window is mapped (and zero-initialised) before the first assembly.
    ``colocate``: renumber the unknowns rank-major (``plan.perm``, new -> old; ``plan.R`` = R[:, perm]) so that a rank
    owns the same block of every variable; ``plan.perm is None`` means the reference's contiguous numbering.
    Raises capi.MgbError("... sharded plans need the element path ...") on every rank alike when the level cannot
    be sharded (the refusal depends on the replicated operators only)."""
    from . import capi
    n, m = D[0].shape[0], R.shape[1]
    row_part, out_part = peer_partitions(n, m, block, nranks)
    perm = None
    if colocate and nranks > 1:
        sizes = variable_blocks(R, n)
        if len(sizes) > 1:
            perm, out_part = colocated_partition(sizes, nranks)
            R = R.tocsr()[:, perm].tocsr()
            R.sort_indices()
    plan = capi.DistPlan(ctx, D, R, x, w, idx, p, rank, nranks, row_part, out_part, slack=slack, idx2=idx2, p2=p2)
    plan.perm, plan.R = perm, R
    handles: List[Optional[bytes]] = [None] * nranks
    if nranks > 1:
        dist.all_gather_object(handles, plan.export_handle(), group=group)
    else:
        handles = [plan.export_handle()]
    plan.attach(handles)
    if nranks > 1:
        dist.barrier(group=group)
    return plan


def destroy_peer_plan(plan, group=None):
    """Collective: nobody unmaps/frees a window while a peer may still store into it."""
    plan.ctx.sync()
    if plan.nranks > 1:
        dist.barrier(group=group)
    plan.close()
