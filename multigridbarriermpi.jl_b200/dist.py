"""Sharding of the assembly over the GPUs of one box (one process per GPU): owner-computes.

Partitioning follows the reference's HPCSparseArrays layout: the outputs - gradient entries
and rows of R'HR - are split in contiguous blocks of the m unknowns, the quadrature rows in contiguous blocks of
whole broken elements.  Because R = blockdiag(R_u, R_s) numbers every u unknown before every s unknown, the two
partitions are uncorrelated: exchanging element contributions would move 50 % (P = 2) to 94 % (P = 8) of everything
a rank assembles.  Instead a rank evaluates every element that touches one of ITS output rows (each element is
evaluated by about two ranks, the owner of its u rows and the owner of its s rows) and completes its rows of R'HR
and its block of the gradient locally: nothing but the three objective scalars crosses NVLink, and those cross as
self-validating peer-memory words written by the gather kernel itself (csrc/kernels.cuh dist_publish / dist_collect).

torch.distributed is used once per plan, to pass the 64-byte CUDA IPC handles of the scalar windows around.
"""
from __future__ import annotations

from typing import List, Optional

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


def create_peer_plan(ctx, D, R, x, w, idx, p: float, block: int, rank: int, nranks: int, group=None, slack: bool = False,
                     idx2=None, p2: float = 2.0):
    """Collective: every rank builds its DistPlan, the scalar windows are cross-mapped over CUDA IPC, and a
    barrier makes sure every window is mapped (and zero-initialised) before the first assembly.
    Raises capi.MgbError("... sharded plans need the element path ...") on every rank alike when the level cannot
    be sharded (the refusal depends on the replicated operators only)."""
    from . import capi
    n, m = D[0].shape[0], R.shape[1]
    row_part, out_part = peer_partitions(n, m, block, nranks)
    plan = capi.DistPlan(ctx, D, R, x, w, idx, p, rank, nranks, row_part, out_part, slack=slack, idx2=idx2, p2=p2)
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
