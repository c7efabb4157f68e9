"""Row-block sharding of the assembly over the GPUs of one box (one process per GPU).

Partitioning follows the reference's HPCSparseArrays layout: quadrature rows are
split in contiguous blocks (whole broken elements, so apply_D needs no halo), the outputs - gradient
entries and rows of R'HR - are split in contiguous blocks of the m unknowns.  Every rank assembles
the contributions of its own quadrature rows on its local pattern; the contributions to rows owned by
another rank (the interface between neighbouring row blocks) travel once per assembly in a single
all-to-all, and are summed at the owner in fixed source-rank order (bit-reproducible).  The scalars
(objective, <c,Dz>, feasibility) ride in the same all-to-all (every rank receives every partial and
sums them in rank order).

The index maps are built once per level (symbolic phase) from replicated structural information plus
one integer all-to-all; the per-assembly work is: pack kernel -> all_to_all_single -> unpack kernels.
The same code runs on CPU tensors with the gloo backend (tests/test_dist_cpu.py) - there the packing
uses torch index ops and the local values come from the test, the CUDA kernels are not involved.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist

from .hpc import uniform_partition


def owner_of(part: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """rank owning each 0-based index under the 1-based offset vector ``part``."""
    return np.searchsorted(part[1:] - 1, idx, side="right")


@dataclass
class ExchangePlan:
    """Frozen maps of one level's interface exchange (per rank)."""
    rank: int
    nranks: int
    m_part: np.ndarray            # 1-based offsets of the unknowns
    own_rowptr: np.ndarray        # CSR of the owned rows of the global pattern (0-based, local row ids)
    own_colidx: np.ndarray
    # Hessian values
    h_send_idx: torch.Tensor      # positions in the local value array, grouped by destination rank
    h_send_splits: List[int]
    h_recv_pos: torch.Tensor      # positions in the owned value array, grouped by source rank
    h_recv_splits: List[int]
    # gradient entries
    g_send_idx: torch.Tensor
    g_send_splits: List[int]
    g_recv_pos: torch.Tensor
    g_recv_splits: List[int]
    n_own_h: int = 0
    n_own_g: int = 0


def _exchange_int_lists(lists: List[np.ndarray], device, group=None) -> List[np.ndarray]:
    """all-to-all of variable-length int64 lists (setup only)."""
    P = len(lists)
    send_counts = torch.tensor([len(x) for x in lists], dtype=torch.int64, device=device)
    recv_counts = torch.empty(P, dtype=torch.int64, device=device)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    rc = [int(v) for v in recv_counts.cpu()]
    sendbuf = torch.from_numpy(np.concatenate(lists).astype(np.int64) if P else np.zeros(0, np.int64)).to(device)
    recvbuf = torch.empty(sum(rc), dtype=torch.int64, device=device)
    dist.all_to_all_single(recvbuf, sendbuf, output_split_sizes=rc, input_split_sizes=[len(x) for x in lists],
                           group=group)
    out, o = [], 0
    rb = recvbuf.cpu().numpy()
    for c in rc:
        out.append(rb[o:o + c])
        o += c
    return out


def build_exchange(rank: int, nranks: int, m: int, glob_rowptr: np.ndarray, glob_colidx: np.ndarray,
                   loc_rowptr: np.ndarray, loc_colidx: np.ndarray, device, m_part: Optional[np.ndarray] = None,
                   group=None) -> ExchangePlan:
    """``glob_*``: global pattern of R'HR (replicated, from a symbolic-only plan over all rows);
    ``loc_*``: pattern of this rank's local plan (m rows, global column ids)."""
    m_part = uniform_partition(m, nranks) if m_part is None else m_part
    lo, hi = int(m_part[rank] - 1), int(m_part[rank + 1] - 1)
    own_rowptr = (glob_rowptr[lo:hi + 1] - glob_rowptr[lo]).astype(np.int64)
    own_colidx = glob_colidx[glob_rowptr[lo]:glob_rowptr[hi]].astype(np.int64)
    # ---- Hessian: every local entry (a,b) -> (owner(a), position inside the owner's block)
    rows_loc = np.repeat(np.arange(m, dtype=np.int64), np.diff(loc_rowptr))
    own = owner_of(m_part, rows_loc)
    # position of (a,b) in the global pattern: search b inside row a of the global CSR
    gpos = np.empty(rows_loc.size, dtype=np.int64)
    key_g = np.repeat(np.arange(m, dtype=np.int64), np.diff(glob_rowptr)) * (m + 1) + glob_colidx.astype(np.int64)
    key_l = rows_loc * (m + 1) + loc_colidx.astype(np.int64)
    gpos = np.searchsorted(key_g, key_l)
    if gpos.size and not np.array_equal(key_g[gpos], key_l):
        raise RuntimeError("local pattern is not contained in the global pattern")
    blk_start = glob_rowptr[(m_part[:-1] - 1).astype(np.int64)].astype(np.int64)  # first global position of each owner
    pos_in_owner = gpos - blk_start[own]
    order = np.argsort(own, kind="stable")
    h_send_idx = order.astype(np.int64)
    h_send_splits = [int(c) for c in np.bincount(own, minlength=nranks)]
    send_lists, o = [], 0
    for r in range(nranks):
        send_lists.append(pos_in_owner[order[o:o + h_send_splits[r]]])
        o += h_send_splits[r]
    recv_lists = _exchange_int_lists(send_lists, device, group)
    # ---- gradient: local rows with at least one entry touch dof a
    touched = np.flatnonzero(np.diff(loc_rowptr) > 0).astype(np.int64)
    gown = owner_of(m_part, touched)
    gorder = np.argsort(gown, kind="stable")
    g_send_splits = [int(c) for c in np.bincount(gown, minlength=nranks)]
    gsend, o = [], 0
    for r in range(nranks):
        sel = touched[gorder[o:o + g_send_splits[r]]]
        gsend.append(sel - int(m_part[r] - 1))
        o += g_send_splits[r]
    grecv = _exchange_int_lists(gsend, device, group)
    td = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(device)
    return ExchangePlan(rank, nranks, m_part, own_rowptr, own_colidx,
                        td(h_send_idx), h_send_splits, td(np.concatenate(recv_lists) if recv_lists else np.zeros(0)),
                        [len(x) for x in recv_lists],
                        td(touched[gorder]), g_send_splits, td(np.concatenate(grecv) if grecv else np.zeros(0)),
                        [len(x) for x in grecv], n_own_h=int(own_colidx.size), n_own_g=hi - lo)


class Exchanger:
    """Per-assembly interface exchange in ONE collective.

    The local outputs live in one buffer ``loc = [hval | grad | scal(4)]`` (``views()`` hands the three
    windows to mgb_assemble).  Per step: one pack kernel (gather_idx) builds the send buffer
    ``[to rank 0 | to rank 1 | ...]``, each segment = H entries owned by that rank, gradient entries
    owned by that rank, the 4 scalars; one ``all_to_all_single``; one owner-side kernel (segsum_idx)
    sums, for every owned output, its contributions in source-rank order (bit-reproducible) into
    ``own = [H rows owned | gradient block owned | scal(4)]``.
    ``ctx`` (capi.Context) selects the CUDA kernels; ``ctx=None`` uses torch index ops (CPU / gloo tests)."""

    def __init__(self, ex: ExchangePlan, device, ctx=None, group=None, n_loc_h: int = None, m: int = None):
        self.ex, self.device, self.ctx, self.group = ex, device, ctx, group
        P = ex.nranks
        f64 = torch.float64
        self.n_loc_h = int(n_loc_h if n_loc_h is not None else ex.h_send_idx.numel())
        self.m = int(m if m is not None else ex.m_part[-1] - 1)
        self.loc = torch.zeros(self.n_loc_h + self.m + 4, dtype=f64, device=device)
        hs, gs = ex.h_send_idx.cpu().numpy(), ex.g_send_idx.cpu().numpy()
        hr, gr = ex.h_recv_pos.cpu().numpy(), ex.g_recv_pos.cpu().numpy()
        scal_pos = np.arange(4, dtype=np.int64) + self.n_loc_h + self.m
        send_idx, self.send_splits, self.recv_splits = [], [], []
        oh = og = 0
        for r in range(P):
            send_idx += [hs[oh:oh + ex.h_send_splits[r]], self.n_loc_h + gs[og:og + ex.g_send_splits[r]], scal_pos]
            self.send_splits.append(ex.h_send_splits[r] + ex.g_send_splits[r] + 4)
            oh += ex.h_send_splits[r]
            og += ex.g_send_splits[r]
        send_idx = np.concatenate(send_idx).astype(np.int64)
        # owner side: destination (in `own`) of every received value, then CSR by destination
        dest, oh, og = [], 0, 0
        for r in range(P):
            dest += [hr[oh:oh + ex.h_recv_splits[r]], ex.n_own_h + gr[og:og + ex.g_recv_splits[r]],
                     ex.n_own_h + ex.n_own_g + np.arange(4, dtype=np.int64)]
            self.recv_splits.append(ex.h_recv_splits[r] + ex.g_recv_splits[r] + 4)
            oh += ex.h_recv_splits[r]
            og += ex.g_recv_splits[r]
        dest = np.concatenate(dest).astype(np.int64)
        n_out = ex.n_own_h + ex.n_own_g + 4
        order = np.argsort(dest, kind="stable")          # stable: contributions stay in source-rank order
        ptr = np.zeros(n_out + 1, dtype=np.int64)
        np.add.at(ptr, dest + 1, 1)
        ptr = np.cumsum(ptr)
        self.n_out = n_out
        self.send = torch.empty(send_idx.size, dtype=f64, device=device)
        self.recv = torch.empty(dest.size, dtype=f64, device=device)
        self.own = torch.zeros(n_out, dtype=f64, device=device)
        td = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(device)
        self.send_idx64 = td(send_idx, torch.int64)
        self.red_idx64 = td(order, torch.int64)
        self.red_ptr64 = td(ptr, torch.int64)
        if ctx is not None:
            assert self.loc.numel() < 2 ** 31 and dest.size < 2 ** 31
            self.send_idx32 = td(send_idx, torch.int32)
            self.red_idx32 = td(order, torch.int32)
            self.red_ptr32 = td(ptr, torch.int32)
        else:  # CPU: segment ids for index_add
            self.red_seg = td(np.repeat(np.arange(n_out), np.diff(ptr)), torch.int64)

    def views(self):
        """(hval, grad, scal) windows of the local output buffer, to be passed to mgb_assemble."""
        h, m = self.n_loc_h, self.m
        return self.loc[:h], self.loc[h:h + m], self.loc[h + m:h + m + 4]

    def exchange(self):
        """pack -> all_to_all -> owner-side sum.  Returns (H values of the owned rows, owned gradient
        block, reduced scalars {f0, all_finite, <c,Dz>, nonfinite count})."""
        ex = self.ex
        if self.ctx is None:
            torch.index_select(self.loc, 0, self.send_idx64, out=self.send)
        else:
            self.ctx.gather_idx(self.loc, self.send_idx32, self.send.numel(), self.send)
        dist.all_to_all_single(self.recv, self.send, output_split_sizes=self.recv_splits,
                               input_split_sizes=self.send_splits, group=self.group)
        if self.ctx is None:
            self.own.zero_()
            self.own.index_add_(0, self.red_seg, self.recv.index_select(0, self.red_idx64))
        else:
            self.ctx.segsum_idx(self.recv, self.red_ptr32, self.red_idx32, self.n_out, self.own)
        h_own = self.own[: ex.n_own_h]
        g_own = self.own[ex.n_own_h: ex.n_own_h + ex.n_own_g]
        scal = self.own[ex.n_own_h + ex.n_own_g:]
        scal[1] = (scal[3] == 0).to(scal.dtype)
        return h_own, g_own, scal


def element_rows(n: int, block: int, rank: int, nranks: int):
    """[row0,row1) of this rank: whole elements, first ranks take the remainder (uniform_partition)."""
    part = uniform_partition(n, nranks, block)
    return int(part[rank] - 1), int(part[rank + 1] - 1)


# ---------------------------------------------------------------------------------------------------
# Fused peer-memory exchange (mgb_dist_*): no collective on the data path.  torch.distributed is used
# once, at setup, to pass the 64-byte CUDA IPC handles of the exchange windows between the processes.

def peer_partitions(n: int, m: int, block: int, nranks: int):
    """0-based offsets (length nranks+1): quadrature rows in whole elements, unknowns uniformly
    (uniform_partition: the HPCSparseArrays-style split, first ``units mod P`` ranks take one extra unit)."""
    return uniform_partition(n, nranks, block) - 1, uniform_partition(m, nranks) - 1


def create_peer_plan(ctx, D, R, x, w, idx, p: float, block: int, rank: int, nranks: int, group=None, slack: bool = False):
    """Collective: every rank builds its DistPlan, the windows are cross-mapped over CUDA IPC, and a
    barrier makes sure every window is mapped (and zero-initialised) before the first assembly."""
    from . import capi
    n, m = D[0].shape[0], R.shape[1]
    row_part, out_part = peer_partitions(n, m, block, nranks)
    plan = capi.DistPlan(ctx, D, R, x, w, idx, p, rank, nranks, row_part, out_part, slack=slack)
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
