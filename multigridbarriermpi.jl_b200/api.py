"""The reference's operator / plugin surface, same names and argument meaning, over the B200 path.

Reference: the ten MultiGridBarrier generics MultiGridBarrierMPI.jl overloads
(src/MultiGridBarrierMPI.jl:62-192) and its public wrappers (src:259-338, 355-528, 559-745).
Julia keyword arguments map to Python keyword arguments one to one.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Sequence

import numpy as np
import scipy.sparse as sp
import torch

from . import capi, geometry as geom_mod, solver
from .amg import DEFAULT_D, DEFAULT_F, DEFAULT_G
from .geometry import Geometry
from .hpc import Backend, HPCMatrix, HPCSparseMatrix, HPCVector, backend_cuda, uniform_partition

_CTX = {}


def _ctx(backend: Backend) -> capi.Context:
    """Backend instance cache, like _GPU_BACKEND_CACHE (src:80-114): one context per device."""
    key = backend.index
    if key not in _CTX:
        torch.cuda.set_device(backend.index)
        _CTX[key] = capi.Context(backend.index, torch.cuda.current_stream().cuda_stream)
    return _CTX[key]


def _dist_ready() -> bool:
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


# ------------------------------------------------------------------ amgb_zeros (src:66-75, 116-117)
def amgb_zeros(proto, m: int, n: Optional[int] = None, backend: Optional[Backend] = None):
    if proto is HPCVector or isinstance(proto, HPCVector):
        be = backend or proto.backend
        return HPCVector(np.zeros(m), be)
    if isinstance(proto, HPCSparseMatrix):
        return HPCSparseMatrix(sp.csr_matrix((m, n)), proto.backend, Ti=proto.Ti)
    if isinstance(proto, HPCMatrix):
        return HPCMatrix(np.zeros((m, n)), proto.backend)
    raise TypeError(f"amgb_zeros: unsupported prototype {type(proto)}")


# ------------------------------------------------------------------ amgb_all_isfinite (src:121-133)
def amgb_all_isfinite(z) -> bool:
    """Local device reduction (mgb_all_isfinite kernel) + AND over ranks."""
    buf = z.v if isinstance(z, HPCVector) else z.A
    ok = _ctx(z.backend).all_isfinite(buf, buf.numel())
    if _dist_ready():
        import torch.distributed as dist
        t = torch.tensor([1 if ok else 0], device=buf.device, dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item())
    return ok


# ------------------------------------------------------------------ dot / sum / norm (HPCVector reductions)
def _reduce(op: str, x, y=None) -> float:
    """local deterministic device reduction (mgb_reduce) + all-reduce of the scalar over the ranks
    (reference: local reduce + Allreduce, tools/profile_scaling.jl:89-109)"""
    bx = x.v if isinstance(x, HPCVector) else x.A
    by = None if y is None else (y.v if isinstance(y, HPCVector) else y.A)
    if by is not None and by.numel() != bx.numel():
        raise ValueError("dot: vectors of different local length (partitions differ)")
    val = _ctx(x.backend).reduce(op, bx, bx.numel(), by)
    if _dist_ready():
        import torch.distributed as dist
        t = torch.tensor([val], device=bx.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "maxabs" else dist.ReduceOp.SUM)
        val = float(t.item())
    return val


def dot(x, y) -> float:
    return _reduce("dot", x, y)


def vsum(x) -> float:
    return _reduce("sum", x)


def norm(x, ord=2) -> float:
    """2-norm (default) or max-norm (ord=inf) of an HPCVector / HPCMatrix"""
    if ord in (np.inf, float("inf"), "inf"):
        return _reduce("maxabs", x)
    if ord != 2:
        raise ValueError("norm: only the 2-norm and the max-norm are implemented")
    return float(np.sqrt(_reduce("norm2sq", x)))


# ------------------------------------------------------------------ amgb_diag (src:137-147)
def amgb_diag(proto, z, m: Optional[int] = None, n: Optional[int] = None) -> HPCSparseMatrix:
    """spdiagm(m, n, 0 => z) as an HPCSparseMatrix with the prototype's index type (Int32 for a dense
    prototype).  The fused assembly never builds these (w .* y is applied inside the kernels); this is
    the drop-in for callers that still ask for the diagonal."""
    zh = to_host(z) if isinstance(z, HPCVector) else np.asarray(z, dtype=np.float64)
    m = len(zh) if m is None else m
    n = len(zh) if n is None else n
    k = min(m, n, len(zh))
    Dm = sp.csr_matrix((zh[:k], (np.arange(k), np.arange(k))), shape=(m, n))
    Ti = proto.Ti if isinstance(proto, HPCSparseMatrix) else np.int32
    return HPCSparseMatrix(Dm, proto.backend, Ti=Ti)


# ------------------------------------------------------------------ amgb_blockdiag (src:150)
def amgb_blockdiag(*args: HPCSparseMatrix) -> HPCSparseMatrix:
    return HPCSparseMatrix(sp.block_diag([a.host for a in args], format="csr"), args[0].backend, Ti=args[0].Ti)


# ------------------------------------------------------------------ map_rows / map_rows_gpu (src:161-170)
class _Row:
    """A whole column block presented as 'one row': row[k] is column k for all local rows at once, so
    a row closure written with indexing / sum / prod evaluates for every row in one pass on the device
    (the reference's map_rows_gpu contract: closures see broadcastable data, no scalar indexing)."""

    def __init__(self, cols: torch.Tensor):
        self._c = cols  # (k, n_local)

    def __getitem__(self, k):
        return self._c[k]

    def __len__(self):
        return self._c.shape[0]

    def __iter__(self):
        return iter(self._c)

    def sum(self):
        return self._c.sum(dim=0)

    def prod(self):
        return self._c.prod(dim=0)


@dataclass
class BarrierFn:
    """F / F1 / F2 of a convex set; recognised by map_rows(_gpu) and routed to the CUDA barrier map
    (mgb_map_barrier) instead of being evaluated as a closure."""
    which: int
    idx: Sequence[int]
    p: float
    slack: bool = False


@dataclass
class ConvexSet:
    idx: Sequence[int]
    p: float
    slack: bool = False

    @property
    def F(self):
        return BarrierFn(0, self.idx, self.p, self.slack)

    @property
    def F1(self):
        return BarrierFn(1, self.idx, self.p, self.slack)

    @property
    def F2(self):
        return BarrierFn(2, self.idx, self.p, self.slack)


def convex_Euclidian_power(idx: Sequence[int], p: float = 2.0) -> ConvexSet:
    """upstream convex_Euclidian_power(idx=..., p=x->p): {(q, s) = y[idx] : s >= |q|^p}; idx 0-based."""
    return ConvexSet(list(idx), float(p))


def _cols(a) -> torch.Tensor:
    return a.v.unsqueeze(0) if isinstance(a, HPCVector) else a.A


def map_rows(f: Callable, A, *args):
    """Apply ``f`` to aligned local rows of every argument (scalar result -> HPCVector, k results ->
    n x k HPCMatrix; KATs: reference test/test_helpers.jl:123-167)."""
    allargs = (A,) + args
    be = A.backend
    part = A.partition if isinstance(A, HPCVector) else A.row_partition
    if isinstance(f, BarrierFn):
        x, Dz = allargs[0], allargs[1]
        nloc, nD = Dz.A.shape[1], Dz.A.shape[0]
        width = (1, nD, nD * nD)[f.which]
        out = torch.empty((width, nloc), dtype=torch.float64, device=Dz.A.device)
        _ctx(be).map_barrier(f.idx, f.p, f.slack, nD, nloc, Dz.A, f.which, out)
        return HPCVector(out[0], be, partition=part, local=True) if f.which == 0 else HPCMatrix(out, be, part, local=True)
    rows = [_Row(_cols(a)) for a in allargs]
    res = f(*rows)
    nloc = rows[0][0].shape[0]
    if isinstance(res, torch.Tensor) and res.dim() == 1:
        return HPCVector(res.contiguous(), be, partition=part, local=True)
    if isinstance(res, (int, float)):
        return HPCVector(torch.full((nloc,), float(res), dtype=torch.float64, device=rows[0][0].device), be, part, True)
    cols = [c if isinstance(c, torch.Tensor) else torch.full((nloc,), float(c), dtype=torch.float64,
                                                             device=rows[0][0].device) for c in res]
    return HPCMatrix(torch.stack(cols).contiguous(), be, part, local=True)


def map_rows_gpu(f: Callable, A, *args):
    return map_rows(f, A, *args)


# ------------------------------------------------------------------ raw-array accessors (src:175-192)
def _raw_array(x):
    return x.v if isinstance(x, HPCVector) else x.A


def _rows_to_svectors(M):
    return M.v if isinstance(M, HPCVector) else M.A


def _to_cpu_array(x):
    if isinstance(x, np.ndarray):
        return x
    return x.v.cpu().numpy() if isinstance(x, HPCVector) else x.A.cpu().numpy().T.copy()


def vertex_indices(A):
    n = len(A) if isinstance(A, HPCVector) else A.shape[0]
    return np.arange(1, n + 1)


# ------------------------------------------------------------------ gathers (Matrix()/Vector()/SparseMatrixCSC())
def to_host(x):
    """Collective gather of an HPC value to a native array on every rank (Vector()/Matrix() of the
    reference, src:357-360, 525-527)."""
    if isinstance(x, HPCSparseMatrix):
        return x.host.copy()
    loc = x.v if isinstance(x, HPCVector) else x.A
    if not _dist_ready():
        h = loc.cpu().numpy()
        return h if isinstance(x, HPCVector) else h.T.copy()
    import torch.distributed as dist
    part = x.partition if isinstance(x, HPCVector) else x.row_partition
    sizes = np.diff(part)
    k = 1 if isinstance(x, HPCVector) else x.k
    outs = [torch.empty((k, int(sz)), dtype=torch.float64, device=loc.device) for sz in sizes]
    dist.all_gather(outs, loc.reshape(k, -1).contiguous())
    full = torch.cat(outs, dim=1).cpu().numpy()
    return full[0] if isinstance(x, HPCVector) else full.T.copy()


# ------------------------------------------------------------------ native_to_mpi / mpi_to_native (src:259-528)
def native_to_mpi(g_native: Geometry, Ti=np.int32, backend: Optional[Backend] = None) -> Geometry:
    be = backend or backend_cuda()
    part = uniform_partition(g_native.x.shape[0], be.nranks, g_native.block)
    conv = lambda op: HPCSparseMatrix(op, be, Ti=Ti)
    operators = {k: conv(g_native.operators[k]) for k in sorted(g_native.operators)}   # sorted keys: src:276
    subspaces = {k: [conv(m) for m in g_native.subspaces[k]] for k in sorted(g_native.subspaces)}  # src:284
    return Geometry(g_native.discretization, HPCMatrix(g_native.x, be, part), HPCVector(g_native.w, be, part),
                    subspaces, operators, [conv(m) for m in g_native.refine], [conv(m) for m in g_native.coarsen],
                    block=g_native.block, meta=dict(g_native.meta))


def _convert_to_native(x):
    if isinstance(x, (HPCMatrix, HPCVector, HPCSparseMatrix)):
        return to_host(x)
    if isinstance(x, (list, tuple)):
        return type(x)(_convert_to_native(v) for v in x)
    if isinstance(x, dict):
        return {k: _convert_to_native(v) for k, v in x.items()}
    return x


def mpi_to_native(obj):
    """Geometry / AMGBSOL / ParabolicSOL with HPC types -> native arrays (collective)."""
    if isinstance(obj, Geometry):
        return Geometry(obj.discretization, to_host(obj.x), to_host(obj.w),
                        {k: [to_host(m) for m in obj.subspaces[k]] for k in sorted(obj.subspaces)},
                        {k: to_host(obj.operators[k]) for k in sorted(obj.operators)},
                        [to_host(m) for m in obj.refine], [to_host(m) for m in obj.coarsen],
                        block=obj.block, meta=dict(obj.meta))
    if isinstance(obj, solver.AMGBSOL):
        return solver.AMGBSOL(_convert_to_native(obj.z), _convert_to_native(obj.SOL_feasibility),
                              _convert_to_native(obj.SOL_main), obj.log, mpi_to_native(obj.geometry), dict(obj.stats))
    if isinstance(obj, solver.ParabolicSOL):
        return solver.ParabolicSOL(mpi_to_native(obj.geometry), obj.ts, [_convert_to_native(u) for u in obj.u])
    return _convert_to_native(obj)


# ------------------------------------------------------------------ public wrappers (src:559-745)
_GEOM_KEYS = {"fem1d": ("L",), "fem2d": ("L", "K"), "fem3d": ("L", "k")}


def _split_kwargs(kind, kwargs):
    gk = {k: v for k, v in kwargs.items() if k in _GEOM_KEYS[kind]}
    return gk


def fem1d_mpi(T=np.float64, Ti=np.int32, backend=None, **kwargs) -> Geometry:
    return native_to_mpi(geom_mod.fem1d(**_split_kwargs("fem1d", kwargs)), Ti=Ti, backend=backend)


def fem2d_mpi(T=np.float64, Ti=np.int32, backend=None, **kwargs) -> Geometry:
    return native_to_mpi(geom_mod.fem2d(**_split_kwargs("fem2d", kwargs)), Ti=Ti, backend=backend)


def fem3d_mpi(T=np.float64, Ti=np.int32, backend=None, **kwargs) -> Geometry:
    return native_to_mpi(geom_mod.fem3d(**_split_kwargs("fem3d", kwargs)), Ti=Ti, backend=backend)


def amgb(g: Geometry, /, **kwargs) -> solver.AMGBSOL:
    """amgb on an HPC-typed geometry (re-exported by the reference, src:752).  Geometry keys that
    femNd_mpi_solve forwards to both calls (src:594-600) are ignored here."""
    for k in ("L", "K", "k", "Ti", "backend", "T"):
        kwargs.pop(k, None)
    be = g.x.backend if isinstance(g.x, HPCMatrix) else backend_cuda()
    g_native = mpi_to_native(g) if isinstance(g.x, HPCMatrix) else g
    sol = solver.amgb(g_native, device=be.index, **kwargs)
    part = g.x.row_partition if isinstance(g.x, HPCMatrix) else None
    z = HPCMatrix(sol.z, be, part) if isinstance(g.x, HPCMatrix) else sol.z
    return solver.AMGBSOL(z, sol.SOL_feasibility, sol.SOL_main, sol.log, g, sol.stats)


def parabolic_solve(g: Geometry, /, **kwargs) -> solver.ParabolicSOL:
    """parabolic_solve on an HPC-typed geometry (reference test/test_parabolic.jl:48): h, t1, p as upstream;
    snapshots come back as HPCMatrix (n x 3: u, s1, s2)."""
    be = g.x.backend if isinstance(g.x, HPCMatrix) else backend_cuda()
    g_native = mpi_to_native(g) if isinstance(g.x, HPCMatrix) else g
    sol = solver.parabolic_solve(g_native, device=be.index, **kwargs)
    if isinstance(g.x, HPCMatrix):
        return solver.ParabolicSOL(g, sol.ts, [HPCMatrix(u, be, g.x.row_partition) for u in sol.u])
    return sol


def fem1d_mpi_solve(T=np.float64, **kwargs):
    return amgb(fem1d_mpi(T, **kwargs), **kwargs)


def fem2d_mpi_solve(T=np.float64, **kwargs):
    return amgb(fem2d_mpi(T, **kwargs), **kwargs)


def fem3d_mpi_solve(T=np.float64, D=None, f=None, g=None, **kwargs):
    """3D defaults as in the reference (src:735-745)."""
    D = DEFAULT_D[3] if D is None else D
    f = DEFAULT_F[3] if f is None else f
    g = DEFAULT_G[3] if g is None else g
    return amgb(fem3d_mpi(T, **kwargs), D=D, f=f, g=g, **kwargs)
