"""Host-side mirror of the HPCSparseArrays types the reference dispatches on
(``HPCVector`` / ``HPCMatrix`` / ``HPCSparseMatrix``, reference src/MultiGridBarrierMPI.jl:47-50) with
device-resident storage.  Layouts follow the reference:

* row-block partition, 1-based start offsets of length P+1 (``[1, n+1]`` for one rank;
  tools/profile_solve.jl:24);
* dense local block column-major (Julia ``Matrix``): stored as a (k, n_local) contiguous tensor so
  column c is the contiguous slice the kernels read;
* sparse local block = CSR rows of the owned row range with global column ids
  (src/MultiGridBarrierMPI.jl:216-221).

torch is used for device memory only; arithmetic on the hot path goes through the C ABI.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import scipy.sparse as sp
import torch


def uniform_partition(n: int, nranks: int, block: int = 1) -> np.ndarray:
    """1-based row-block offsets, length nranks+1.  The first ``r = units mod P`` ranks get one extra
    unit (HPCSparseArrays' split rule is not pinned by any reference test - test/test_partitions.jl:36-39
    only prints it - so it lives in this one function).  ``block`` keeps whole broken elements
    (7 / 2 / (k+1)^3 rows) on one rank so that apply_D needs no halo."""
    units = n // block
    assert units * block == n, "n must be a multiple of the element block"
    base, extra = divmod(units, nranks)
    sizes = np.array([(base + (1 if r < extra else 0)) * block for r in range(nranks)], dtype=np.int64)
    return np.concatenate([[1], 1 + np.cumsum(sizes)]).astype(np.int64)


@dataclass
class Backend:
    """HPCBackend{T,Ti,Device,Comm,Solver} (reference src/MultiGridBarrierMPI.jl:86-114)."""
    T: type = np.float64
    Ti: type = np.int32
    device: str = "cuda"     # DeviceCUDA
    comm: str = "nccl"       # one rank per GPU (test/test_2d.jl:14-25)
    solver: str = "host-lu"  # the solve seam stays outside the graft
    index: int = 0
    rank: int = 0
    nranks: int = 1

    @property
    def torch_device(self):
        return torch.device("cuda", self.index) if self.device == "cuda" else torch.device("cpu")


def backend_cuda(index: Optional[int] = None, rank: Optional[int] = None, nranks: Optional[int] = None) -> Backend:
    """Without arguments: the running job's layout - under an initialised torch.distributed job (one process per
    GPU, the reference's `mpiexec -n P` with one rank per GPU, test/test_2d.jl:17-20) the rank / world size of the
    default group and this process's current CUDA device; otherwise one rank on device 0."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        rank = dist.get_rank() if rank is None else rank
        nranks = dist.get_world_size() if nranks is None else nranks
        if index is None:
            index = torch.cuda.current_device() if torch.cuda.is_available() else 0
    return Backend(index=index or 0, rank=rank or 0, nranks=nranks or 1)


class HPCVector:
    def __init__(self, v, backend: Backend, partition: Optional[np.ndarray] = None, local: bool = False):
        """``v``: full host vector (scattered by ``partition``) or, with ``local=True``, this rank's block."""
        self.backend = backend
        if isinstance(v, torch.Tensor) and local:
            n_glob = None
            self.v = v
        else:
            v = np.asarray(v, dtype=np.float64)
            n_glob = v.shape[0]
            self.partition = uniform_partition(n_glob, backend.nranks) if partition is None else np.asarray(partition)
            lo, hi = self.partition[backend.rank] - 1, self.partition[backend.rank + 1] - 1
            self.v = torch.from_numpy(np.ascontiguousarray(v[lo:hi])).to(backend.torch_device)
        if n_glob is None:
            assert partition is not None
            self.partition = np.asarray(partition)
        self.n = int(self.partition[-1] - 1)

    def __len__(self):
        return self.n

    @property
    def shape(self):
        return (self.n,)


class HPCMatrix:
    def __init__(self, A, backend: Backend, row_partition: Optional[np.ndarray] = None, local: bool = False):
        self.backend = backend
        if isinstance(A, torch.Tensor) and local:
            assert row_partition is not None
            self.A = A  # (k, n_local) contiguous == column-major n_local x k
            self.row_partition = np.asarray(row_partition)
        else:
            A = np.asarray(A, dtype=np.float64)
            self.row_partition = uniform_partition(A.shape[0], backend.nranks) if row_partition is None else np.asarray(row_partition)
            lo, hi = self.row_partition[backend.rank] - 1, self.row_partition[backend.rank + 1] - 1
            self.A = torch.from_numpy(np.ascontiguousarray(A[lo:hi].T)).to(backend.torch_device)
        self.n = int(self.row_partition[-1] - 1)
        self.k = int(self.A.shape[0])

    @property
    def shape(self):
        return (self.n, self.k)


GATHER_STATS = dict(calls=0, seconds=0.0, bytes=0)   # collective gathers of row blocks (setup cost of a solve)


class HPCSparseMatrix:
    """Row-partitioned sparse matrix: every rank STORES ONLY ITS OWN ROW BLOCK (global column ids), like the reference
    (src/MultiGridBarrierMPI.jl:216-221).  ``gather()`` / ``.host`` reassemble the whole matrix on every rank with one
    collective (the `SparseMatrixCSC(A)` gather the reference uses in mpi_to_native, src:357-371) - the symbolic phase
    of a level needs R whole, once; the assembly itself never gathers."""

    def __init__(self, A: sp.spmatrix, backend: Backend, row_partition: Optional[np.ndarray] = None,
                 Ti=np.int32, local_block: bool = False, shape=None):
        """``A``: the whole matrix (every rank builds the same native geometry, src:239-240; the rows of other ranks
        are dropped here) or, with ``local_block=True``, this rank's rows only (then ``shape`` = global shape)."""
        self.backend = backend
        A = sp.csr_matrix(A)
        A.sort_indices()
        self.Ti = Ti
        if local_block:
            assert shape is not None and row_partition is not None
            self._shape = tuple(shape)
            self.row_partition = np.asarray(row_partition)
            self._local = A
        else:
            if A.nnz >= np.iinfo(Ti).max:
                raise OverflowError("index type too small for this matrix; pass Ti=np.int64 (src/MultiGridBarrierMPI.jl:233-234)")
            self._shape = A.shape
            self.row_partition = uniform_partition(A.shape[0], backend.nranks) if row_partition is None else np.asarray(row_partition)
            lo, hi = self.row_partition[backend.rank] - 1, self.row_partition[backend.rank + 1] - 1
            self._local = A[lo:hi].copy()
        self.col_partition = uniform_partition(self._shape[1], backend.nranks)
        self._whole = self._local if backend.nranks == 1 else None

    @property
    def shape(self):
        return self._shape

    @property
    def local(self) -> sp.csr_matrix:
        return self._local

    def gather(self) -> sp.csr_matrix:
        """the whole matrix on every rank: collective over the ranks of the running torch.distributed job"""
        if self._whole is None:
            import time
            import torch.distributed as dist
            be = self.backend
            if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() == be.nranks):
                raise RuntimeError("HPCSparseMatrix.gather: this rank holds rows "
                                   f"[{self.row_partition[be.rank]}, {self.row_partition[be.rank + 1]}) only and no "
                                   f"{be.nranks}-rank torch.distributed job is running to gather the others")
            t0 = time.perf_counter()
            blocks = [None] * be.nranks
            loc = self._local
            dist.all_gather_object(blocks, (loc.indptr, loc.indices, loc.data))
            whole = sp.vstack([sp.csr_matrix((d, i, p), shape=(len(p) - 1, self._shape[1])) for p, i, d in blocks], format="csr")
            assert whole.shape == self._shape
            self._whole = whole
            GATHER_STATS["calls"] += 1
            GATHER_STATS["seconds"] += time.perf_counter() - t0
            GATHER_STATS["bytes"] += int(sum(len(d) * 12 + len(p) * 4 for p, i, d in blocks))
        return self._whole

    @property
    def host(self) -> sp.csr_matrix:
        return self.gather()

    @property
    def rowptr(self):
        """1-based local row pointer in the index type Ti (reference src/MultiGridBarrierMPI.jl:364)."""
        return (self.local.indptr + 1).astype(self.Ti)

    @property
    def nnz(self):
        """stored entries of this rank's block"""
        return self._local.nnz

    # ---- the sparse algebra the reference's f2 loop is written in (HPCSparseArrays methods `*`, `'`, `+`:
    # test/test_map_rows_compare.jl:102-123).  Host-side structural operations with Julia's conventions; the
    # hot path never calls them (it replays the frozen pattern on the device) - they exist so that code written
    # against the reference's operator surface runs, and so that the conventions are stated in one place.
    def __matmul__(self, other):
        """A * B keeps structural zeros (Julia spmatmul does not drop cancelled entries)."""
        if isinstance(other, HPCSparseMatrix):
            a, b = self.host.tocsr(), other.host.tocsr()
            pat = (sp.csr_matrix((np.ones(a.nnz), a.indices, a.indptr), shape=a.shape) @
                   sp.csr_matrix((np.ones(b.nnz), b.indices, b.indptr), shape=b.shape)).tocsr()   # structural pattern
            val = (a @ b).tocsr()              # scipy prunes cancelled entries: put the values on the structural pattern
            pat.sort_indices()
            rows = np.repeat(np.arange(pat.shape[0]), np.diff(pat.indptr))
            data = np.asarray(val[rows, pat.indices]).ravel() if pat.nnz else np.zeros(0)
            out = sp.csr_matrix((data, pat.indices, pat.indptr), shape=pat.shape)
            return HPCSparseMatrix(out, self.backend, row_partition=self.row_partition, Ti=self.Ti)
        return NotImplemented

    def __add__(self, other):
        """A + B drops entries that cancel to exactly 0.0, like Julia's SparseMatrixCSC `+` - which makes the
        structure of a sum value-dependent (reference test/test_matrix_addition.jl:22-24)."""
        if isinstance(other, HPCSparseMatrix):
            c = (self.host + other.host).tocsr()
            c.eliminate_zeros()
            return HPCSparseMatrix(c, self.backend, row_partition=self.row_partition, Ti=self.Ti)
        return NotImplemented

    @property
    def T(self):
        """A' (reference test/test_transpose_only.jl): rows of the transpose are partitioned like A's columns"""
        return HPCSparseMatrix(self.host.T.tocsr(), self.backend, row_partition=self.col_partition, Ti=self.Ti)

    def local_storage(self) -> dict:
        """This rank's block in the reference's own field layout (constructor order
        src/MultiGridBarrierMPI.jl:216-221; older names test/test_dump_matrices.jl:62-71): the local rows stored
        as CSC of the transpose - ``colptr`` indexes local rows, ``rowval`` holds COMPRESSED column ids,
        ``col_indices[c]`` is the global column of compressed column c; everything 1-based in ``Ti``."""
        loc = self.local
        lo = int(self.row_partition[self.backend.rank] - 1)
        cols = np.unique(loc.indices)                       # sorted global columns this block touches
        comp = np.searchsorted(cols, loc.indices)
        return dict(nrows_local=loc.shape[0], ncols_compressed=len(cols), ncols_global=loc.shape[1], row0=lo,
                    colptr=(loc.indptr + 1).astype(self.Ti), rowval=(comp + 1).astype(self.Ti),
                    nzval=np.ascontiguousarray(loc.data, dtype=np.float64), col_indices=(cols + 1).astype(self.Ti),
                    index_base=1, has_sorted_rows=True)
