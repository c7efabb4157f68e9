"""ctypes binding of libmgb_b200.so (include/mgb_b200.h).

This is the Python twin of the Julia ``ccall`` shim in ``julia/MGBB200.jl``: the reference's host
language (Julia) is absent from this image, so the C ABI is exercised from here.  There is no CPU
fallback: if the library or a GPU is missing every numeric call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np
import scipy.sparse as sp

from . import build as _build

WANT_F0, WANT_GRAD, WANT_HESS, STORE_DZ = 1, 2, 4, 8
PATH_ELEMENT, PATH_CSR = 1, 2
PLAN_NO_HESSIAN = 16
PLAN_TWO_STAGE = 32
BARRIER_EUCLIDIAN_POWER = 1

EXPORTS = [
    "mgb_last_error", "mgb_version", "mgb_ctx_create", "mgb_ctx_destroy", "mgb_ctx_sync", "mgb_plan_create",
    "mgb_plan_create_local", "mgb_plan_destroy", "mgb_plan_info", "mgb_plan_pattern", "mgb_assemble", "mgb_assemble_host", "mgb_apply_D",
    "mgb_map_barrier", "mgb_all_isfinite", "mgb_reduce", "mgb_diag_scale", "mgb_time_assemble", "mgb_launch_count",
    "mgb_spmat_create", "mgb_spmat_destroy", "mgb_spmat_mv", "mgb_gather_idx", "mgb_scatter_add_idx", "mgb_segsum_idx",
    "mgb_plan_create_rows", "mgb_dist_plan_create", "mgb_dist_info", "mgb_dist_rows", "mgb_dist_pattern", "mgb_dist_window",
    "mgb_dist_export", "mgb_dist_attach", "mgb_dist_attach_local", "mgb_dist_begin", "mgb_dist_end", "mgb_dist_assemble", "mgb_dist_s_publish", "mgb_dist_s_wait", "mgb_dist_assemble_s", "mgb_copy_to_host", "mgb_host_register", "mgb_host_unregister",
    "mgb_graph_begin", "mgb_graph_end", "mgb_graph_launch", "mgb_graph_destroy", "mgb_graph_stats", "mgb_pdl_active",
]


class MgbError(RuntimeError):
    pass


class _Csr(C.Structure):
    _fields_ = [("nrows", C.c_int64), ("ncols", C.c_int64), ("nnz", C.c_int64),
                ("rowptr", C.c_void_p), ("colidx", C.c_void_p), ("vals", C.c_void_p),
                ("index_base", C.c_int32)]


class _HpcBlock(C.Structure):
    """mgb_hpc_block: one rank's HPCSparseMatrix storage in the reference's field layout (src:216-221)."""
    _fields_ = [("nrows_local", C.c_int64), ("ncols_compressed", C.c_int64), ("ncols_global", C.c_int64),
                ("row0", C.c_int64), ("colptr", C.c_void_p), ("rowval", C.c_void_p), ("nzval", C.c_void_p),
                ("col_indices", C.c_void_p), ("index_base", C.c_int32)]


class _Barrier(C.Structure):
    _fields_ = [("kind", C.c_int32), ("nidx", C.c_int32), ("idx", C.c_int32 * 8), ("p", C.c_double),
                ("slack", C.c_int32), ("nidx2", C.c_int32), ("idx2", C.c_int32 * 8), ("p2", C.c_double)]


_lib = None


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """Load the shared library (building it in-tree with nvcc when absent and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if build_if_missing and _build.needs_build():
        try:
            nvcc = _build._nvcc()
        except RuntimeError as exc:   # no compiler on this box: an existing (possibly older) library is all there is
            nvcc = None
            if not os.path.exists(path):
                raise MgbError(f"libmgb_b200.so is not built and cannot be built here: {exc}") from exc
            import warnings
            warnings.warn("libmgb_b200.so is older than its sources and nvcc is not available: loading the stale library")
        if nvcc is not None:
            try:
                _build.build()
            except Exception as exc:   # a failed rebuild must not silently fall back to stale kernels
                raise MgbError(f"libmgb_b200.so is out of date and the rebuild failed: {exc}") from exc
    if not os.path.exists(path):
        raise MgbError("libmgb_b200.so missing: run __graft_entry__.build() (there is no CPU fallback)")
    lib = C.CDLL(path)
    lib.mgb_last_error.restype = C.c_char_p
    lib.mgb_version.restype = C.c_int
    lib.mgb_launch_count.restype = C.c_int64
    lib.mgb_ctx_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.mgb_ctx_destroy.argtypes = [C.c_void_p]
    lib.mgb_ctx_sync.argtypes = [C.c_void_p]
    lib.mgb_plan_create.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(_Csr), C.POINTER(_Csr), C.c_int32,
                                    C.c_void_p, C.c_void_p, C.POINTER(_Barrier), C.c_int64, C.c_int64, C.c_int32,
                                    C.POINTER(C.c_void_p)]
    lib.mgb_plan_create_local.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(_HpcBlock), C.POINTER(_Csr), C.c_int32,
                                          C.c_void_p, C.c_void_p, C.POINTER(_Barrier), C.c_int32, C.POINTER(C.c_void_p)]
    lib.mgb_plan_destroy.argtypes = [C.c_void_p]
    lib.mgb_plan_info.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    lib.mgb_plan_pattern.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mgb_assemble.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int32,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mgb_assemble_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_double,
                                      C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mgb_apply_D.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mgb_map_barrier.argtypes = [C.c_void_p, C.POINTER(_Barrier), C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]
    lib.mgb_all_isfinite.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int32)]
    lib.mgb_reduce.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_double)]
    lib.mgb_diag_scale.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p]
    lib.mgb_time_assemble.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int32,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                      C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.mgb_spmat_create.argtypes = [C.c_void_p, C.POINTER(_Csr), C.POINTER(C.c_void_p)]
    lib.mgb_spmat_destroy.argtypes = [C.c_void_p]
    lib.mgb_spmat_mv.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    lib.mgb_gather_idx.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.mgb_scatter_add_idx.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.mgb_segsum_idx.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.mgb_dist_plan_create.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(_Csr), C.POINTER(_Csr), C.c_int32,
                                         C.c_void_p, C.c_void_p, C.POINTER(_Barrier), C.c_int32, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.mgb_copy_to_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    lib.mgb_host_register.argtypes = [C.c_void_p, C.c_int64]
    lib.mgb_host_unregister.argtypes = [C.c_void_p]
    lib.mgb_graph_begin.argtypes = [C.c_void_p]
    lib.mgb_graph_end.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    lib.mgb_graph_launch.argtypes = [C.c_void_p]
    lib.mgb_graph_destroy.argtypes = [C.c_void_p]
    lib.mgb_graph_stats.argtypes = [C.c_void_p, C.c_void_p]
    lib.mgb_plan_create_rows.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(_Csr), C.POINTER(_Csr), C.c_int32,
                                         C.c_void_p, C.c_void_p, C.POINTER(_Barrier), C.c_int64, C.c_void_p, C.c_int64,
                                         C.c_int64, C.c_int64, C.c_int32, C.POINTER(C.c_void_p)]
    lib.mgb_dist_info.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    lib.mgb_dist_rows.argtypes = [C.c_void_p, C.c_void_p]
    lib.mgb_dist_pattern.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mgb_dist_window.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    lib.mgb_dist_export.argtypes = [C.c_void_p, C.c_void_p]
    lib.mgb_dist_attach.argtypes = [C.c_void_p, C.c_void_p]
    lib.mgb_dist_attach_local.argtypes = [C.c_void_p, C.c_void_p]
    lib.mgb_dist_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int32]
    lib.mgb_dist_end.argtypes = [C.c_void_p, C.c_double, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                 C.POINTER(C.c_void_p)]
    lib.mgb_dist_assemble.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int32,
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    lib.mgb_dist_s_publish.argtypes = [C.c_void_p, C.c_void_p]
    lib.mgb_dist_s_wait.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    lib.mgb_dist_assemble_s.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int32,
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        raise MgbError(load().mgb_last_error().decode("utf-8", "replace"))


def _ptr(a) -> Optional[int]:
    """device pointer of a torch tensor / host pointer of a numpy array / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if isinstance(a, int):
        return a
    return a.data_ptr()


class DeviceView:
    """``count`` float64 items at device address ``ptr`` as a __cuda_array_interface__ object
    (``torch.as_tensor(DeviceView(...), device=...)`` gives a zero-copy tensor over library-owned memory,
    e.g. the exchange window results of DistPlan.end)."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def host_register(a: np.ndarray):
    """page-lock a caller-owned numpy array in place (mgb_host_register)"""
    _check(load().mgb_host_register(a.ctypes.data, a.nbytes))


def host_unregister(a: np.ndarray):
    _check(load().mgb_host_unregister(a.ctypes.data))


def launch_count() -> int:
    return int(load().mgb_launch_count())


def pdl_active() -> bool:
    """programmatic dependent launches in use (False: MGB_NO_PDL=1 or the driver refused the attribute once)"""
    return bool(load().mgb_pdl_active())


class Context:
    """One GPU + one stream (mgb_ctx)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        lib = load()
        h = C.c_void_p()
        _check(lib.mgb_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)))
        self._h = h
        self.device = int(device)

    def sync(self):
        _check(load().mgb_ctx_sync(self._h))

    def all_isfinite(self, v_dev, length: int) -> bool:
        flag = C.c_int32(0)
        _check(load().mgb_all_isfinite(self._h, _ptr(v_dev), int(length), C.byref(flag)))
        return bool(flag.value)

    REDUCE = {"dot": 0, "sum": 1, "norm2sq": 2, "maxabs": 3}

    def reduce(self, op: str, x_dev, length: int, y_dev=None) -> float:
        """dot / sum / squared 2-norm / max-abs of device vectors (mgb_reduce), result on the host"""
        out = C.c_double(0.0)
        _check(load().mgb_reduce(self._h, self.REDUCE[op], _ptr(x_dev), _ptr(y_dev), int(length), None, C.byref(out)))
        return out.value

    def diag_scale(self, w_dev, y_dev, n: int, ld: int, col: int, out_dev):
        _check(load().mgb_diag_scale(self._h, _ptr(w_dev), _ptr(y_dev), int(n), int(ld), int(col), _ptr(out_dev)))

    def map_barrier(self, idx, p: float, slack: bool, nD: int, n: int, Dz_dev, which: int, out_dev):
        """map_rows of the barrier F / F1 / F2 over the rows of Dz (n x nD column-major)."""
        bar = _Barrier()
        bar.kind, bar.nidx, bar.p, bar.slack = BARRIER_EUCLIDIAN_POWER, len(idx), float(p), int(bool(slack))
        for j, v in enumerate(idx):
            bar.idx[j] = int(v)
        _check(load().mgb_map_barrier(self._h, C.byref(bar), int(nD), int(n), _ptr(Dz_dev), int(which), _ptr(out_dev)))

    def gather_idx(self, src_dev, idx_dev, count: int, out_dev):
        _check(load().mgb_gather_idx(self._h, _ptr(src_dev), _ptr(idx_dev), int(count), _ptr(out_dev)))

    def scatter_add_idx(self, src_dev, idx_dev, count: int, dst_dev):
        _check(load().mgb_scatter_add_idx(self._h, _ptr(src_dev), _ptr(idx_dev), int(count), _ptr(dst_dev)))

    def segsum_idx(self, src_dev, ptr_dev, idx_dev, nout: int, dst_dev):
        _check(load().mgb_segsum_idx(self._h, _ptr(src_dev), _ptr(ptr_dev), _ptr(idx_dev), int(nout), _ptr(dst_dev)))

    def to_host(self, src_dev: int, count: int, dtype=np.float64) -> np.ndarray:
        """stream-ordered copy of ``count`` items at device address ``src_dev`` to a new numpy array"""
        out = np.empty(int(count), dtype=dtype)
        _check(load().mgb_copy_to_host(self._h, out.ctypes.data, C.c_void_p(int(src_dev)), out.nbytes))
        return out

    def graph_begin(self):
        """record (instead of execute) the library's launches on this context until graph_end (mgb_graph_begin)"""
        _check(load().mgb_graph_begin(self._h))

    def graph_end(self) -> "Graph":
        h = C.c_void_p()
        _check(load().mgb_graph_end(self._h, C.byref(h)))
        return Graph(h)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            load().mgb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Graph:
    """instantiated CUDA graph of a recorded call sequence (mgb_graph)"""

    def __init__(self, h):
        self._h = h

    def launch(self):
        _check(load().mgb_graph_launch(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            load().mgb_graph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _csr_struct(A: sp.spmatrix, keep: list) -> _Csr:
    A = sp.csr_matrix(A)
    A.sort_indices()
    if A.nnz >= 2 ** 31:
        raise MgbError("operator has >= 2^31 stored entries; Int32 indices (reference default Ti=Int32) overflow")
    rp = np.ascontiguousarray(A.indptr, dtype=np.int32)
    ci = np.ascontiguousarray(A.indices, dtype=np.int32)
    va = np.ascontiguousarray(A.data, dtype=np.float64)
    keep.extend([rp, ci, va])
    return _Csr(A.shape[0], A.shape[1], A.nnz, rp.ctypes.data, ci.ctypes.data, va.ctypes.data, 0)


class Plan:
    """Per-level symbolic plan (mgb_plan): frozen sparsity of R'HR plus replay lists.
    ``ctx=None`` builds a symbolic-only plan (pattern/info queries; numeric calls raise)."""

    INFO = ["path", "n_local", "nD", "m", "nnzH", "elements", "nodes_per_element", "cols_per_var",
            "slots_per_element", "hess_contribs", "grad_contribs", "plan_bytes", "N", "nu", "alg_bytes",
            "hess_stored"]

    def __init__(self, ctx: Context, D: Sequence[sp.spmatrix], R: sp.spmatrix, x: np.ndarray, w: np.ndarray,
                 idx: Sequence[int], p: float, slack: bool = False, rows=None, force_path: int = 0,
                 idx2: Optional[Sequence[int]] = None, p2: float = 2.0):
        lib = load()
        self.ctx = ctx
        n = D[0].shape[0]
        keep: list = []
        Ds = (_Csr * len(D))(*[_csr_struct(d, keep) for d in D])
        Rs = _csr_struct(R, keep)
        bar = _Barrier()
        bar.kind, bar.nidx, bar.p, bar.slack = BARRIER_EUCLIDIAN_POWER, len(idx), float(p), int(bool(slack))
        for j, v in enumerate(idx):
            bar.idx[j] = int(v)
        if idx2:
            bar.nidx2, bar.p2 = len(idx2), float(p2)
            for j, v in enumerate(idx2):
                bar.idx2[j] = int(v)
        x = np.asfortranarray(x, dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        h = C.c_void_p()
        if rows is not None and not (isinstance(rows, tuple) and len(rows) == 2):
            # explicit list of quadrature rows (whole elements, any order): mgb_plan_create_rows, all outputs
            ridx = np.ascontiguousarray(rows, dtype=np.int64)
            _check(lib.mgb_plan_create_rows(ctx._h if ctx is not None else None, n, len(D), Ds, C.byref(Rs), x.shape[1], x.ctypes.data,
                                            w.ctypes.data, C.byref(bar), int(ridx.size), ridx.ctypes.data, int(ridx.size), 0, -1,
                                            int(force_path), C.byref(h)))
        else:
            row0, row1 = (0, n) if rows is None else rows
            _check(lib.mgb_plan_create(ctx._h if ctx is not None else None, n, len(D), Ds, C.byref(Rs), x.shape[1], x.ctypes.data, w.ctypes.data,
                                       C.byref(bar), int(row0), int(row1), int(force_path), C.byref(h)))
        self._finish_init(h)

    @classmethod
    def from_local_blocks(cls, ctx: Optional["Context"], blocks: Sequence[dict], R: sp.spmatrix, n: int, x_local: np.ndarray,
                          w_local: np.ndarray, idx: Sequence[int], p: float, slack: bool = False, force_path: int = 0,
                          idx2: Optional[Sequence[int]] = None, p2: float = 2.0) -> "Plan":
        """Plan from this rank's HPCSparseMatrix storage (mgb_plan_create_local): every entry of ``blocks`` is the
        dict ``HPCSparseMatrix.local_storage()`` returns (reference field names colptr / rowval / nzval /
        col_indices, 1-based)."""
        lib = load()
        self = cls.__new__(cls)
        self.ctx = ctx
        keep: list = []
        arr = (_HpcBlock * len(blocks))()
        for k, b in enumerate(blocks):
            cp = np.ascontiguousarray(b["colptr"], dtype=np.int32)
            rv = np.ascontiguousarray(b["rowval"], dtype=np.int32)
            nz = np.ascontiguousarray(b["nzval"], dtype=np.float64)
            ci = np.ascontiguousarray(b["col_indices"], dtype=np.int32)
            keep.extend([cp, rv, nz, ci])
            arr[k] = _HpcBlock(int(b["nrows_local"]), int(b["ncols_compressed"]), int(b["ncols_global"]), int(b["row0"]),
                               cp.ctypes.data, rv.ctypes.data, nz.ctypes.data, ci.ctypes.data, int(b.get("index_base", 1)))
        Rs = _csr_struct(R, keep)
        bar = _Barrier()
        bar.kind, bar.nidx, bar.p, bar.slack = BARRIER_EUCLIDIAN_POWER, len(idx), float(p), int(bool(slack))
        for j, v in enumerate(idx):
            bar.idx[j] = int(v)
        if idx2:
            bar.nidx2, bar.p2 = len(idx2), float(p2)
            for j, v in enumerate(idx2):
                bar.idx2[j] = int(v)
        x_local = np.asfortranarray(x_local, dtype=np.float64)
        w_local = np.ascontiguousarray(w_local, dtype=np.float64)
        h = C.c_void_p()
        _check(lib.mgb_plan_create_local(ctx._h if ctx is not None else None, int(n), len(blocks), arr, C.byref(Rs),
                                         x_local.shape[1], x_local.ctypes.data, w_local.ctypes.data, C.byref(bar),
                                         int(force_path), C.byref(h)))
        self._finish_init(h)
        return self

    def _finish_init(self, h):
        lib = load()
        self._h = h
        info = np.zeros(16, dtype=np.int64)
        _check(lib.mgb_plan_info(h, info.ctypes.data, 16))
        self.info = dict(zip(self.INFO, (int(v) for v in info)))
        self.n_local, self.nD, self.m, self.nnzH = (self.info[k] for k in ("n_local", "nD", "m", "nnzH"))
        self._pattern = None

    def pattern(self):
        """(rowptr, colidx) of the fixed CSR pattern of R'HR (0-based)."""
        if self._pattern is None:
            rp = np.zeros(self.m + 1, dtype=np.int32)
            ci = np.zeros(max(self.nnzH, 1), dtype=np.int32)
            _check(load().mgb_plan_pattern(self._h, rp.ctypes.data, ci.ctypes.data))
            self._pattern = (rp, ci[: self.nnzH])
        return self._pattern

    def assemble(self, s_dev, Dz0_dev, c_dev, t: float, flags: int, scal_dev=None, grad_dev=None, hval_dev=None,
                 Dz_dev=None):
        _check(load().mgb_assemble(self._h, _ptr(s_dev), _ptr(Dz0_dev), _ptr(c_dev), float(t), int(flags),
                                   _ptr(scal_dev), _ptr(grad_dev), _ptr(hval_dev), _ptr(Dz_dev)))

    def assemble_host(self, s, Dz0, c, t: float, flags: int, upload_inputs: bool = True, out: Optional[dict] = None):
        """Host-buffer call; returns dict(scal, grad, hval, Dz) of numpy arrays (only requested ones).
        ``out``: reuse the arrays of a previous call (e.g. page-locked with host_register)."""
        s = np.ascontiguousarray(s, dtype=np.float64)
        assert s.shape == (self.m,)
        nd = self.n_local * self.nD
        Dz0 = None if Dz0 is None else np.asfortranarray(Dz0, dtype=np.float64)
        c = None if c is None else np.asfortranarray(c, dtype=np.float64)
        if out is not None:
            scal, grad, hval, Dz = out["scal"], out["grad"], out["hval"], out["Dz"]
        else:
            scal = np.zeros(4)
            grad = np.zeros(self.m) if flags & WANT_GRAD else None
            hval = np.zeros(self.nnzH) if flags & WANT_HESS else None
            Dz = np.zeros((self.n_local, self.nD), order="F") if flags & STORE_DZ else None
        _check(load().mgb_assemble_host(self._h, _ptr(s), _ptr(Dz0), _ptr(c), int(upload_inputs), float(t), int(flags),
                                        _ptr(scal), _ptr(grad), _ptr(hval), _ptr(Dz)))
        assert nd >= 0
        return dict(scal=scal, grad=grad, hval=hval, Dz=Dz)

    def graph_stats(self) -> dict:
        v = np.zeros(3, dtype=np.int64)
        _check(load().mgb_graph_stats(self._h, v.ctypes.data))
        return dict(state=int(v[0]), captures=int(v[1]), launches=int(v[2]))

    def apply_D(self, s_dev, Dz0_dev, Dz_dev):
        _check(load().mgb_apply_D(self._h, _ptr(s_dev), _ptr(Dz0_dev), _ptr(Dz_dev)))

    def time_assemble(self, s_dev, Dz0_dev, c_dev, t, flags, scal_dev, grad_dev, hval_dev, reps: int,
                      flush_l2: bool, split: bool = True):
        tot, tel, tga = C.c_float(0), C.c_float(0), C.c_float(0)
        _check(load().mgb_time_assemble(self._h, _ptr(s_dev), _ptr(Dz0_dev), _ptr(c_dev), float(t), int(flags),
                                        _ptr(scal_dev), _ptr(grad_dev), _ptr(hval_dev), int(reps), int(flush_l2),
                                        C.byref(tot), C.byref(tel) if split else None, C.byref(tga) if split else None))
        return tot.value, tel.value, tga.value

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            load().mgb_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DistPlan(Plan):
    """Sharded plan, owner-computes (mgb_dist_*): one instance per rank.

    ``row_part`` / ``out_part``: 0-based offsets (length nranks+1) of the quadrature rows (whole elements) and of
    the unknowns.  The plan evaluates this rank's block of quadrature rows plus the halo elements touching its
    output rows (``rows``: their global ids, in the order of the Dz0 / c blocks the numeric calls take) and
    completes the owned rows of R'HR / block of the gradient locally; only the objective scalars cross the ranks.
    ``ctx=None`` builds the symbolic part only (CPU tests of the host logic)."""

    DINFO = ["rank", "nranks", "n_own_h", "n_own_g", "own0", "own1", "n_rows", "n_primary", "elements", "_9", "_10",
             "window_words", "epoch", "err", "_14", "_15"]

    def __init__(self, ctx: Optional[Context], D, R, x, w, idx, p: float, rank: int, nranks: int, row_part, out_part,
                 slack: bool = False, idx2: Optional[Sequence[int]] = None, p2: float = 2.0):
        lib = load()
        self.ctx = ctx
        n = D[0].shape[0]
        keep: list = []
        Ds = (_Csr * len(D))(*[_csr_struct(d, keep) for d in D])
        Rs = _csr_struct(R, keep)
        bar = _Barrier()
        bar.kind, bar.nidx, bar.p, bar.slack = BARRIER_EUCLIDIAN_POWER, len(idx), float(p), int(bool(slack))
        for j, v in enumerate(idx):
            bar.idx[j] = int(v)
        if idx2:
            bar.nidx2, bar.p2 = len(idx2), float(p2)
            for j, v in enumerate(idx2):
                bar.idx2[j] = int(v)
        x = np.asfortranarray(x, dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        rp = np.ascontiguousarray(row_part, dtype=np.int64)
        op = np.ascontiguousarray(out_part, dtype=np.int64)
        assert rp.size == nranks + 1 and op.size == nranks + 1
        h = C.c_void_p()
        _check(lib.mgb_dist_plan_create(ctx._h if ctx is not None else None, n, len(D), Ds, C.byref(Rs), x.shape[1],
                                        x.ctypes.data, w.ctypes.data, C.byref(bar), int(rank), int(nranks),
                                        rp.ctypes.data, op.ctypes.data, C.byref(h)))
        self._finish_init(h)
        self.rank, self.nranks = int(rank), int(nranks)
        self.dinfo = self.dist_info()
        self.rows = np.zeros(self.dinfo["n_rows"], dtype=np.int64)
        _check(lib.mgb_dist_rows(h, self.rows.ctypes.data))

    def dist_info(self) -> dict:
        v = np.zeros(16, dtype=np.int64)
        _check(load().mgb_dist_info(self._h, v.ctypes.data, 16))
        return {k: int(a) for k, a in zip(self.DINFO, v) if not k.startswith("_")}

    def own_pattern(self):
        """(rowptr, colidx) of the owned rows of R'HR (rowptr relative to the block, global column ids)."""
        if self._pattern is None:
            d = self.dinfo
            rp = np.zeros(d["own1"] - d["own0"] + 1, dtype=np.int32)
            ci = np.zeros(max(d["n_own_h"], 1), dtype=np.int32)
            _check(load().mgb_dist_pattern(self._h, rp.ctypes.data, ci.ctypes.data))
            self._pattern = (rp, ci[: d["n_own_h"]])
        return self._pattern

    def pattern(self):
        return self.own_pattern()

    def window(self):
        p, b = C.c_void_p(), C.c_int64(0)
        _check(load().mgb_dist_window(self._h, C.byref(p), C.byref(b)))
        return int(p.value), int(b.value)

    def export_handle(self) -> bytes:
        buf = (C.c_ubyte * 64)()
        _check(load().mgb_dist_export(self._h, buf))
        return bytes(buf)

    def attach(self, handles: Sequence[bytes]):
        assert len(handles) == self.nranks and all(len(h) == 64 for h in handles)
        buf = (C.c_ubyte * (64 * self.nranks)).from_buffer_copy(b"".join(handles))
        _check(load().mgb_dist_attach(self._h, buf))

    def attach_local(self, windows: Sequence[int]):
        arr = (C.c_void_p * self.nranks)(*[C.c_void_p(int(w)) for w in windows])
        _check(load().mgb_dist_attach_local(self._h, arr))

    def begin(self, s_dev, Dz0_dev, c_dev, t: float, flags: int):
        _check(load().mgb_dist_begin(self._h, _ptr(s_dev), _ptr(Dz0_dev), _ptr(c_dev), float(t), int(flags)))

    def end(self, t: float, flags: int):
        """-> device pointers (hval_own, grad_own, scal), library-owned"""
        hp, gp, sp_ = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(load().mgb_dist_end(self._h, float(t), int(flags), C.byref(hp), C.byref(gp), C.byref(sp_)))
        return int(hp.value), int(gp.value), int(sp_.value)

    def s_publish(self, s_own_dev):
        """this rank's block of a row-distributed Newton unknown -> every rank's copy of the whole vector"""
        _check(load().mgb_dist_s_publish(self._h, _ptr(s_own_dev)))

    def s_wait(self) -> int:
        """enqueue the wait for every rank's block; device pointer of the gathered vector (m doubles, library-owned)"""
        p = C.c_void_p()
        _check(load().mgb_dist_s_wait(self._h, C.byref(p)))
        return int(p.value)

    def dist_assemble_s(self, s_own_dev, Dz0_dev, c_dev, t: float, flags: int):
        """mgb_dist_assemble_s: all-gather of the distributed unknown inside the library, then the sharded assembly"""
        hp, gp, sp_ = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(load().mgb_dist_assemble_s(self._h, _ptr(s_own_dev), _ptr(Dz0_dev), _ptr(c_dev), float(t), int(flags),
                                          C.byref(hp), C.byref(gp), C.byref(sp_)))
        return int(hp.value), int(gp.value), int(sp_.value)

    def dist_assemble(self, s_dev, Dz0_dev, c_dev, t: float, flags: int):
        hp, gp, sp_ = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(load().mgb_dist_assemble(self._h, _ptr(s_dev), _ptr(Dz0_dev), _ptr(c_dev), float(t), int(flags),
                                        C.byref(hp), C.byref(gp), C.byref(sp_)))
        return int(hp.value), int(gp.value), int(sp_.value)


class SpMat:
    """Device CSR matrix with y = alpha*op(A) x + beta*y0 (mgb_spmat)."""

    def __init__(self, ctx: Context, A: sp.spmatrix):
        keep: list = []
        cs = _csr_struct(A, keep)
        h = C.c_void_p()
        _check(load().mgb_spmat_create(ctx._h, C.byref(cs), C.byref(h)))
        self._h = h
        self.shape = A.shape
        self.ctx = ctx

    def mv(self, x_dev, y_dev, trans: bool = False, alpha: float = 1.0, beta: float = 0.0, y0_dev=None):
        _check(load().mgb_spmat_mv(self._h, int(trans), float(alpha), _ptr(x_dev), float(beta), _ptr(y0_dev),
                                   _ptr(y_dev)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            load().mgb_spmat_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
