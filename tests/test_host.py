"""CPU-side checks: the C-ABI library loads and exports every symbol include/mgb_b200.h declares;
the symbolic phase (host C++) reproduces the oracle's structural sparsity bit-exactly; numeric entry
points refuse to run without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

import mgb_b200
from mgb_b200 import capi
import mgb_oracle as O

from helpers import problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    header = open(os.path.join(ROOT, "include", "mgb_b200.h")).read()
    declared = set(re.findall(r"\b(mgb_[a-z_0-9A-Z]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(capi.EXPORTS) == declared


def _structural_pattern(pr):
    n = pr["x"].shape[0]
    nD = len(pr["D"])
    y2 = np.abs(np.random.default_rng(0).normal(size=(n, nD * nD))) + 0.1
    Hs = O.hessian_fine(y2, pr["w"], pr["D"])
    H = (pr["R"].T.tocsc() @ Hs @ pr["R"].tocsc()).tocsr()
    H.sort_indices()
    return H


@pytest.mark.parametrize("gen,L,slack,level", [
    ("fem1d", 1, False, None), ("fem1d", 4, False, None), ("fem1d", 4, False, 1), ("fem1d", 3, True, None),
    ("fem2d", 1, False, None), ("fem2d", 3, False, None), ("fem2d", 3, False, 0), ("fem2d", 4, False, 2),
    ("fem2d", 2, True, None), ("fem2d", 3, True, 1)])
@pytest.mark.parametrize("path", [0, capi.PATH_CSR])
def test_symbolic_pattern_bit_exact(gen, L, slack, level, path):
    geom = getattr(mgb_b200, gen)(L)
    pr = problem(geom, slack=slack, level=level)
    plan = capi.Plan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0, slack=slack, force_path=path)
    rp, ci = plan.pattern()
    H = _structural_pattern(pr)
    assert plan.nnzH == H.nnz
    assert np.array_equal(rp, H.indptr) and np.array_equal(ci, H.indices)
    if path == 0:
        assert plan.info["path"] == capi.PATH_ELEMENT
        assert plan.info["nodes_per_element"] == geom.block


def test_fem3d_uses_csr_path_for_now():
    geom = mgb_b200.fem3d(2, k=1)
    pr = problem(geom)
    plan = capi.Plan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0)
    H = _structural_pattern(pr)
    rp, ci = plan.pattern()
    assert np.array_equal(rp, H.indptr) and np.array_equal(ci, H.indices)


def test_numeric_calls_fail_loudly_without_gpu_context():
    pr = problem(mgb_b200.fem1d(2))
    plan = capi.Plan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0)
    with pytest.raises(capi.MgbError, match="no CPU path"):
        plan.assemble_host(pr["s"], None, pr["c"], 1.0, capi.WANT_F0)


def test_bad_inputs_return_errors_not_crashes():
    pr = problem(mgb_b200.fem1d(2))
    with pytest.raises(capi.MgbError):
        capi.Plan(None, pr["D"], pr["R"][:-1], pr["x"], pr["w"], pr["idx"], 1.0)   # R rows != D cols
    with pytest.raises(capi.MgbError):
        capi.Plan(None, pr["D"], pr["R"], pr["x"], pr["w"], [1, 9], 1.0)           # idx outside 0..nD-1
    with pytest.raises(capi.MgbError):
        capi.Plan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 0.5)        # p < 1


def test_one_based_csr_input_accepted():
    """Julia passes 1-based rowptr/colval (index_base=1) zero-copy."""
    import ctypes as C
    pr = problem(mgb_b200.fem1d(3))
    lib = capi.load()
    keep = []

    def one_based(A):
        A = sp.csr_matrix(A); A.sort_indices()
        rp = (A.indptr + 1).astype(np.int32); ci = (A.indices + 1).astype(np.int32); va = A.data.astype(np.float64)
        keep.extend([rp, ci, va])
        return capi._Csr(A.shape[0], A.shape[1], A.nnz, rp.ctypes.data, ci.ctypes.data, va.ctypes.data, 1)

    Ds = (capi._Csr * len(pr["D"]))(*[one_based(d) for d in pr["D"]])
    Rs = one_based(pr["R"])
    bar = capi._Barrier(); bar.kind = 1; bar.nidx = 2; bar.idx[0] = 1; bar.idx[1] = 2; bar.p = 1.0; bar.slack = 0
    x = np.asfortranarray(pr["x"]); w = pr["w"]
    h = C.c_void_p()
    n = pr["x"].shape[0]
    rc = lib.mgb_plan_create(None, n, len(pr["D"]), Ds, C.byref(Rs), 1, x.ctypes.data, w.ctypes.data, C.byref(bar),
                             0, n, 0, C.byref(h))
    assert rc == 0, lib.mgb_last_error()
    info = np.zeros(15, dtype=np.int64)
    lib.mgb_plan_info(h, info.ctypes.data, 15)
    ref = capi.Plan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0)
    assert info[4] == ref.nnzH and info[3] == ref.m
    lib.mgb_plan_destroy(h)


@pytest.mark.parametrize("gen,L,nranks", [("fem1d", 4, 3), ("fem2d", 3, 2), ("fem2d", 3, 4)])
def test_plan_from_hpc_local_blocks_matches_global_plan(gen, L, nranks):
    """mgb_plan_create_local takes a rank's HPCSparseMatrix storage as the reference lays it out (compressed
    column ids + col_indices, 1-based; src/MultiGridBarrierMPI.jl:216-221) and must give the plan that
    mgb_plan_create builds from the global operators restricted to the same rows."""
    from mgb_b200.hpc import Backend, HPCSparseMatrix, uniform_partition
    geom = getattr(mgb_b200, gen)(L)
    pr = problem(geom)
    n = pr["x"].shape[0]
    part = uniform_partition(n, nranks, block=geom.block)
    for rank in range(nranks):
        be = Backend(device="cpu", rank=rank, nranks=nranks)
        blocks = [HPCSparseMatrix(Dk, be, row_partition=part).local_storage() for Dk in pr["D"]]
        lo, hi = int(part[rank] - 1), int(part[rank + 1] - 1)
        assert blocks[0]["row0"] == lo and blocks[0]["nrows_local"] == hi - lo
        # compressed ids really are compressed: dx touches only this block's element columns
        assert blocks[1]["ncols_compressed"] < blocks[1]["ncols_global"]
        loc = capi.Plan.from_local_blocks(None, blocks, pr["R"], n, pr["x"][lo:hi], pr["w"][lo:hi], pr["idx"], 1.0)
        ref = capi.Plan(None, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0, rows=(lo, hi))
        assert loc.info == ref.info
        for a, b in zip(loc.pattern(), ref.pattern()):
            assert np.array_equal(a, b)


def test_plan_from_hpc_local_blocks_rejects_bad_ids():
    from mgb_b200.hpc import Backend, HPCSparseMatrix
    pr = problem(mgb_b200.fem1d(3))
    n = pr["x"].shape[0]
    be = Backend(device="cpu")
    blocks = [HPCSparseMatrix(Dk, be).local_storage() for Dk in pr["D"]]
    bad = dict(blocks[1]); bad["rowval"] = blocks[1]["rowval"].copy(); bad["rowval"][0] = bad["ncols_compressed"] + 1
    with pytest.raises(capi.MgbError, match="compressed column id"):
        capi.Plan.from_local_blocks(None, [blocks[0], bad, blocks[2]], pr["R"], n, pr["x"], pr["w"], pr["idx"], 1.0)
    bad = dict(blocks[1]); bad["col_indices"] = blocks[1]["col_indices"].copy(); bad["col_indices"][0] = 0
    with pytest.raises(capi.MgbError, match="col_indices"):
        capi.Plan.from_local_blocks(None, [blocks[0], bad, blocks[2]], pr["R"], n, pr["x"], pr["w"], pr["idx"], 1.0)


def test_sell_chunk_layout_replays_every_list(tmp_path):
    """the SELL-32-sigma / chunk layout of the CSR path (host C++ in csrc/kernels_csr.cuh) replayed on the CPU exactly
    as the kernels walk it: unchunked lists bit-identical to the sequential sum, chunked ones to rounding, every
    output written exactly once, no lane longer than one chunk"""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = tmp_path / "sell_layout_check"
    csrc = os.path.join(ROOT, "multigridbarriermpi.jl_b200", "csrc")
    cmd = [nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-I", csrc, "-I", os.path.join(ROOT, "include"),
           "-o", str(exe), os.path.join(ROOT, "tests", "native", "sell_layout_check.cu"), os.path.join(csrc, "plan_host.cpp")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0 and "bad 0" in run.stdout, run.stdout[-2000:]


def test_host_sparse_algebra_known_answers(tmp_path):
    """spgemm / transpose of the symbolic phase (csrc/plan_host.cpp) on the literals of the reference's own test
    (test/test_basic_ops.jl:27-66: A*B and A'A of a 3x2 and a 2x3 matrix) + the keep-structural-zeros convention"""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    exe = tmp_path / "host_algebra_check"
    csrc = os.path.join(ROOT, "multigridbarriermpi.jl_b200", "csrc")
    res = subprocess.run([gxx, "-std=c++17", "-O1", "-I", csrc, "-I", os.path.join(ROOT, "include"), "-o", str(exe),
                          os.path.join(ROOT, "tests", "native", "host_algebra_check.cpp"), os.path.join(csrc, "plan_host.cpp")],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert run.returncode == 0 and "bad 0" in run.stdout, run.stdout[-2000:]


def test_hpc_sparse_algebra_conventions():
    """`*`, `'`, `+` on the HPCSparseMatrix mirror with the reference's conventions: the literals of
    test/test_basic_ops.jl:27-66, structural zeros kept by products, exact cancellations dropped by sums
    (test/test_matrix_addition.jl:22-24), and the f2 accumulation of test/test_matrix_addition.jl:38-80 equal to the oracle"""
    from mgb_b200.hpc import Backend, HPCSparseMatrix
    be = Backend(device="cpu")
    A = HPCSparseMatrix(sp.csr_matrix(np.array([[1.0, 0.0], [2.0, 3.0], [0.0, 4.0]])), be)
    B = HPCSparseMatrix(sp.csr_matrix(np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])), be)
    assert np.array_equal((A @ B).host.toarray(), [[1, 2, 3], [14, 19, 24], [16, 20, 24]])
    assert np.array_equal((A.T @ A).host.toarray(), [[5, 6], [6, 25]])
    # a product keeps the structural entry of a cancellation, a sum drops it
    R1 = HPCSparseMatrix(sp.csr_matrix(np.array([[1.0, -1.0]])), be)
    C1 = HPCSparseMatrix(sp.csr_matrix(np.array([[1.0], [1.0]])), be)
    assert (R1 @ C1).host.nnz == 1 and (R1 @ C1).host.data[0] == 0.0
    P = HPCSparseMatrix(sp.csr_matrix(np.array([[1.0, 2.0]])), be)
    M = HPCSparseMatrix(sp.csr_matrix(np.array([[-1.0, 2.0]])), be)
    assert (P + M).host.nnz == 1 and (P + M).host.toarray().tolist() == [[0.0, 4.0]]
    # the f2 accumulation written with these operators = the oracle's Hessian (fem1d L=2, literal weights)
    g = mgb_b200.fem1d(2)
    n = g.x.shape[0]
    D = [HPCSparseMatrix(g.operators["dx"], be), HPCSparseMatrix(g.operators["id"], be)]
    y = np.zeros((n, 4)); y[:, 0] = 0.5; y[:, 1] = 0.1; y[:, 2] = 0.1; y[:, 3] = 0.3
    diag = lambda v: HPCSparseMatrix(sp.diags(g.w * v).tocsr(), be)
    ret = None
    for j in range(2):
        bar = D[j].T @ diag(y[:, j * 2 + j]) @ D[j]
        ret = bar if ret is None else ret + bar
        for k in range(j):
            t1 = D[j].T @ diag(y[:, j * 2 + k]) @ D[k]
            ret = ret + t1 + t1.T
    Ho = O.hessian_fine(y, g.w, [g.operators["dx"], g.operators["id"]])
    assert abs(ret.host - Ho).max() < 1e-13
    Hd = ret.host.copy(); Hd.eliminate_zeros(); Hoz = Ho.copy(); Hoz.eliminate_zeros()
    assert Hd.nnz == Hoz.nnz


def test_amgb_rejects_unknown_keywords_instead_of_dropping_them():
    """a custom convex set `Q` (or any key the GPU path does not implement) must not be ignored silently; the geometry
    constructor's keys that femNd_mpi_solve forwards to both calls (reference src:594-600) are accepted"""
    from mgb_b200 import solver
    with pytest.raises(TypeError, match="unsupported keyword"):
        solver.amgb(mgb_b200.fem1d(2), Q=object())
