// C ABI of libmgb_b200.so (see include/mgb_b200.h).  Host glue only: contexts, plans, uploads,
// kernel dispatch.  No torch types, no CPU numeric fallback: every numeric entry point launches
// CUDA kernels or fails.
#include "../../include/mgb_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "kernels_csr.cuh"
#include "kernels_dense.cuh"
#include "kernels_dist.cuh"
#include "launch.h"
#include "plan_host.h"

namespace {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};
// programmatic dependent launches (PDL) are used until a launch with the attribute is refused once; MGB_NO_PDL=1
// disables them for A/B runs.  mgb_pdl_active() reports the current state (no silent downgrade).
bool g_pdl_ok = getenv("MGB_NO_PDL") == nullptr;

int fail(const std::string& msg) {
    g_err = msg;
    return 1;
}

#define CUDA_OK(expr)                                                                        \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
            throw std::runtime_error(std::string(#expr) + ": " + cudaGetErrorString(_e));    \
    } while (0)

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    void alloc(size_t count) {
        if (p) cudaFree(p), p = nullptr;
        n = count;
        if (count) CUDA_OK(cudaMalloc(&p, count * sizeof(T)));
    }
    void upload(const std::vector<T>& h, cudaStream_t st) {
        alloc(h.size());
        if (!h.empty()) CUDA_OK(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    }
    size_t bytes() const { return n * sizeof(T); }
};

// Coarse multigrid levels have few outputs with very long contribution lists (level 0 of fem2d L=8: 25 entries fed
// by all 32,768 elements).  One warp per output walks such a list in hundreds of dependent steps and sets the kernel
// time alone; instead every list is cut into chunks of GATHER_CHUNK contributions (one warp each, stage 1) whose
// partial sums a second launch adds per output in list order (stage 2).  Deterministic; the chunking depends on the
// plan only.
struct ChunkedList {
    DevBuf<int64_t> kptr;   // nchunks + 1: contribution range of every chunk
    DevBuf<int64_t> pptr;   // nout + 1: partial range of every output
    DevBuf<double> part;    // nchunks
    int64_t nchunks = 0;
    void build(const std::vector<int64_t>& cptr, int64_t chunk, cudaStream_t st) {
        const int64_t nout = (int64_t)cptr.size() - 1;
        std::vector<int64_t> k(1, 0), pp((size_t)nout + 1, 0);
        for (int64_t e = 0; e < nout; ++e) {
            for (int64_t c0 = cptr[e]; c0 < cptr[e + 1]; c0 += chunk) k.push_back(std::min(cptr[e + 1], c0 + chunk));
            pp[(size_t)e + 1] = (int64_t)k.size() - 1;
        }
        nchunks = (int64_t)k.size() - 1;
        kptr.upload(k, st); pptr.upload(pp, st);
        part.alloc((size_t)std::max<int64_t>(nchunks, 1));
        CUDA_OK(cudaStreamSynchronize(st));   // k / pp die with this scope
    }
    size_t bytes() const { return kptr.bytes() + pptr.bytes() + part.bytes(); }
};

static int64_t gather_chunk() {
    const char* e = std::getenv("MGB_GATHER_CHUNK");
    const long v = e ? std::atol(e) : 128;
    return v > 0 ? v : 0;   // 0: one warp per output (no chunking)
}

}  // namespace

struct mgb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    bool capturing = false;   // between mgb_graph_begin and mgb_graph_end: launches are recorded, not executed
    DevBuf<double> flush;  // L2 flush scratch (lazy)
    DevBuf<int> flag;
    DevBuf<double> red;       // reduce_kernel: REDUCE_BLOCKS partials + result (lazy)
    DevBuf<unsigned> ticket;
};

struct mgb_plan {
    mgb_ctx* ctx = nullptr;
    int path = 0;
    int64_t n = 0, nloc = 0, N = 0, m = 0, nnzH = 0;
    int64_t out0 = 0, m_out = 0;   // output rows [out0, out0 + m_out) of the m unknowns (all of them unless sharded)
    int64_t n_primary = 0;         // leading local quadrature rows that count in the scalars (nloc unless sharded)
    int ND = 0, NU = 0, dim = 0;
    mgb::BarrierDesc bar;
    std::vector<int32_t> h_rowptr, h_colidx;
    int64_t alg_bytes = 0;
    size_t dev_bytes = 0;
    // ---- element path
    mgb::ElementPlan ep;  // host arrays are released after upload except the pattern
    DevBuf<int32_t> d_lcols, d_hcidx, d_gcidx;
    DevBuf<int64_t> d_hcptr, d_gcptr, d_hlptr;
    // coarse levels: contribution lists cut into chunks (one warp each) + partial sums (see ChunkedList)
    ChunkedList ck_h, ck_g;
    DevBuf<int2> d_hsrc2;
    DevBuf<int32_t> d_hlidx, d_hlt;
    int64_t n_long = 0;
    DevBuf<double> d_prec, d_w, d_sel, d_rel, d_part, d_scal_tmp;
    // dense path (coarse levels of large elements): group dofs, chunk table, dense operator rows
    DevBuf<int32_t> d_dgdof;
    DevBuf<int64_t> d_dchunk;
    DevBuf<double> d_drows;
    int64_t nblocks_elem = 0, n_hcontrib = 0, n_gcontrib = 0, n_hstored = 0;
    bool has_hessian = true;
    bool long_lists = false;
    // ---- csr path
    mgb::CsrDev csr;
    // ---- host staging for the *_host entry point
    DevBuf<double> st_s, st_dz0, st_c, st_scal, st_grad, st_hval, st_dz;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // ---- multi-GPU, owner-computes sharding (mgb_dist_*)
    struct Dist {
        int rank = 0, nranks = 1;
        std::vector<int64_t> rows;        // global quadrature rows of this plan, in plan order (primary block first)
        DevBuf<double> hval, grad, scal;  // owned rows of R'HR (mgb_dist_pattern order), owned gradient block, scalars
        DevBuf<int> err;
        void* window = nullptr;           // scalar exchange words [2 parities][DIST_LL_RANKS][8]; cudaMalloc'ed, IPC-exported
        size_t window_bytes = 0;
        void* peer[mgb::DIST_LL_RANKS] = {nullptr};
        bool peer_ipc[mgb::DIST_LL_RANKS] = {false};
        bool attached = false;
        unsigned epoch = 0;
        bool finish_pending = false;
        double timeout_s = 30.0;
        // all-gather of a row-distributed Newton unknown (mgb_dist_s_publish / _wait / mgb_dist_assemble_s): the window
        // is [scalar words | 16 flags | whole vector, parity 0 | whole vector, parity 1]
        size_t off_flags = 0, off_full = 0;
        unsigned long long s_epoch = 0;
        DevBuf<unsigned int> counter;
        ~Dist() {
            for (int p = 0; p < mgb::DIST_LL_RANKS; ++p)
                if (peer_ipc[p] && peer[p]) cudaIpcCloseMemHandle(peer[p]);
            if (window) cudaFree(window);
        }
    };
    std::unique_ptr<Dist> dist;
    // ---- CUDA-graph cache of mgb_assemble: one instantiated graph per distinct argument tuple (the Newton loop calls
    // with a handful: trial / accepted iterate x objective-only / full, t changing once per central-path step)
    struct GraphEntry {
        const void* key[8] = {nullptr};
        double t = 0.0;
        int flags = 0;
        cudaGraphExec_t exec = nullptr;
        int kernels = 0;
        uint64_t last_use = 0;
    };
    std::vector<GraphEntry> graphs;
    uint64_t graph_clock = 0;
    int graph_state = 0;   // 0 untested, 1 in use, -1 disabled (env MGB_GRAPH=0, legacy default stream, or capture refused)
    int64_t graph_hits = 0, graph_captures = 0;
};

struct mgb_graph {
    mgb_ctx* ctx = nullptr;
    cudaGraphExec_t exec = nullptr;
    int kernels = 0;
};

struct mgb_spmat {
    mgb_ctx* ctx = nullptr;
    int64_t nrows = 0, ncols = 0;
    DevBuf<int64_t> ptr, tptr;
    DevBuf<int32_t> idx, tidx;
    DevBuf<double> val, tval;
    ChunkedList ck, tck;   // built when an orientation has rows of >= 1024 entries (thread-per-row would crawl)
};

namespace {

// sort columns inside each row (HPCSparseMatrix has_sorted_rows may be false)
void sort_rows(mgb::HostCSR& H) {
    std::vector<std::pair<int32_t, double>> tmp;
    for (int64_t i = 0; i < H.nrows; ++i) {
        bool sorted = true;
        for (int64_t p = H.ptr[i] + 1; p < H.ptr[i + 1]; ++p)
            if (H.idx[p] < H.idx[p - 1]) { sorted = false; break; }
        if (sorted) continue;
        tmp.clear();
        for (int64_t p = H.ptr[i]; p < H.ptr[i + 1]; ++p) tmp.emplace_back(H.idx[p], H.val[p]);
        std::sort(tmp.begin(), tmp.end());
        for (int64_t p = H.ptr[i]; p < H.ptr[i + 1]; ++p) {
            H.idx[p] = tmp[p - H.ptr[i]].first;
            H.val[p] = tmp[p - H.ptr[i]].second;
        }
    }
}

mgb::HostCSR to_host_csr(const mgb_csr& A, int64_t row0, int64_t row1) {
    mgb::HostCSR H;
    const int base = A.index_base;
    if (row0 < 0 || row1 > A.nrows || row0 > row1) throw std::runtime_error("row range outside matrix");
    H.nrows = row1 - row0;
    H.ncols = A.ncols;
    H.ptr.resize(H.nrows + 1);
    const int64_t p0 = A.rowptr[row0] - base;
    for (int64_t i = 0; i <= H.nrows; ++i) H.ptr[i] = (int64_t)A.rowptr[row0 + i] - base - p0;
    const int64_t cnt = H.ptr[H.nrows];
    H.idx.resize(cnt);
    H.val.resize(cnt);
    for (int64_t p = 0; p < cnt; ++p) {
        const int32_t j = A.colidx[p0 + p] - base;
        if (j < 0 || j >= A.ncols) throw std::runtime_error("column index outside matrix");
        H.idx[p] = j;
        H.val[p] = A.vals[p0 + p];
    }
    sort_rows(H);
    return H;
}

// rows `rows[0..cnt)` of A (global 0-based row ids, any order) as a host CSR block
mgb::HostCSR to_host_csr_rows(const mgb_csr& A, const int64_t* rows, int64_t cnt) {
    mgb::HostCSR H;
    const int base = A.index_base;
    H.nrows = cnt; H.ncols = A.ncols;
    H.ptr.assign(cnt + 1, 0);
    for (int64_t i = 0; i < cnt; ++i) {
        if (rows[i] < 0 || rows[i] >= A.nrows) throw std::runtime_error("row id outside matrix");
        H.ptr[i + 1] = H.ptr[i] + (A.rowptr[rows[i] + 1] - A.rowptr[rows[i]]);
    }
    H.idx.resize(H.ptr[cnt]); H.val.resize(H.ptr[cnt]);
    for (int64_t i = 0; i < cnt; ++i) {
        const int64_t p0 = A.rowptr[rows[i]] - base;
        for (int64_t q = H.ptr[i]; q < H.ptr[i + 1]; ++q) {
            const int32_t j = A.colidx[p0 + (q - H.ptr[i])] - base;
            if (j < 0 || j >= A.ncols) throw std::runtime_error("column index outside matrix");
            H.idx[q] = j; H.val[q] = A.vals[p0 + (q - H.ptr[i])];
        }
    }
    sort_rows(H);
    return H;
}

// unfused-minimum algorithmic bytes of one assembly (gradient + restricted Hessian)
int64_t algorithmic_bytes(int64_t n, int64_t N, int nD, int dim, int64_t nnzD, int64_t nnzH_fine, int64_t nnzR,
                          int64_t nnzH) {
    const int64_t y2u = (int64_t)nD * (nD + 1) / 2;
    int64_t b = 0;
    b += N * 8 + nnzD * 12 + (int64_t)nD * (n + 1) * 4 + n * nD * 8;  // apply_D
    b += n * (int64_t)(nD + dim + 1 + nD) * 8;                         // barrier reads Dz, x, w, c
    b += n * (int64_t)nD * 8;                                          // y1 write
    b += n * y2u * 8;                                                  // y2 unique write
    b += nnzD * 12 + n * (int64_t)nD * 8 + N * 8;                      // gradient
    b += n * y2u * 8 + nnzD * 8 + nnzH_fine * 8;                       // Hessian numeric
    b += nnzH_fine * 8 + nnzR * 12 * 2 + nnzH * 8;                     // restriction
    return b;
}

bool elem_supported(int B, int dim) { return mgb::element_supported(B, dim); }

// Every index the numeric kernels will dereference comes from these frozen lists: check all of them once, on the
// host, at plan creation (compute-sanitizer is not available on the target pool, so the bounds are enforced here).
void validate_element_plan(const mgb::ElementPlan& ep) {
    const int64_t nrec = (ep.E + ep.agg - 1) / ep.agg;
    const int64_t nsel = nrec * (int64_t)ep.lay.NS, nrel = nrec * (int64_t)ep.NU * ep.LPE;
    auto bad = [](const char* what) { throw std::runtime_error(std::string("internal: element plan validation failed: ") + what); };
    for (int32_t a : ep.lcols) if (a < -1 || a >= ep.m) bad("dof id outside -1..m-1");
    for (int32_t sl : ep.h_cidx) if (sl < 0 || sl >= nsel) bad("Hessian contribution slot outside the record buffer");
    for (int32_t sl : ep.g_cidx) if (sl < 0 || sl >= nrel) bad("gradient contribution slot outside the record buffer");
    if ((int64_t)ep.h_rowptr.size() != ep.m_out + 1 || (int64_t)ep.g_cptr.size() != ep.m_out + 1) bad("row pointer length");
    if (!ep.h_cptr.empty() && (ep.h_cptr.front() != 0 || ep.h_cptr.back() != (int64_t)ep.h_cidx.size())) bad("contribution pointers");
    for (size_t t = 1; t < ep.h_cptr.size(); ++t) if (ep.h_cptr[t] < ep.h_cptr[t - 1]) bad("contribution pointers not monotone");
    for (int64_t a = 0; a < ep.m_out; ++a) {
        if (ep.h_rowptr[a + 1] < ep.h_rowptr[a]) bad("row pointers not monotone");
        for (int32_t q = ep.h_rowptr[a]; q < ep.h_rowptr[a + 1]; ++q) {
            if (ep.h_colidx[q] < 0 || ep.h_colidx[q] >= ep.m) bad("column id outside 0..m-1");
            if (q > ep.h_rowptr[a] && ep.h_colidx[q] <= ep.h_colidx[q - 1]) bad("columns not strictly increasing in a row");
        }
    }
    if (ep.g_cptr.back() != (int64_t)ep.g_cidx.size()) bad("gradient pointers");
    for (double v : ep.prec) if (!(v == v)) bad("NaN in an operator record");
}

void launch_dense(const mgb_plan* pl, const mgb::ElemParams& E, int flags) {
    mgb::DenseParams P{};
    P.nchunks = pl->ep.d_nchunks; P.nloc = pl->ep.nloc;
    P.chunk = pl->d_dchunk.p; P.gdof = pl->d_dgdof.p; P.rows = pl->d_drows.p; P.w = pl->d_w.p;
    P.s = E.s; P.Dz0 = E.Dz0; P.c = E.c; P.t = E.t; P.p = E.p;
    P.sel = E.sel; P.rel = E.rel; P.part = E.part; P.Dz = E.Dz;
    static bool opted = false;
    if (!opted) {
        CUDA_OK(cudaFuncSetAttribute(mgb::dense_element_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mgb::D_SMEM));
        CUDA_OK(cudaFuncSetAttribute(mgb::dense_element_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, mgb::D_SMEM));
        CUDA_OK(cudaFuncSetAttribute(mgb::dense_element_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, mgb::D_SMEM));
        CUDA_OK(cudaFuncSetAttribute(mgb::dense_element_kernel<15>, cudaFuncAttributeMaxDynamicSharedMemorySize, mgb::D_SMEM));
        opted = true;
    }
    const dim3 g((unsigned)P.nchunks), b(mgb::D_THREADS);
    cudaStream_t st = pl->ctx->stream;
    switch (mgb::canonical_flags(flags)) {
        case 1: mgb::dense_element_kernel<1><<<g, b, mgb::D_SMEM, st>>>(P); break;
        case 7: mgb::dense_element_kernel<7><<<g, b, mgb::D_SMEM, st>>>(P); break;
        case 8: mgb::dense_element_kernel<8><<<g, b, mgb::D_SMEM, st>>>(P); break;
        case 15: mgb::dense_element_kernel<15><<<g, b, mgb::D_SMEM, st>>>(P); break;
        default: throw std::runtime_error("assemble: empty flags");
    }
    g_launches++;
    CUDA_OK(cudaGetLastError());
}

void launch_elem(const mgb_plan* pl, const mgb::ElemParams& P, int flags) {
    const auto& ep = pl->ep;
    if (ep.dense) { launch_dense(pl, P, flags); return; }
    mgb::launch_element(ep.B, ep.dim, ep.mode, ep.fine, P, flags, pl->nblocks_elem, pl->ctx->stream);
    g_launches++;
    CUDA_OK(cudaGetLastError());
}

// Launch as a programmatic dependent of the previous kernel in the stream (PDL): the kernel may start while
// the element kernel's last wave drains and blocks at griddepcontrol.wait until its records are complete.
// Falls back to a plain launch when the attribute is rejected (MGB_NO_PDL=1 disables it for A/B runs).

template <class Params>
void launch_dependent(void (*kernel)(Params), unsigned grid, unsigned block, cudaStream_t st, const Params& params, bool allow_pdl) {
    if (allow_pdl && g_pdl_ok) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (cudaLaunchKernelEx(&cfg, kernel, params) == cudaSuccess) return;
        cudaGetLastError();
        g_pdl_ok = false;
    }
    kernel<<<grid, block, 0, st>>>(params);
}

mgb::ElemParams make_elem_params(mgb_plan* pl, const double* s, const double* Dz0, const double* c, double t, double* Dz) {
    const auto& ep = pl->ep;
    mgb::ElemParams P{};
    P.E = ep.E; P.nloc = ep.nloc; P.agg = ep.agg; P.Eprim = pl->n_primary / std::max(ep.B, 1);
    P.lcols = pl->d_lcols.p; P.prec = pl->d_prec.p;
    P.s = s; P.Dz0 = Dz0; P.c = c; P.t = t; P.p = pl->bar.p; P.p2 = pl->bar.p2;
    P.sel = pl->d_sel.p; P.rel = pl->d_rel.p; P.part = pl->d_part.p; P.Dz = Dz;
    P.off_uu = ep.lay.off_uu; P.off_us = ep.lay.off_us; P.off_ss = ep.lay.off_ss;
    P.off_ut = ep.lay.off_ut; P.off_st = ep.lay.off_st; P.off_tt = ep.lay.off_tt; P.NS = ep.lay.NS;
    P.mma = ep.mma ? 1 : 0;
    return P;
}

// two-wide ELL + long-list replay parameters of the thread-per-entry gather (fine levels)
mgb::GatherParams make_gather_params(mgb_plan* pl, int flags, double t, double* scal, double* grad, double* hval) {
    mgb::GatherParams G{};
    G.nnzH = pl->nnzH; G.m = pl->m_out;
    G.h_src2 = pl->d_hsrc2.p; G.h_lptr = pl->d_hlptr.p; G.h_lidx = pl->d_hlidx.p; G.h_lt = pl->d_hlt.p;
    G.g_cptr = pl->d_gcptr.p; G.g_cidx = pl->d_gcidx.p;
    G.sel = pl->d_sel.p; G.rel = pl->d_rel.p; G.hval = hval; G.grad = grad;
    G.part = pl->d_part.p; G.nparts = pl->nblocks_elem; G.scal = scal; G.t = t;
    G.want_h = (flags & MGB_WANT_HESS) ? 1 : 0;
    G.want_g = (flags & MGB_WANT_GRAD) ? 1 : 0;
    return G;
}

void size_gather_grid(mgb_plan* pl, mgb::GatherParams& G) {
    G.nblk_h = G.want_h ? (pl->nnzH + 256 * mgb::GATHER_UNROLL - 1) / (256 * mgb::GATHER_UNROLL) : 0;
    G.nblk_g = G.want_g ? (pl->m_out + 255) / 256 : 0;
    G.n_long = G.want_h ? pl->n_long : 0;
    G.nblk_l = (G.n_long + 255) / 256;
}

void assemble_element(mgb_plan* pl, const double* s, const double* Dz0, const double* c, double t, int flags,
                      double* scal, double* grad, double* hval, double* Dz, cudaEvent_t mid = nullptr,
                      const mgb::DistScal* dscal = nullptr) {
    cudaStream_t st = pl->ctx->stream;
    const auto& ep = pl->ep;
    mgb::ElemParams P = make_elem_params(pl, s, Dz0, c, t, Dz);
    if ((flags & MGB_STORE_DZ) && !Dz) throw std::runtime_error("MGB_STORE_DZ without Dz buffer");
    if ((flags & MGB_WANT_GRAD) && !grad) throw std::runtime_error("MGB_WANT_GRAD without grad buffer");
    if ((flags & MGB_WANT_HESS) && !hval) throw std::runtime_error("MGB_WANT_HESS without hval buffer");
    launch_elem(pl, P, flags);
    if (mid) CUDA_OK(cudaEventRecord(mid, st));

    mgb::GatherParams G = make_gather_params(pl, flags, t, scal ? scal : pl->d_scal_tmp.p, grad, hval);
    if (dscal) G.dist = *dscal;
    if (pl->long_lists) {
        if (dscal) throw std::runtime_error("sharded plans use the thread-per-entry gather (fine levels)");
        // coarse levels: few output entries with long lists.  Stage 1: one warp (or 8 lanes) per output, or per chunk
        // of a chunked list; stage 2 (only with chunked lists): the partial sums of every output.  The scalar fold
        // rides in the last launch.
        auto stage1 = [&](ChunkedList& ck, int64_t nout, int64_t ncontrib, const int64_t* cptr, const int32_t* cidx, const double* src,
                          double* dst) {
            mgb::WarpList w{};
            if (ck.nchunks > 0) { w.nout = ck.nchunks; w.ptr = ck.kptr.p; w.idx = cidx; w.src = src; w.dst = ck.part.p; w.lpe = 32; }
            else { w.nout = nout; w.ptr = cptr; w.idx = cidx; w.src = src; w.dst = dst; w.lpe = (ncontrib < 48 * nout) ? 8 : 32; }
            w.nblk = (w.nout * w.lpe + 255) / 256;
            return w;
        };
        auto stage2 = [&](ChunkedList& ck, int64_t nout, double* dst) {
            mgb::WarpList w{};
            if (ck.nchunks > 0) { w.nout = nout; w.ptr = ck.pptr.p; w.idx = nullptr; w.src = ck.part.p; w.dst = dst; w.lpe = 32; w.nblk = (nout * 32 + 255) / 256; }
            return w;
        };
        mgb::WarpGatherParams W1{}, W2{};
        if (G.want_h) { W1.a = stage1(pl->ck_h, pl->nnzH, pl->n_hcontrib, pl->d_hcptr.p, pl->d_hcidx.p, G.sel, hval); W2.a = stage2(pl->ck_h, pl->nnzH, hval); }
        if (G.want_g) { W1.b = stage1(pl->ck_g, pl->m_out, pl->n_gcontrib, G.g_cptr, G.g_cidx, G.rel, grad); W2.b = stage2(pl->ck_g, pl->m_out, grad); }
        const bool two = W2.a.nblk + W2.b.nblk > 0;
        mgb::WarpGatherParams& last = two ? W2 : W1;
        last.part = G.part; last.nparts = G.nparts; last.t = G.t; last.scal = G.scal;
        launch_dependent(mgb::warp_gather_kernel, (unsigned)(W1.a.nblk + W1.b.nblk + (two ? 0 : 1)), 256u, st, W1, /*allow_pdl=*/mid == nullptr);
        g_launches++;
        if (two) {
            launch_dependent(mgb::warp_gather_kernel, (unsigned)(W2.a.nblk + W2.b.nblk + 1), 256u, st, W2, /*allow_pdl=*/true);
            g_launches++;
        }
        CUDA_OK(cudaGetLastError());
        return;
    }
    size_gather_grid(pl, G);
    launch_dependent(mgb::gather_kernel, (unsigned)(G.nblk_h + G.nblk_l + G.nblk_g + 1), 256u, st, G,
                     /*allow_pdl=*/mid == nullptr && !pl->long_lists);
    g_launches++;
    CUDA_OK(cudaGetLastError());
}

}  // namespace

namespace {
// Everything after the inputs are on the host in canonical form: symbolic phase, uploads, replay lists.
// Dh: the operators restricted to this plan's quadrature rows (nloc x N), wloc: their weights.
void finish_plan(std::unique_ptr<mgb_plan>& pl, std::vector<mgb::HostCSR>& Dh, mgb::HostCSR& Rh, std::vector<double>& wloc,
                 int64_t n, int nD, int dim, int force_path, int64_t nnzD, int64_t out0 = 0, int64_t out1 = -1, int64_t n_primary = -1) {
        mgb_ctx* ctx = pl->ctx;
        const bool host_only = (ctx == nullptr);
        pl->NU = (int)(pl->N / n);
        cudaStream_t st = host_only ? nullptr : ctx->stream;
        bool use_elem = false;
        const bool want_hess = (force_path & MGB_PLAN_NO_HESSIAN) == 0;
        force_path &= 3;
        pl->has_hessian = want_hess;
        if (out1 < 0) out1 = pl->m;
        pl->out0 = out0; pl->m_out = out1 - out0;
        pl->n_primary = n_primary < 0 ? pl->nloc : n_primary;
        const bool sharded = out0 != 0 || out1 != pl->m || pl->n_primary != pl->nloc;
        if (sharded && force_path == MGB_PATH_CSR) throw std::runtime_error("sharded plans need the element path");
        if (force_path != MGB_PATH_CSR) {
            mgb::build_element_plan(Dh, Rh, n, wloc.data(), pl->bar, pl->ep, want_hess, out0, out1,
                                    /*allow_agg=*/getenv("MGB_NO_AGG") == nullptr);
            use_elem = pl->ep.ok && (pl->ep.dense || elem_supported(pl->ep.B, pl->ep.dim));
            if (!use_elem && sharded) throw std::runtime_error(std::string("sharded plans need the element path: ") + (pl->ep.ok ? "element type not instantiated" : pl->ep.why));
            if (!use_elem && force_path == MGB_PATH_ELEMENT)
                throw std::runtime_error("element path unavailable: " + (pl->ep.ok ? std::string("element type not instantiated") : pl->ep.why));
        }
        if (!host_only) pl->d_w.upload(wloc, st);
        if (use_elem) {
            auto& ep = pl->ep;
            pl->path = MGB_PATH_ELEMENT;
            pl->nnzH = (int64_t)ep.h_colidx.size();
            pl->h_rowptr = ep.h_rowptr; pl->h_colidx = ep.h_colidx;
            pl->n_hcontrib = (int64_t)ep.h_cidx.size(); pl->n_gcontrib = (int64_t)ep.g_cidx.size();
            validate_element_plan(ep);
            {
                const double avg = pl->nnzH ? (double)ep.h_cidx.size() / (double)pl->nnzH : 0.0;
                // lists averaging more than six contributions go to the lanes-per-output gather (fem1d L=16 level 12,
                // 10.7 per entry: 36.5 -> 21.5 us; at 5.4 per entry the thread-per-entry gather still wins: 69 vs 105 us;
                // MGB_LONG_AVG overrides for tuning runs)
                const char* ev = getenv("MGB_LONG_AVG");
                pl->long_lists = avg > (ev ? atof(ev) : 6.0);
            }
          if (!host_only) {
            if (ep.dense) { pl->d_dgdof.upload(ep.d_gdof, st); pl->d_dchunk.upload(ep.d_chunk, st); pl->d_drows.upload(ep.d_rows, st); }
            else { pl->d_lcols.upload(ep.lcols, st); pl->d_prec.upload(ep.prec, st); }
            if (pl->long_lists) {  // coarse levels: warp-per-entry over the CSR lists
                pl->d_hcptr.upload(ep.h_cptr, st); pl->d_hcidx.upload(ep.h_cidx, st);
                // chunk only where it pays: lists of a thousand contributions and more (coarsest levels)
                const int64_t ch = gather_chunk();
                if (ch > 0 && (int64_t)ep.h_cidx.size() >= 1024 * std::max<int64_t>(pl->nnzH, 1)) pl->ck_h.build(ep.h_cptr, ch, st);
                if (ch > 0 && (int64_t)ep.g_cidx.size() >= 1024 * std::max<int64_t>(pl->m_out, 1)) pl->ck_g.build(ep.g_cptr, ch, st);
            } else {   // two-wide ELL + long list for the thread-per-entry gather
                std::vector<int2> src2(pl->nnzH);
                std::vector<int64_t> lptr(1, 0);
                std::vector<int32_t> lidx, lt;
                for (int64_t t = 0; t < pl->nnzH; ++t) {
                    const int64_t c0 = ep.h_cptr[t], c1 = ep.h_cptr[t + 1];
                    if (c1 - c0 <= 2) {
                        src2[t].x = (c1 > c0) ? ep.h_cidx[c0] : 0;
                        src2[t].y = (c1 - c0 == 2) ? ep.h_cidx[c0 + 1] : -1;
                        if (c1 == c0) throw std::runtime_error("internal: Hessian entry without contribution");
                    } else {
                        src2[t].x = (int32_t)(-1 - (int64_t)(lptr.size() - 1));
                        src2[t].y = -1;
                        lidx.insert(lidx.end(), ep.h_cidx.begin() + c0, ep.h_cidx.begin() + c1);
                        lptr.push_back((int64_t)lidx.size());
                        lt.push_back((int32_t)t);
                    }
                }
                pl->d_hsrc2.upload(src2, st); pl->d_hlptr.upload(lptr, st); pl->d_hlidx.upload(lidx, st);
                pl->d_hlt.upload(lt, st); pl->n_long = (int64_t)lt.size();
                CUDA_OK(cudaStreamSynchronize(st));
            }
            pl->d_gcptr.upload(ep.g_cptr, st); pl->d_gcidx.upload(ep.g_cidx, st);
            const int64_t nrec = (ep.E + ep.agg - 1) / ep.agg;   // one record per aggregation group
            pl->d_sel.alloc((size_t)nrec * ep.lay.NS);
            pl->d_rel.alloc((size_t)std::max<int64_t>(nrec * ep.NU * ep.LPE, pl->m_out));
            if (pl->d_sel.p) CUDA_OK(cudaMemsetAsync(pl->d_sel.p, 0, pl->d_sel.bytes(), st));
            CUDA_OK(cudaMemsetAsync(pl->d_rel.p, 0, pl->d_rel.bytes(), st));
            const int epb = ep.dense ? 1 : MGB_ELEM_THREADS / ep.LPE;   // dense path: one CTA per chunk
            pl->nblocks_elem = (ep.E + epb - 1) / epb;

            pl->d_part.alloc((size_t)pl->nblocks_elem * 4);
            pl->d_scal_tmp.alloc(4);
            CUDA_OK(cudaStreamSynchronize(st));
            pl->dev_bytes = pl->d_lcols.bytes() + pl->d_prec.bytes() + pl->d_hcptr.bytes() + pl->d_hcidx.bytes() + pl->d_gcptr.bytes() +
                            pl->d_gcidx.bytes() + pl->d_dgdof.bytes() + pl->d_dchunk.bytes() + pl->d_drows.bytes() + pl->ck_h.bytes() + pl->ck_g.bytes() + pl->d_hsrc2.bytes() + pl->d_hlptr.bytes() + pl->d_hlidx.bytes() + pl->d_sel.bytes() + pl->d_rel.bytes() + pl->d_w.bytes();
          }
            // release host copies that are no longer needed
            std::vector<int32_t>().swap(ep.h_cidx); std::vector<int64_t>().swap(ep.h_cptr);
            std::vector<double>().swap(ep.prec); std::vector<double>().swap(ep.d_rows);
            std::vector<int32_t>().swap(ep.h_rowptr); std::vector<int32_t>().swap(ep.h_colidx);
        } else {
            pl->path = MGB_PATH_CSR;
            mgb::CsrPlan cp;
            mgb::build_csr_plan(Dh, Rh, pl->bar, cp, want_hess);
            pl->nnzH = (int64_t)cp.h_colidx.size();
            pl->h_rowptr = cp.h_rowptr; pl->h_colidx = cp.h_colidx;
            pl->n_hcontrib = (int64_t)cp.prod_coef.size();
            pl->n_hstored = mgb::sell_stored(cp.up_t.size(), cp.prod_ptr.data(), mgb::sell_sigma());
            if (!host_only) {
                pl->dev_bytes = mgb::csr_upload(cp, pl->bar, pl->csr, st) + pl->d_w.bytes();
                pl->d_scal_tmp.alloc(4);
                CUDA_OK(cudaStreamSynchronize(st));
            }
        }
        // nnz of the fine-space Hessian sum_jk D_j' diag D_k (structural), for the SURVEY 8(d) byte formula
        const int64_t nnzS = mgb::count_gram_pattern(Dh);
        pl->alg_bytes = algorithmic_bytes(pl->nloc, pl->N, nD, dim, nnzD, nnzS, Rh.nnz(), pl->nnzH);
        if (!host_only) for (auto& e : pl->ev) CUDA_OK(cudaEventCreate(&e));
}
}  // namespace

namespace {
// argument checks + the fields every plan needs; returns an error text or nullptr
const char* init_plan(mgb_plan& pl, mgb_ctx* ctx, int64_t n, int32_t nD, const mgb_csr* D, const mgb_csr* R, int32_t dim,
                      const mgb_barrier* barrier) {
    if (nD < 1 || nD > 8) return "nD must be 1..8";
    if (barrier->kind != MGB_BARRIER_EUCLIDIAN_POWER) return "unknown barrier kind";
    if (barrier->nidx < 1 || barrier->nidx > 4) return "barrier needs 1..4 idx entries (<=3 derivatives + s)";
    if (!(barrier->p >= 1.0)) return "p must be >= 1";
    pl.ctx = ctx;
    pl.n = n; pl.ND = nD; pl.dim = dim;
    pl.N = D[0].ncols; pl.m = R->ncols;
    if (R->nrows != pl.N) return "R rows must equal D columns";
    if (n <= 0 || pl.N % n) return "operator columns are not a multiple of n";
    pl.bar.kind = barrier->kind; pl.bar.nidx = barrier->nidx; pl.bar.p = barrier->p; pl.bar.slack = barrier->slack;
    for (int j = 0; j < barrier->nidx; ++j) {
        if (barrier->idx[j] < 0 || barrier->idx[j] >= nD) return "barrier idx outside 0..nD-1";
        pl.bar.idx[j] = barrier->idx[j];
    }
    if (barrier->nidx2 < 0 || barrier->nidx2 > 4) return "second cone needs 0 or 2..4 idx entries";
    if (barrier->nidx2 > 0) {
        if (barrier->nidx2 < 2 || !(barrier->p2 >= 1.0)) return "bad second cone";
        if (barrier->slack) return "slack variant supports a single cone";
        pl.bar.nidx2 = barrier->nidx2; pl.bar.p2 = barrier->p2;
        for (int j = 0; j < barrier->nidx2; ++j) {
            if (barrier->idx2[j] < 0 || barrier->idx2[j] >= nD) return "barrier idx2 outside 0..nD-1";
            pl.bar.idx2[j] = barrier->idx2[j];
        }
    }
    for (int k = 0; k < nD; ++k)
        if (D[k].nrows != n || D[k].ncols != pl.N) return "operator shapes differ";
    return nullptr;
}
}  // namespace

extern "C" {

const char* mgb_last_error(void) { return g_err.c_str(); }
int mgb_version(void) { return 100; }
int64_t mgb_launch_count(void) { return g_launches.load(); }
int mgb_pdl_active(void) { return g_pdl_ok ? 1 : 0; }

int mgb_ctx_create(int device, void* stream, mgb_ctx** out) {
    try {
        if (!out) return fail("mgb_ctx_create: out is NULL");
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            return fail(std::string("mgb_ctx_create: no CUDA device available (") + cudaGetErrorString(e) +
                        "); this library has no CPU path");
        if (device < 0 || device >= count) return fail("mgb_ctx_create: bad device index");
        CUDA_OK(cudaSetDevice(device));
        cudaDeviceProp prop{};
        CUDA_OK(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) return fail("mgb_ctx_create: device is not sm_100 class; kernels are built for sm_100a only");
        auto ctx = std::make_unique<mgb_ctx>();
        ctx->device = device;
        ctx->sm_count = prop.multiProcessorCount;
        // stream == NULL selects the legacy default stream (0): it orders with the caller's other
        // default-stream work (CUDA.jl / torch enqueue there unless told otherwise)
        ctx->stream = (cudaStream_t)stream;
        ctx->flag.alloc(1);
        *out = ctx.release();
        return 0;
    } catch (const std::exception& ex) { return fail(ex.what()); }
}

int mgb_ctx_destroy(mgb_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

int mgb_ctx_sync(mgb_ctx* ctx) {
    try {
        if (!ctx) return fail("mgb_ctx_sync: ctx is NULL");
        CUDA_OK(cudaStreamSynchronize(ctx->stream));
        return 0;
    } catch (const std::exception& ex) { return fail(ex.what()); }
}

int mgb_plan_create(mgb_ctx* ctx, int64_t n, int32_t nD, const mgb_csr* D, const mgb_csr* R, int32_t dim,
                    const double* x_host, const double* w_host, const mgb_barrier* barrier, int64_t row0,
                    int64_t row1, int32_t force_path, mgb_plan** out) {
    try {
        if (!D || !R || !w_host || !barrier || !out) return fail("mgb_plan_create: NULL argument");
        const bool host_only = (ctx == nullptr);  // symbolic-only plan: pattern/info queries, no numeric calls
        auto pl = std::make_unique<mgb_plan>();
        if (const char* why = init_plan(*pl, ctx, n, nD, D, R, dim, barrier)) return fail(std::string("mgb_plan_create: ") + why);
        pl->nloc = row1 - row0;
        (void)x_host;  // the Euclidian power barrier does not depend on x (p is constant); kept for the f(x,.) signature
        if (!host_only) CUDA_OK(cudaSetDevice(ctx->device));
        std::vector<mgb::HostCSR> Dh(nD);
        int64_t nnzD = 0;
        for (int k = 0; k < nD; ++k) {
            Dh[k] = to_host_csr(D[k], row0, row1);
            nnzD += Dh[k].nnz();
        }
        mgb::HostCSR Rh = to_host_csr(*R, 0, R->nrows);
        std::vector<double> wloc(w_host + row0, w_host + row1);
        finish_plan(pl, Dh, Rh, wloc, n, nD, dim, force_path, nnzD);
        *out = pl.release();
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_plan_create: ") + ex.what()); }
}

int mgb_plan_create_rows(mgb_ctx* ctx, int64_t n, int32_t nD, const mgb_csr* D, const mgb_csr* R, int32_t dim,
                         const double* x_host, const double* w_host, const mgb_barrier* barrier, int64_t nrows_sel,
                         const int64_t* rows_sel, int64_t n_primary, int64_t out0, int64_t out1, int32_t force_path,
                         mgb_plan** out) {
    try {
        if (!D || !R || !w_host || !barrier || !out || (!rows_sel && nrows_sel > 0)) return fail("mgb_plan_create_rows: NULL argument");
        if (nrows_sel < 0 || n_primary < 0 || n_primary > nrows_sel) return fail("mgb_plan_create_rows: bad row counts");
        const bool host_only = (ctx == nullptr);
        auto pl = std::make_unique<mgb_plan>();
        if (const char* why = init_plan(*pl, ctx, n, nD, D, R, dim, barrier)) return fail(std::string("mgb_plan_create_rows: ") + why);
        pl->nloc = nrows_sel;
        (void)x_host;
        if (!host_only) CUDA_OK(cudaSetDevice(ctx->device));
        std::vector<mgb::HostCSR> Dh(nD);
        int64_t nnzD = 0;
        for (int k = 0; k < nD; ++k) { Dh[k] = to_host_csr_rows(D[k], rows_sel, nrows_sel); nnzD += Dh[k].nnz(); }
        mgb::HostCSR Rh = to_host_csr(*R, 0, R->nrows);
        std::vector<double> wloc((size_t)nrows_sel);
        for (int64_t i = 0; i < nrows_sel; ++i) wloc[i] = w_host[rows_sel[i]];
        finish_plan(pl, Dh, Rh, wloc, n, nD, dim, force_path, nnzD, out0, out1, n_primary);
        *out = pl.release();
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_plan_create_rows: ") + ex.what()); }
}

int mgb_plan_create_local(mgb_ctx* ctx, int64_t n, int32_t nD, const mgb_hpc_block* D, const mgb_csr* R, int32_t dim,
                          const double* x_local_host, const double* w_local_host, const mgb_barrier* barrier,
                          int32_t force_path, mgb_plan** out) {
    try {
        if (!D || !R || !w_local_host || !barrier || !out) return fail("mgb_plan_create_local: NULL argument");
        const bool host_only = (ctx == nullptr);
        const int64_t nloc = D[0].nrows_local;
        std::vector<mgb_csr> shape(nD > 0 && nD <= 8 ? nD : 0);
        for (auto& c : shape) { c = mgb_csr{}; c.nrows = n; c.ncols = D[0].ncols_global; }
        for (int k = 0; k < (int)shape.size(); ++k) {
            if (D[k].nrows_local != nloc || D[k].row0 != D[0].row0) return fail("mgb_plan_create_local: operator row blocks differ");
            shape[k].ncols = D[k].ncols_global;
        }
        if (D[0].row0 < 0 || D[0].row0 + nloc > n) return fail("mgb_plan_create_local: row block outside 0..n");
        auto pl = std::make_unique<mgb_plan>();
        if (const char* why = init_plan(*pl, ctx, n, nD, shape.data(), R, dim, barrier)) return fail(std::string("mgb_plan_create_local: ") + why);
        pl->nloc = nloc;
        (void)x_local_host;
        if (!host_only) CUDA_OK(cudaSetDevice(ctx->device));
        // expand the compressed column ids (rowval -> col_indices[rowval]) to global ids; rows stay local
        std::vector<mgb::HostCSR> Dh(nD);
        int64_t nnzD = 0;
        for (int k = 0; k < nD; ++k) {
            const mgb_hpc_block& b = D[k];
            const int base = b.index_base;
            if (!b.colptr || (!b.rowval && nloc > 0 && b.colptr[nloc] - base > 0) || !b.nzval || (!b.col_indices && b.ncols_compressed > 0))
                return fail("mgb_plan_create_local: NULL array in operator block");
            mgb::HostCSR& H = Dh[k];
            H.nrows = nloc; H.ncols = b.ncols_global;
            H.ptr.resize(nloc + 1);
            for (int64_t i = 0; i <= nloc; ++i) H.ptr[i] = (int64_t)b.colptr[i] - base;
            if (H.ptr[0] != 0) return fail("mgb_plan_create_local: colptr must start at index_base");
            const int64_t cnt = H.ptr[nloc];
            H.idx.resize(cnt); H.val.assign(b.nzval, b.nzval + cnt);
            for (int64_t q = 0; q < cnt; ++q) {
                const int64_t cc = (int64_t)b.rowval[q] - base;
                if (cc < 0 || cc >= b.ncols_compressed) return fail("mgb_plan_create_local: compressed column id outside 0..ncols_compressed-1");
                const int64_t gc = (int64_t)b.col_indices[cc] - base;
                if (gc < 0 || gc >= b.ncols_global) return fail("mgb_plan_create_local: col_indices entry outside the matrix");
                H.idx[q] = (int32_t)gc;
            }
            sort_rows(H);
            nnzD += H.nnz();
        }
        mgb::HostCSR Rh = to_host_csr(*R, 0, R->nrows);
        std::vector<double> wloc(w_local_host, w_local_host + nloc);
        finish_plan(pl, Dh, Rh, wloc, n, nD, dim, force_path, nnzD);
        *out = pl.release();
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_plan_create_local: ") + ex.what()); }
}

int mgb_plan_destroy(mgb_plan* plan) {
    if (!plan) return 0;
    if (plan->ctx) {
        cudaSetDevice(plan->ctx->device);
        cudaStreamSynchronize(plan->ctx->stream);
    }
    for (auto& e : plan->ev) if (e) cudaEventDestroy(e);
    for (auto& g : plan->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    delete plan;
    return 0;
}

int mgb_plan_info(const mgb_plan* pl, int64_t* info, int32_t ninfo) {
    if (!pl || !info) return fail("mgb_plan_info: NULL argument");
    int64_t v[16] = {pl->path, pl->nloc, pl->ND, pl->m, pl->nnzH, pl->ep.E, pl->ep.B, pl->ep.B, pl->ep.lay.NS,
                     pl->n_hcontrib, pl->n_gcontrib, (int64_t)pl->dev_bytes, pl->N, pl->NU, pl->alg_bytes, pl->n_hstored};
    if (pl->path == MGB_PATH_CSR) { v[5] = 0; v[6] = 0; v[7] = 0; v[8] = 0; v[10] = 0; }
    for (int i = 0; i < ninfo && i < 16; ++i) info[i] = v[i];
    return 0;
}

int mgb_plan_pattern(const mgb_plan* pl, int32_t* rowptr_host, int32_t* colidx_host) {
    if (!pl || !rowptr_host || !colidx_host) return fail("mgb_plan_pattern: NULL argument");
    std::memcpy(rowptr_host, pl->h_rowptr.data(), pl->h_rowptr.size() * sizeof(int32_t));
    std::memcpy(colidx_host, pl->h_colidx.data(), pl->h_colidx.size() * sizeof(int32_t));
    return 0;
}

namespace {
void assemble_direct(mgb_plan* pl, const double* s_dev, const double* Dz0_dev, const double* c_dev, double t, int32_t flags,
                     double* scal_dev, double* grad_dev, double* hval_dev, double* Dz_dev) {
    if (pl->path == MGB_PATH_ELEMENT)
        assemble_element(pl, s_dev, Dz0_dev, c_dev, t, flags, scal_dev, grad_dev, hval_dev, Dz_dev);
    else
        g_launches += mgb::csr_assemble(pl->csr, pl->d_w.p, s_dev, Dz0_dev, c_dev, t, flags,
                                        scal_dev ? scal_dev : pl->d_scal_tmp.p, grad_dev, hval_dev, Dz_dev, pl->ctx->stream);
}

constexpr size_t GRAPH_CACHE = 8;

// One cudaGraphLaunch instead of two to four kernel launches: the launches of an assembly (with their programmatic
// dependent-launch edges) are captured once per argument tuple and replayed.  Returns false when graphs are off.
bool assemble_graph(mgb_plan* pl, const double* s_dev, const double* Dz0_dev, const double* c_dev, double t, int32_t flags,
                    double* scal_dev, double* grad_dev, double* hval_dev, double* Dz_dev) {
    mgb_ctx* ctx = pl->ctx;
    if (pl->graph_state < 0 || ctx->capturing) return false;
    if (pl->graph_state == 0) {
        const char* ev = getenv("MGB_GRAPH");
        // the legacy default stream cannot be captured
        pl->graph_state = (ctx->stream == nullptr || (ev && atoi(ev) == 0)) ? -1 : 1;
        if (pl->graph_state < 0) return false;
    }
    const void* key[8] = {s_dev, Dz0_dev, c_dev, scal_dev, grad_dev, hval_dev, Dz_dev, nullptr};
    pl->graph_clock++;
    for (auto& g : pl->graphs)
        if (g.exec && g.flags == flags && g.t == t && std::memcmp(g.key, key, sizeof(key)) == 0) {
            g.last_use = pl->graph_clock;
            CUDA_OK(cudaGraphLaunch(g.exec, ctx->stream));
            g_launches += g.kernels;
            pl->graph_hits++;
            return true;
        }
    // capture
    const int64_t before = g_launches.load();
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        pl->graph_state = -1;
        return false;
    }
    bool ok = true;
    std::string why;
    try { assemble_direct(pl, s_dev, Dz0_dev, c_dev, t, flags, scal_dev, grad_dev, hval_dev, Dz_dev); }
    catch (const std::exception& ex) { ok = false; why = ex.what(); }
    const cudaError_t ec = cudaStreamEndCapture(ctx->stream, &graph);
    const int kernels = (int)(g_launches.load() - before);
    g_launches = before;   // recorded, not executed
    cudaGraphExec_t exec = nullptr;
    if (ok && ec == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
        cudaGraphDestroy(graph);
    } else {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        if (!ok && ec == cudaSuccess) throw std::runtime_error(why);   // a genuine argument error: report it
        pl->graph_state = -1;   // capture / instantiation refused: plain launches from now on (visible in mgb_graph_stats)
        return false;
    }
    mgb_plan::GraphEntry* slot = nullptr;
    if (pl->graphs.size() < GRAPH_CACHE) { pl->graphs.emplace_back(); slot = &pl->graphs.back(); }
    else {
        slot = &pl->graphs[0];
        for (auto& g : pl->graphs) if (g.last_use < slot->last_use) slot = &g;
        if (slot->exec) cudaGraphExecDestroy(slot->exec);
    }
    std::memcpy(slot->key, key, sizeof(key));
    slot->t = t; slot->flags = flags; slot->exec = exec; slot->kernels = kernels; slot->last_use = pl->graph_clock;
    pl->graph_captures++;
    CUDA_OK(cudaGraphLaunch(exec, ctx->stream));
    g_launches += kernels;
    return true;
}
}  // namespace

int mgb_assemble(mgb_plan* pl, const double* s_dev, const double* Dz0_dev, const double* c_dev, double t,
                 int32_t flags, double* scal_dev, double* grad_dev, double* hval_dev, double* Dz_dev) {
    try {
        if (!pl || !s_dev || !c_dev) return fail("mgb_assemble: NULL argument");
        if (!pl->ctx) return fail("mgb_assemble: symbolic-only plan (created without a GPU context); no CPU path exists");
        if ((flags & MGB_WANT_HESS) && !pl->has_hessian) return fail("mgb_assemble: plan was created with MGB_PLAN_NO_HESSIAN");
        CUDA_OK(cudaSetDevice(pl->ctx->device));
        if (!assemble_graph(pl, s_dev, Dz0_dev, c_dev, t, flags, scal_dev, grad_dev, hval_dev, Dz_dev))
            assemble_direct(pl, s_dev, Dz0_dev, c_dev, t, flags, scal_dev, grad_dev, hval_dev, Dz_dev);
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_assemble: ") + ex.what()); }
}

int mgb_graph_stats(const mgb_plan* pl, int64_t* stats3) {
    if (!pl || !stats3) return fail("mgb_graph_stats: NULL argument");
    stats3[0] = pl->graph_state; stats3[1] = pl->graph_captures; stats3[2] = pl->graph_hits;
    return 0;
}

int mgb_graph_begin(mgb_ctx* ctx) {
    try {
        if (!ctx) return fail("mgb_graph_begin: ctx is NULL");
        if (ctx->capturing) return fail("mgb_graph_begin: a capture is already open on this context");
        if (!ctx->stream) return fail("mgb_graph_begin: the legacy default stream cannot be captured; create the context on a real stream");
        CUDA_OK(cudaSetDevice(ctx->device));
        CUDA_OK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        ctx->capturing = true;
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_graph_begin: ") + ex.what()); }
}

int mgb_graph_end(mgb_ctx* ctx, mgb_graph** out) {
    try {
        if (!ctx || !out) return fail("mgb_graph_end: NULL argument");
        if (!ctx->capturing) return fail("mgb_graph_end: no capture open (mgb_graph_begin)");
        ctx->capturing = false;
        cudaGraph_t graph = nullptr;
        CUDA_OK(cudaStreamEndCapture(ctx->stream, &graph));
        auto g = std::make_unique<mgb_graph>();
        g->ctx = ctx;
        size_t nn = 0;
        CUDA_OK(cudaGraphGetNodes(graph, nullptr, &nn));
        g->kernels = (int)nn;
        const cudaError_t e = cudaGraphInstantiate(&g->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) throw std::runtime_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
        *out = g.release();
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_graph_end: ") + ex.what()); }
}

int mgb_graph_launch(mgb_graph* g) {
    try {
        if (!g || !g->exec) return fail("mgb_graph_launch: NULL graph");
        CUDA_OK(cudaSetDevice(g->ctx->device));
        CUDA_OK(cudaGraphLaunch(g->exec, g->ctx->stream));
        g_launches += g->kernels;
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_graph_launch: ") + ex.what()); }
}

int mgb_graph_destroy(mgb_graph* g) {
    if (!g) return 0;
    if (g->exec) { cudaSetDevice(g->ctx->device); cudaStreamSynchronize(g->ctx->stream); cudaGraphExecDestroy(g->exec); }
    delete g;
    return 0;
}

int mgb_assemble_host(mgb_plan* pl, const double* s_host, const double* Dz0_host, const double* c_host,
                      int32_t upload_inputs, double t, int32_t flags, double* scal_host, double* grad_host,
                      double* hval_host, double* Dz_host) {
    try {
        if (!pl || !s_host) return fail("mgb_assemble_host: NULL argument");
        if (!pl->ctx) return fail("mgb_assemble_host: symbolic-only plan (created without a GPU context); no CPU path exists");
        CUDA_OK(cudaSetDevice(pl->ctx->device));
        cudaStream_t st = pl->ctx->stream;
        const size_t nd = (size_t)pl->nloc * pl->ND;
        if (!pl->st_s.p) {
            pl->st_s.alloc(pl->m); pl->st_dz0.alloc(nd); pl->st_c.alloc(nd); pl->st_scal.alloc(4);
            pl->st_grad.alloc(std::max<int64_t>(pl->m_out, 1)); pl->st_hval.alloc(pl->nnzH); pl->st_dz.alloc(nd);
            CUDA_OK(cudaMemsetAsync(pl->st_dz0.p, 0, nd * 8, st));
            CUDA_OK(cudaMemsetAsync(pl->st_c.p, 0, nd * 8, st));
            upload_inputs = 1;
        }
        CUDA_OK(cudaMemcpyAsync(pl->st_s.p, s_host, pl->m * 8, cudaMemcpyHostToDevice, st));
        if (upload_inputs) {
            if (Dz0_host) CUDA_OK(cudaMemcpyAsync(pl->st_dz0.p, Dz0_host, nd * 8, cudaMemcpyHostToDevice, st));
            if (c_host) CUDA_OK(cudaMemcpyAsync(pl->st_c.p, c_host, nd * 8, cudaMemcpyHostToDevice, st));
        }
        int rc = mgb_assemble(pl, pl->st_s.p, pl->st_dz0.p, pl->st_c.p, t, flags, pl->st_scal.p, pl->st_grad.p,
                              pl->st_hval.p, (flags & MGB_STORE_DZ) ? pl->st_dz.p : nullptr);
        if (rc) return rc;
        if (scal_host) CUDA_OK(cudaMemcpyAsync(scal_host, pl->st_scal.p, 32, cudaMemcpyDeviceToHost, st));
        if ((flags & MGB_WANT_GRAD) && grad_host)
            CUDA_OK(cudaMemcpyAsync(grad_host, pl->st_grad.p, pl->m_out * 8, cudaMemcpyDeviceToHost, st));
        if ((flags & MGB_WANT_HESS) && hval_host)
            CUDA_OK(cudaMemcpyAsync(hval_host, pl->st_hval.p, pl->nnzH * 8, cudaMemcpyDeviceToHost, st));
        if ((flags & MGB_STORE_DZ) && Dz_host)
            CUDA_OK(cudaMemcpyAsync(Dz_host, pl->st_dz.p, nd * 8, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_assemble_host: ") + ex.what()); }
}

int mgb_apply_D(mgb_plan* pl, const double* s_dev, const double* Dz0_dev, double* Dz_dev) {
    try {
        if (!pl || !s_dev || !Dz_dev) return fail("mgb_apply_D: NULL argument");
        if (!pl->ctx) return fail("mgb_apply_D: symbolic-only plan; no CPU path exists");
        CUDA_OK(cudaSetDevice(pl->ctx->device));
        // c is only read for the <c,Dz> scalar: reuse Dz0 or s as a harmless stand-in is not allowed,
        // so keep a zero block
        const size_t nd = (size_t)pl->nloc * pl->ND;
        if (!pl->st_c.p) { pl->st_c.alloc(nd); CUDA_OK(cudaMemsetAsync(pl->st_c.p, 0, nd * 8, pl->ctx->stream)); }
        return mgb_assemble(pl, s_dev, Dz0_dev, pl->st_c.p, 0.0, MGB_STORE_DZ, nullptr, nullptr, nullptr, Dz_dev);
    } catch (const std::exception& ex) { return fail(std::string("mgb_apply_D: ") + ex.what()); }
}

int mgb_map_barrier(mgb_ctx* ctx, const mgb_barrier* barrier, int32_t nD, int64_t n, const double* Dz_dev,
                    int32_t which, double* out_dev) {
    try {
        if (!ctx || !barrier || !Dz_dev || !out_dev) return fail("mgb_map_barrier: NULL argument");
        if (which < 0 || which > 2) return fail("mgb_map_barrier: which must be 0,1,2");
        if (barrier->kind != MGB_BARRIER_EUCLIDIAN_POWER || barrier->nidx < 2 || barrier->nidx > 4)
            return fail("mgb_map_barrier: unsupported barrier");
        if (nD < barrier->nidx + 1 + (barrier->slack ? 1 : 0)) return fail("mgb_map_barrier: nD too small for idx");
        CUDA_OK(cudaSetDevice(ctx->device));
        mgb::BarrierDesc bd;
        bd.kind = barrier->kind; bd.nidx = barrier->nidx; bd.p = barrier->p; bd.slack = barrier->slack;
        for (int j = 0; j < barrier->nidx; ++j) bd.idx[j] = barrier->idx[j];
        if (n > 0) g_launches += mgb::csr_map_barrier(bd, nD, n, Dz_dev, which, out_dev, ctx->stream);
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_map_barrier: ") + ex.what()); }
}

int mgb_all_isfinite(mgb_ctx* ctx, const double* v_dev, int64_t len, int32_t* flag_host) {
    try {
        if (!ctx || !flag_host) return fail("mgb_all_isfinite: NULL argument");
        CUDA_OK(cudaSetDevice(ctx->device));
        int one = 1;
        CUDA_OK(cudaMemcpyAsync(ctx->flag.p, &one, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        if (len > 0) {
            const int nb = (int)std::min<int64_t>((len + 255) / 256, (int64_t)ctx->sm_count * 8);
            mgb::isfinite_kernel<<<nb, 256, 0, ctx->stream>>>(v_dev, len, ctx->flag.p);
            g_launches++;
        }
        int res = 1;
        CUDA_OK(cudaMemcpyAsync(&res, ctx->flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_OK(cudaStreamSynchronize(ctx->stream));
        *flag_host = res;
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_all_isfinite: ") + ex.what()); }
}

int mgb_reduce(mgb_ctx* ctx, int32_t op, const double* x_dev, const double* y_dev, int64_t len, double* out_dev,
               double* out_host) {
    try {
        if (!ctx || (!x_dev && len > 0)) return fail("mgb_reduce: NULL argument");
        if (op < MGB_REDUCE_DOT || op > MGB_REDUCE_MAXABS) return fail("mgb_reduce: unknown operation");
        if (op == MGB_REDUCE_DOT && !y_dev && len > 0) return fail("mgb_reduce: dot needs two vectors");
        if (len < 0) return fail("mgb_reduce: negative length");
        CUDA_OK(cudaSetDevice(ctx->device));
        if (!ctx->red.p) {
            ctx->red.alloc(mgb::REDUCE_BLOCKS + 1);
            ctx->ticket.alloc(1);
            CUDA_OK(cudaMemsetAsync(ctx->ticket.p, 0, sizeof(unsigned), ctx->stream));
        }
        double* out = out_dev ? out_dev : ctx->red.p + mgb::REDUCE_BLOCKS;
        mgb::reduce_kernel<<<mgb::REDUCE_BLOCKS, 256, 0, ctx->stream>>>(x_dev, y_dev, len, op, ctx->red.p, ctx->ticket.p, out);
        g_launches++;
        CUDA_OK(cudaGetLastError());
        if (out_host) {
            CUDA_OK(cudaMemcpyAsync(out_host, out, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_OK(cudaStreamSynchronize(ctx->stream));
        }
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_reduce: ") + ex.what()); }
}

int mgb_diag_scale(mgb_ctx* ctx, const double* w_dev, const double* y_dev, int64_t n, int64_t ld, int32_t col,
                   double* out_dev) {
    try {
        if (!ctx || !w_dev || !y_dev || !out_dev) return fail("mgb_diag_scale: NULL argument");
        CUDA_OK(cudaSetDevice(ctx->device));
        if (n > 0) {
            mgb::diag_scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(w_dev, y_dev + (int64_t)col * ld, n, out_dev);
            g_launches++;
        }
        CUDA_OK(cudaGetLastError());
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_diag_scale: ") + ex.what()); }
}

int mgb_spmat_create(mgb_ctx* ctx, const mgb_csr* A, mgb_spmat** out) {
    try {
        if (!ctx || !A || !out) return fail("mgb_spmat_create: NULL argument");
        CUDA_OK(cudaSetDevice(ctx->device));
        mgb::HostCSR H = to_host_csr(*A, 0, A->nrows);
        mgb::HostCSR T = mgb::transpose(H);
        auto M = std::make_unique<mgb_spmat>();
        M->ctx = ctx; M->nrows = H.nrows; M->ncols = H.ncols;
        cudaStream_t st = ctx->stream;
        M->ptr.upload(H.ptr, st); M->idx.upload(H.idx, st); M->val.upload(H.val, st);
        M->tptr.upload(T.ptr, st); M->tidx.upload(T.idx, st); M->tval.upload(T.val, st);
        auto longest = [](const mgb::HostCSR& C) { int64_t mx = 0; for (int64_t i = 0; i < C.nrows; ++i) mx = std::max(mx, C.ptr[i + 1] - C.ptr[i]); return mx; };
        if (longest(H) >= 1024) M->ck.build(H.ptr, 256, st);
        if (longest(T) >= 1024) M->tck.build(T.ptr, 256, st);
        CUDA_OK(cudaStreamSynchronize(st));
        *out = M.release();
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_spmat_create: ") + ex.what()); }
}

int mgb_spmat_destroy(mgb_spmat* A) {
    if (!A) return 0;
    cudaSetDevice(A->ctx->device);
    cudaStreamSynchronize(A->ctx->stream);
    delete A;
    return 0;
}

int mgb_spmat_mv(mgb_spmat* A, int32_t trans, double alpha, const double* x_dev, double beta, const double* y0_dev,
                 double* y_dev) {
    try {
        if (!A || !x_dev || !y_dev) return fail("mgb_spmat_mv: NULL argument");
        CUDA_OK(cudaSetDevice(A->ctx->device));
        const int64_t nr = trans ? A->ncols : A->nrows;
        if (nr > 0) {
            cudaStream_t st = A->ctx->stream;
            ChunkedList& ck = trans ? A->tck : A->ck;
            const int64_t* ptr = trans ? A->tptr.p : A->ptr.p;
            const int32_t* idx = trans ? A->tidx.p : A->idx.p;
            const double* val = trans ? A->tval.p : A->val.p;
            if (ck.nchunks > 0) {
                mgb::spmv_chunk_kernel<<<(unsigned)((ck.nchunks * 32 + 255) / 256), 256, 0, st>>>(ck.nchunks, ck.kptr.p, idx, val, x_dev, ck.part.p);
                mgb::spmv_fold_kernel<<<(unsigned)((nr * 32 + 255) / 256), 256, 0, st>>>(nr, ck.pptr.p, ck.part.p, alpha, beta, y0_dev, y_dev);
                g_launches++;
            } else
                mgb::spmv_kernel<<<(unsigned)((nr + 255) / 256), 256, 0, st>>>(nr, ptr, idx, val, alpha, x_dev, beta, y0_dev, y_dev);
            g_launches++;
        }
        CUDA_OK(cudaGetLastError());
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_spmat_mv: ") + ex.what()); }
}

int mgb_gather_idx(mgb_ctx* ctx, const double* src_dev, const int32_t* idx_dev, int64_t count, double* out_dev) {
    try {
        if (!ctx) return fail("mgb_gather_idx: NULL argument");
        CUDA_OK(cudaSetDevice(ctx->device));
        if (count > 0) {
            mgb::gather_idx_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(src_dev, idx_dev, count, out_dev);
            g_launches++;
        }
        CUDA_OK(cudaGetLastError());
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_gather_idx: ") + ex.what()); }
}

int mgb_scatter_add_idx(mgb_ctx* ctx, const double* src_dev, const int32_t* idx_dev, int64_t count, double* dst_dev) {
    try {
        if (!ctx) return fail("mgb_scatter_add_idx: NULL argument");
        CUDA_OK(cudaSetDevice(ctx->device));
        if (count > 0) {
            mgb::scatter_add_idx_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(src_dev, idx_dev, count, dst_dev);
            g_launches++;
        }
        CUDA_OK(cudaGetLastError());
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_scatter_add_idx: ") + ex.what()); }
}

int mgb_segsum_idx(mgb_ctx* ctx, const double* src_dev, const int32_t* ptr_dev, const int32_t* idx_dev, int64_t nout,
                   double* dst_dev) {
    try {
        if (!ctx) return fail("mgb_segsum_idx: NULL argument");
        CUDA_OK(cudaSetDevice(ctx->device));
        if (nout > 0) {
            mgb::segsum_idx_kernel<<<(unsigned)((nout + 255) / 256), 256, 0, ctx->stream>>>(src_dev, ptr_dev, idx_dev, nout, dst_dev);
            g_launches++;
        }
        CUDA_OK(cudaGetLastError());
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_segsum_idx: ") + ex.what()); }
}

int mgb_time_assemble(mgb_plan* pl, const double* s_dev, const double* Dz0_dev, const double* c_dev, double t,
                      int32_t flags, double* scal_dev, double* grad_dev, double* hval_dev, int32_t reps,
                      int32_t flush_l2, float* ms_total, float* ms_kernel_element, float* ms_kernel_gather) {
    try {
        if (!pl || reps < 1) return fail("mgb_time_assemble: bad argument");
        if (!pl->ctx) return fail("mgb_time_assemble: symbolic-only plan; no CPU path exists");
        mgb_ctx* ctx = pl->ctx;
        CUDA_OK(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        const int64_t flush_len = (int64_t)256 << 20 >> 3;  // 256 MiB > 126 MB L2
        if (flush_l2 && !ctx->flush.p) { ctx->flush.alloc(flush_len); CUDA_OK(cudaMemsetAsync(ctx->flush.p, 0, flush_len * 8, st)); }
        double tot = 0.0, tel = 0.0, tga = 0.0;
        const bool split = pl->path == MGB_PATH_ELEMENT && (ms_kernel_element || ms_kernel_gather);
        for (int r = 0; r < reps; ++r) {
            if (flush_l2) {
                mgb::l2_flush_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(ctx->flush.p, flush_len, (double)r, flush_l2 == 2 ? 1 : 0);
            }
            CUDA_OK(cudaEventRecord(pl->ev[0], st));
            int rc = mgb_assemble(pl, s_dev, Dz0_dev, c_dev, t, flags, scal_dev, grad_dev, hval_dev, nullptr);
            if (rc) return rc;
            CUDA_OK(cudaEventRecord(pl->ev[1], st));
            CUDA_OK(cudaEventSynchronize(pl->ev[1]));
            float ms = 0.f;
            CUDA_OK(cudaEventElapsedTime(&ms, pl->ev[0], pl->ev[1]));
            tot += ms;
        }
        if (split) {
            // second pass with an event between the two kernels (kept out of the totals above)
            for (int r = 0; r < reps; ++r) {
                if (flush_l2) mgb::l2_flush_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(ctx->flush.p, flush_len, (double)r, flush_l2 == 2 ? 1 : 0);
                CUDA_OK(cudaEventRecord(pl->ev[0], st));
                assemble_element(pl, s_dev, Dz0_dev, c_dev, t, flags, scal_dev, grad_dev, hval_dev, nullptr, pl->ev[2]);
                CUDA_OK(cudaEventRecord(pl->ev[1], st));
                CUDA_OK(cudaEventSynchronize(pl->ev[1]));
                float ms = 0.f, ms2 = 0.f;
                CUDA_OK(cudaEventElapsedTime(&ms, pl->ev[0], pl->ev[2]));
                CUDA_OK(cudaEventElapsedTime(&ms2, pl->ev[2], pl->ev[1]));
                tel += ms;
                tga += ms2;
            }
        }
        if (ms_total) *ms_total = (float)(tot / reps);
        if (ms_kernel_element) *ms_kernel_element = (float)(tel / reps);
        if (ms_kernel_gather) *ms_kernel_gather = (float)(tga / reps);
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_time_assemble: ") + ex.what()); }
}

int mgb_host_register(void* ptr_host, int64_t bytes) {
    try {
        if (!ptr_host || bytes <= 0) return fail("mgb_host_register: bad argument");
        CUDA_OK(cudaHostRegister(ptr_host, (size_t)bytes, cudaHostRegisterDefault));
        return 0;
    } catch (const std::exception& ex) { cudaGetLastError(); return fail(std::string("mgb_host_register: ") + ex.what()); }
}

int mgb_host_unregister(void* ptr_host) {
    try {
        if (!ptr_host) return fail("mgb_host_unregister: NULL argument");
        CUDA_OK(cudaHostUnregister(ptr_host));
        return 0;
    } catch (const std::exception& ex) { cudaGetLastError(); return fail(std::string("mgb_host_unregister: ") + ex.what()); }
}

int mgb_copy_to_host(mgb_ctx* ctx, void* dst_host, const void* src_dev, int64_t bytes) {
    try {
        if (!ctx || (bytes > 0 && (!dst_host || !src_dev))) return fail("mgb_copy_to_host: NULL argument");
        CUDA_OK(cudaSetDevice(ctx->device));
        if (bytes > 0) CUDA_OK(cudaMemcpyAsync(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_OK(cudaStreamSynchronize(ctx->stream));
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_copy_to_host: ") + ex.what()); }
}

// ------------------------------------------------------------------ multi-GPU: owner-computes sharding
int mgb_dist_plan_create(mgb_ctx* ctx, int64_t n, int32_t nD, const mgb_csr* D, const mgb_csr* R, int32_t dim,
                         const double* x_host, const double* w_host, const mgb_barrier* barrier, int32_t rank,
                         int32_t nranks, const int64_t* row_part, const int64_t* out_part, mgb_plan** out) {
    try {
        if (!D || !R || !w_host || !barrier || !row_part || !out_part || !out) return fail("mgb_dist_plan_create: NULL argument");
        if (nranks < 1 || nranks > mgb::DIST_LL_RANKS || rank < 0 || rank >= nranks) return fail("mgb_dist_plan_create: bad rank / nranks (1..16)");
        if (nD < 1 || nD > 8) return fail("mgb_dist_plan_create: nD must be 1..8");
        if (row_part[0] != 0 || row_part[nranks] != n || out_part[0] != 0 || out_part[nranks] != R->ncols)
            return fail("mgb_dist_plan_create: partitions must cover [0,n) and [0,m)");
        for (int r = 0; r < nranks; ++r)
            if (row_part[r + 1] < row_part[r] || out_part[r + 1] < out_part[r]) return fail("mgb_dist_plan_create: partition offsets must be non-decreasing");
        // replicated element structure (host only, no Hessian lists): which elements touch this rank's output rows
        std::vector<int64_t> rows;
        int64_t n_primary = 0;
        {
            mgb_plan tmp;
            if (const char* why = init_plan(tmp, nullptr, n, nD, D, R, dim, barrier)) return fail(std::string("mgb_dist_plan_create: ") + why);
            std::vector<mgb::HostCSR> Dh(nD);
            for (int k = 0; k < nD; ++k) Dh[k] = to_host_csr(D[k], 0, n);
            mgb::HostCSR Rh = to_host_csr(*R, 0, R->nrows);
            mgb::ElementPlan gp;
            mgb::build_element_plan(Dh, Rh, n, w_host, tmp.bar, gp, /*want_hessian=*/false);
            if (!gp.ok) return fail("mgb_dist_plan_create: sharded plans need the element path: " + gp.why);
            if (!elem_supported(gp.B, gp.dim)) return fail("mgb_dist_plan_create: sharded plans need the element path: element type not instantiated");
            // decided from the replicated (global) structure so that every rank takes the same branch: only levels whose
            // id-like operators own one column per row (the fine ones, where the work is) are sharded
            if (!gp.fine) return fail("mgb_dist_plan_create: sharded plans need the element path with thread-per-entry gather: coarse level");
            for (int r = 0; r <= nranks; ++r)
                if (row_part[r] % gp.B) return fail("mgb_dist_plan_create: row partition splits a broken element");
            mgb::dist_select_elements(gp.lcols, gp.E, gp.NU, gp.LPE, gp.B, rank, nranks, out_part, rows, n_primary);
        }
        mgb_plan* plraw = nullptr;
        int rc = mgb_plan_create_rows(ctx, n, nD, D, R, dim, x_host, w_host, barrier, (int64_t)rows.size(), rows.data(), n_primary,
                                      out_part[rank], out_part[rank + 1], MGB_PATH_ELEMENT | MGB_PLAN_TWO_STAGE, &plraw);
        if (rc) return rc;
        std::unique_ptr<mgb_plan, int (*)(mgb_plan*)> pl(plraw, mgb_plan_destroy);
        if (pl->long_lists)
            return fail("mgb_dist_plan_create: sharded plans need the element path: this level's gather runs lanes-per-entry (coarse level)");
        auto dd = std::make_unique<mgb_plan::Dist>();
        dd->rank = rank; dd->nranks = nranks; dd->rows = std::move(rows);
        if (ctx) {
            CUDA_OK(cudaSetDevice(ctx->device));
            cudaStream_t st = ctx->stream;
            dd->hval.alloc((size_t)std::max<int64_t>(pl->nnzH, 1)); dd->grad.alloc((size_t)std::max<int64_t>(pl->m_out, 1));
            dd->scal.alloc(4); dd->err.alloc(1);
            CUDA_OK(cudaMemsetAsync(dd->err.p, 0, sizeof(int), st));
            dd->off_flags = (size_t)2 * mgb::DIST_LL_RANKS * 8 * sizeof(unsigned long long);
            dd->off_full = dd->off_flags + 128;
            dd->window_bytes = dd->off_full + (size_t)2 * std::max<int64_t>(pl->m, 1) * sizeof(double);
            dd->counter.alloc(1);
            CUDA_OK(cudaMemsetAsync(dd->counter.p, 0, sizeof(unsigned int), st));
            CUDA_OK(cudaMalloc(&dd->window, dd->window_bytes));
            CUDA_OK(cudaMemsetAsync(dd->window, 0, dd->window_bytes, st));   // epoch tag 0 never matches a live epoch
            CUDA_OK(cudaStreamSynchronize(st));
            pl->dev_bytes += dd->window_bytes + dd->hval.bytes() + dd->grad.bytes();
            if (const char* ev = getenv("MGB_DIST_TIMEOUT_S")) dd->timeout_s = atof(ev) > 0 ? atof(ev) : dd->timeout_s;
        }
        pl->dist = std::move(dd);
        *out = pl.release();
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_dist_plan_create: ") + ex.what()); }
}

int mgb_dist_info(const mgb_plan* pl, int64_t* info, int32_t ninfo) {
    try {
        if (!pl || !pl->dist || !info) return fail("mgb_dist_info: not a distributed plan");
        const auto& dd = *pl->dist;
        int err = 0;
        if (pl->ctx && dd.err.p) {
            CUDA_OK(cudaSetDevice(pl->ctx->device));
            CUDA_OK(cudaMemcpy(&err, dd.err.p, sizeof(int), cudaMemcpyDeviceToHost));
        }
        int64_t v[16] = {dd.rank, dd.nranks, pl->nnzH, pl->m_out, pl->out0, pl->out0 + pl->m_out, pl->nloc, pl->n_primary,
                         pl->ep.E, 0, 0, (int64_t)(dd.window_bytes / 8), (int64_t)dd.epoch, err, 0, 0};
        for (int i = 0; i < ninfo && i < 16; ++i) info[i] = v[i];
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_dist_info: ") + ex.what()); }
}

int mgb_dist_rows(const mgb_plan* pl, int64_t* rows_host) {
    if (!pl || !pl->dist || !rows_host) return fail("mgb_dist_rows: not a distributed plan");
    std::memcpy(rows_host, pl->dist->rows.data(), pl->dist->rows.size() * sizeof(int64_t));
    return 0;
}

int mgb_dist_pattern(const mgb_plan* pl, int32_t* rowptr_host, int32_t* colidx_host) {
    if (!pl || !pl->dist) return fail("mgb_dist_pattern: not a distributed plan");
    return mgb_plan_pattern(pl, rowptr_host, colidx_host);
}

int mgb_dist_window(mgb_plan* pl, void** window_dev, int64_t* bytes) {
    if (!pl || !pl->dist || !pl->dist->window) return fail("mgb_dist_window: plan has no exchange window (symbolic-only or not distributed)");
    if (window_dev) *window_dev = pl->dist->window;
    if (bytes) *bytes = (int64_t)pl->dist->window_bytes;
    return 0;
}

int mgb_dist_export(mgb_plan* pl, mgb_ipc_handle* handle) {
    try {
        if (!pl || !pl->dist || !pl->dist->window || !handle) return fail("mgb_dist_export: plan has no exchange window");
        static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(mgb_ipc_handle), "IPC handle size");
        CUDA_OK(cudaSetDevice(pl->ctx->device));
        cudaIpcMemHandle_t h;
        CUDA_OK(cudaIpcGetMemHandle(&h, pl->dist->window));
        std::memcpy(handle->bytes, &h, sizeof(h));
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_dist_export: ") + ex.what()); }
}

int mgb_dist_attach(mgb_plan* pl, const mgb_ipc_handle* handles) {
    try {
        if (!pl || !pl->dist || !pl->dist->window || !handles) return fail("mgb_dist_attach: plan has no exchange window");
        auto& dd = *pl->dist;
        CUDA_OK(cudaSetDevice(pl->ctx->device));
        for (int p = 0; p < dd.nranks; ++p) {
            if (p == dd.rank) { dd.peer[p] = dd.window; continue; }
            cudaIpcMemHandle_t h;
            std::memcpy(&h, handles[p].bytes, sizeof(h));
            void* ptr = nullptr;
            CUDA_OK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
            dd.peer[p] = ptr; dd.peer_ipc[p] = true;
        }
        dd.attached = true;
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_dist_attach: ") + ex.what()); }
}

int mgb_dist_attach_local(mgb_plan* pl, void* const* windows_dev) {
    if (!pl || !pl->dist || !pl->dist->window || !windows_dev) return fail("mgb_dist_attach_local: plan has no exchange window");
    auto& dd = *pl->dist;
    for (int p = 0; p < dd.nranks; ++p) {
        if (!windows_dev[p]) return fail("mgb_dist_attach_local: NULL window");
        dd.peer[p] = (p == dd.rank) ? dd.window : windows_dev[p];
    }
    dd.attached = true;
    return 0;
}

namespace {
mgb::DistScal make_dist_scal(mgb_plan* pl, bool publish_only) {
    auto& dd = *pl->dist;
    mgb::DistScal S{};
    S.rank = dd.rank; S.nranks = dd.nranks; S.epoch = dd.epoch; S.publish_only = publish_only ? 1 : 0;
    for (int p = 0; p < dd.nranks; ++p) S.win[p] = static_cast<unsigned long long*>(dd.peer[p]);
    S.timeout_ns = (unsigned long long)(dd.timeout_s * 1e9);
    S.err = dd.err.p;
    return S;
}

void dist_outputs(mgb_plan* pl, const double** hval_own_dev, const double** grad_own_dev, const double** scal_dev) {
    auto& dd = *pl->dist;
    if (hval_own_dev) *hval_own_dev = dd.hval.p;
    if (grad_own_dev) *grad_own_dev = dd.grad.p;
    if (scal_dev) *scal_dev = dd.scal.p;
}

// element kernel + gather kernel of a new epoch on this rank's elements; the gather's scalar block publishes the
// partial sums to every rank and - fused mode - collects the peers' words and sums them in rank order
void dist_launch(mgb_plan* pl, const double* s_dev, const double* Dz0_dev, const double* c_dev, double t, int flags, bool fused) {
    auto& dd = *pl->dist;
    if (dd.nranks > 1 && !dd.attached) throw std::runtime_error("peers not attached (mgb_dist_attach)");
    if ((flags & MGB_WANT_HESS) && !pl->has_hessian) throw std::runtime_error("plan was created without a Hessian");
    CUDA_OK(cudaSetDevice(pl->ctx->device));
    dd.epoch++;
    if (dd.epoch == 0) dd.epoch = 2;   // tag 0 marks "never written"; keep the parity sequence (…, 0xFFFFFFFF, 2, 3, …)
    dd.finish_pending = !fused;
    const mgb::DistScal S = make_dist_scal(pl, !fused);
    assemble_element(pl, s_dev, Dz0_dev, c_dev, t, flags & 7, dd.scal.p, dd.grad.p, dd.hval.p, nullptr, nullptr, &S);
}
}  // namespace

namespace {
mgb::DistGather make_dist_gather(mgb_plan* pl, const double* s_own_dev) {
    auto& dd = *pl->dist;
    mgb::DistGather G{};
    G.rank = dd.rank; G.nranks = dd.nranks; G.epoch = dd.s_epoch;
    const size_t par = (size_t)(dd.s_epoch & 1ull);
    for (int p = 0; p < dd.nranks; ++p) {
        char* base = static_cast<char*>(dd.peer[p]);
        G.full[p] = reinterpret_cast<double*>(base + dd.off_full) + par * (size_t)pl->m;
        G.flag[p] = reinterpret_cast<unsigned long long*>(base + dd.off_flags);
    }
    G.own = s_own_dev; G.off = pl->out0; G.count = pl->m_out;
    G.counter = dd.counter.p; G.timeout_ns = (unsigned long long)(dd.timeout_s * 1e9); G.err = dd.err.p;
    return G;
}
}  // namespace

int mgb_dist_s_publish(mgb_plan* pl, const double* s_own_dev) {
    try {
        if (!pl || !pl->dist || !pl->ctx || (!s_own_dev && pl->m_out > 0)) return fail("mgb_dist_s_publish: NULL argument / not a distributed device plan");
        auto& dd = *pl->dist;
        if (!dd.attached) return fail("mgb_dist_s_publish: peers not attached (mgb_dist_attach)");
        CUDA_OK(cudaSetDevice(pl->ctx->device));
        dd.s_epoch++;
        const mgb::DistGather G = make_dist_gather(pl, s_own_dev);
        mgb::dist_s_scatter_kernel<<<(unsigned)std::max<int64_t>((pl->m_out + 255) / 256, 1), 256, 0, pl->ctx->stream>>>(G);
        g_launches++;
        CUDA_OK(cudaGetLastError());
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_dist_s_publish: ") + ex.what()); }
}

int mgb_dist_s_wait(mgb_plan* pl, const double** s_full_dev) {
    try {
        if (!pl || !pl->dist || !pl->ctx) return fail("mgb_dist_s_wait: not a distributed device plan");
        auto& dd = *pl->dist;
        if (dd.s_epoch == 0) return fail("mgb_dist_s_wait: no mgb_dist_s_publish issued");
        CUDA_OK(cudaSetDevice(pl->ctx->device));
        const mgb::DistGather G = make_dist_gather(pl, nullptr);
        mgb::dist_s_wait_kernel<<<1, 32, 0, pl->ctx->stream>>>(G);
        g_launches++;
        CUDA_OK(cudaGetLastError());
        if (s_full_dev) *s_full_dev = G.full[dd.rank];
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_dist_s_wait: ") + ex.what()); }
}

int mgb_dist_assemble_s(mgb_plan* pl, const double* s_own_dev, const double* Dz0_dev, const double* c_dev, double t,
                        int32_t flags, const double** hval_own_dev, const double** grad_own_dev, const double** scal_dev) {
    const double* s_full = nullptr;
    int rc = mgb_dist_s_publish(pl, s_own_dev);
    if (rc) return rc;
    rc = mgb_dist_s_wait(pl, &s_full);
    if (rc) return rc;
    return mgb_dist_assemble(pl, s_full, Dz0_dev, c_dev, t, flags, hval_own_dev, grad_own_dev, scal_dev);
}

int mgb_dist_begin(mgb_plan* pl, const double* s_dev, const double* Dz0_dev, const double* c_dev, double t, int32_t flags) {
    try {
        if (!pl || !pl->dist || !s_dev || !c_dev) return fail("mgb_dist_begin: NULL argument / not a distributed plan");
        if (!pl->ctx) return fail("mgb_dist_begin: symbolic-only plan; no CPU path exists");
        dist_launch(pl, s_dev, Dz0_dev, c_dev, t, flags, /*fused=*/false);
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_dist_begin: ") + ex.what()); }
}

int mgb_dist_end(mgb_plan* pl, double t, int32_t flags, const double** hval_own_dev, const double** grad_own_dev,
                 const double** scal_dev) {
    try {
        if (!pl || !pl->dist || !pl->ctx) return fail("mgb_dist_end: not a distributed device plan");
        auto& dd = *pl->dist;
        if (!dd.finish_pending) return fail("mgb_dist_end: no mgb_dist_begin in flight");
        (void)flags;
        CUDA_OK(cudaSetDevice(pl->ctx->device));
        if (dd.nranks > 1) {
            mgb::dist_finish_kernel<<<1, 128, 0, pl->ctx->stream>>>(make_dist_scal(pl, false), t, dd.scal.p);
            g_launches++;
            CUDA_OK(cudaGetLastError());
        }
        dd.finish_pending = false;
        dist_outputs(pl, hval_own_dev, grad_own_dev, scal_dev);
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_dist_end: ") + ex.what()); }
}

int mgb_dist_assemble(mgb_plan* pl, const double* s_dev, const double* Dz0_dev, const double* c_dev, double t,
                      int32_t flags, const double** hval_own_dev, const double** grad_own_dev, const double** scal_dev) {
    try {
        if (!pl || !pl->dist || !s_dev || !c_dev) return fail("mgb_dist_assemble: NULL argument / not a distributed plan");
        if (!pl->ctx) return fail("mgb_dist_assemble: symbolic-only plan; no CPU path exists");
        dist_launch(pl, s_dev, Dz0_dev, c_dev, t, flags, /*fused=*/true);
        dist_outputs(pl, hval_own_dev, grad_own_dev, scal_dev);
        return 0;
    } catch (const std::exception& ex) { return fail(std::string("mgb_dist_assemble: ") + ex.what()); }
}

}  // extern "C"
