// General CSR path: the same numeric phase for operators WITHOUT broken-element block structure
// (or operator tables other than [u.id; u.d*; s.id]).  One kernel per seam of the reference:
//   csr_apply_kernel    Dz = Dz0 + E_k s            (apply_D, reference test/test_apply_d.jl:44)
//   csr_barrier_kernel  w.*F1, w.*F2, objective    (map_rows src:161-170 + amgb_diag src:137-147)
//   csr_replay_kernel   gradient g = sum_k E_k' (w.*(y1_k + t c_k)) (gather over the stored transpose) and, in the
//                       same launch, the Hessian: lane per upper-triangle output entry, numeric-only replay of
//                       sum_jk E_j' diag E_k on the frozen pattern from precomputed (coefficient, V index) product
//                       lists, no atomics (test/test_map_rows_compare.jl:102-123 with R folded in: E_k = D_k R)
//   csr_finish_kernel   partial sums of the lists that were cut into chunks + the scalar fold
// The three sparse steps share one list layout (sliced ELL with sorting windows and chunks, see SellHost) and one
// replay routine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "plan_host.h"

namespace mgb {

// sorting window of the SELL lists (outputs); MGB_SELL_SIGMA in the environment overrides it (tuning runs)
static int sell_sigma() {
    const char* e = std::getenv("MGB_SELL_SIGMA");
    const int v = e ? std::atoi(e) : 0;
    return v >= 1 ? v : 16384;
}
// apply_D lists are nearly uniform on fine levels: a short window keeps the Dz0 reads / Dz stores local
static int sell_sigma_apply() { return std::min(sell_sigma(), 256); }
static int sell_sigma_grad() {
    const char* e = std::getenv("MGB_SELL_SIGMA_GRAD");
    const int v = e ? std::atoi(e) : 0;
    return v >= 1 ? v : 1024;
}
// longest run one lane replays (MGB_SELL_CHUNK); longer lists are cut into chunks, see below
static int sell_chunk() {
    const char* e = std::getenv("MGB_SELL_CHUNK");
    const int v = e ? std::atoi(e) : 0;
    return v >= 1 ? v : 16;
}

// Sliced-ELL replay list (SELL-32-sigma).  Every output value (a Dz entry, a gradient entry, an upper-triangle
// Hessian entry) owns a list of (coefficient, source index) contributions that is fixed per level.  Outputs are
// grouped in slices of 32 lanes = one warp; contribution r of lane l of slice s sits at (off[s] + r) * 32 + l, so
// every warp load of coefficients / source ids is one contiguous, fully used run (the thread-per-list CSR walk
// touched 32 different sectors per step and idled the warp on its longest list).  Inside every window of `sigma`
// consecutive outputs the outputs are sorted by list length first (stable), which removes the padding and the
// divergence.  Padding contributions carry src = -1 and are skipped by a predicate - a 0 * Inf of a non-finite
// iterate must not leak into other entries.
//
// Chunks.  A replay is a chain of dependent loads (source id -> value -> fma), and the kernel cannot end before its
// longest chain does: on the fem3d mesh the mean Hessian list has 3.8 products, the longest 144 (a vertex shared by
// eight elements), and that one lane set the kernel time (113 us at 36 % of the DRAM peak, ncu long-scoreboard
// stalls).  Lists longer than `chunk` are therefore cut into runs of `chunk` contributions that different lanes
// replay; each run drops its partial sum into `part`, and a small second kernel adds the partials of an output in
// list order.  code[slot] says what a slot stands for: >= 0 the output itself, -1 padding, <= -2 partial slot
// -(code + 2).  Lists that fit one chunk keep their order and value bit for bit; chunked lists are the fixed-order
// sum of their runs (deterministic, differs by rounding only).
struct SellHost {
    std::vector<uint32_t> off;   // nslices + 1, in units of 32 contributions
    std::vector<int32_t> code;   // nslices * 32
    std::vector<int32_t> src;
    std::vector<double> coef;
    std::vector<int32_t> comb_ptr, comb_out;   // chunked outputs: partial range, output id
};

template <class PtrT>
static void build_sell(int64_t nent, const PtrT* ptr, const double* coef, const int32_t* src, int sigma, int chunk, SellHost& out) {
    // virtual entries: (first contribution, length, code)
    std::vector<int64_t> vbeg;
    std::vector<int32_t> vlen, vcode;
    vbeg.reserve((size_t)nent); vlen.reserve((size_t)nent); vcode.reserve((size_t)nent);
    out.comb_ptr.assign(1, 0);
    int64_t npart = 0;
    for (int64_t e = 0; e < nent; ++e) {
        const int64_t b0 = (int64_t)ptr[e], len = (int64_t)ptr[e + 1] - b0;
        if (len <= chunk) { vbeg.push_back(b0); vlen.push_back((int32_t)len); vcode.push_back((int32_t)e); continue; }
        for (int64_t c0 = 0; c0 < len; c0 += chunk) {
            vbeg.push_back(b0 + c0); vlen.push_back((int32_t)std::min<int64_t>(chunk, len - c0));
            vcode.push_back((int32_t)(-2 - npart)); ++npart;
            if (npart > INT32_MAX - 4) throw std::runtime_error("csr path: partial slots exceed int32");
        }
        out.comb_ptr.push_back((int32_t)npart);
        out.comb_out.push_back((int32_t)e);
    }
    const int64_t nv = (int64_t)vbeg.size();
    const int64_t nsl = (nv + 31) / 32;
    std::vector<int32_t> order((size_t)nsl * 32, -1), win;
    for (int64_t w0 = 0; w0 < nv; w0 += sigma) {
        const int64_t w1 = std::min<int64_t>(nv, w0 + sigma);
        win.resize((size_t)(w1 - w0));
        for (int64_t j = w0; j < w1; ++j) win[(size_t)(j - w0)] = (int32_t)j;
        if (sigma > 1) std::stable_sort(win.begin(), win.end(), [&](int32_t a, int32_t b) { return vlen[(size_t)a] > vlen[(size_t)b]; });
        for (int64_t j = w0; j < w1; ++j) order[(size_t)j] = win[(size_t)(j - w0)];
    }
    out.code.assign((size_t)nsl * 32, -1);
    out.off.assign((size_t)nsl + 1, 0);
    for (int64_t sl = 0; sl < nsl; ++sl) {
        int64_t len = 0;
        for (int l = 0; l < 32; ++l) {
            const int32_t v = order[(size_t)sl * 32 + l];
            if (v < 0) continue;
            out.code[(size_t)sl * 32 + l] = vcode[(size_t)v];
            len = std::max<int64_t>(len, vlen[(size_t)v]);
        }
        const int64_t nxt = (int64_t)out.off[(size_t)sl] + len;
        if (nxt > (int64_t)UINT32_MAX) throw std::runtime_error("csr path: replay list too long for 32-bit slice offsets");
        out.off[(size_t)sl + 1] = (uint32_t)nxt;
    }
    const size_t tot = (size_t)out.off[(size_t)nsl] * 32;
    out.coef.assign(tot, 0.0);
    out.src.assign(tot, -1);
    for (int64_t sl = 0; sl < nsl; ++sl)
        for (int l = 0; l < 32; ++l) {
            const int32_t v = order[(size_t)sl * 32 + l];
            if (v < 0) continue;
            size_t q = (size_t)out.off[(size_t)sl] * 32 + l;
            for (int64_t r = vbeg[(size_t)v]; r < vbeg[(size_t)v] + vlen[(size_t)v]; ++r, q += 32) { out.coef[q] = coef[r]; out.src[q] = src[r]; }
        }
}

// contributions a SELL list stores, padding included (plan statistics; no arrays are built)
template <class PtrT>
static int64_t sell_stored(int64_t nent, const PtrT* ptr, int sigma) {
    std::vector<int64_t> len;
    int64_t tot = 0;
    for (int64_t w0 = 0; w0 < nent; w0 += std::max(sigma, 32)) {
        const int64_t w1 = std::min<int64_t>(nent, w0 + std::max(sigma, 32));
        len.clear();
        for (int64_t j = w0; j < w1; ++j) len.push_back((int64_t)(ptr[j + 1] - ptr[j]));
        if (sigma > 1) std::sort(len.begin(), len.end(), std::greater<int64_t>());
        for (size_t q = 0; q < len.size(); q += 32) tot += 32 * *std::max_element(len.begin() + q, len.begin() + std::min(len.size(), q + 32));
    }
    return tot;
}

struct SellDev {
    const uint32_t* off = nullptr;
    const int32_t* code = nullptr;
    const int32_t* src = nullptr;
    const double* coef = nullptr;
    int64_t nslot = 0;   // slices * 32
    const int32_t* comb_ptr = nullptr;   // chunked outputs (ncomb + 1)
    const int32_t* comb_out = nullptr;
    int64_t ncomb = 0;
    double* part = nullptr;
};

struct CsrDev {
    int ND = 0, npair = 0;
    int pair_a[36] = {0}, pair_b[36] = {0};
    int64_t nloc = 0, m = 0, nnzH = 0, nup = 0;
    BarrierDesc bar;
    std::vector<void*> owned;  // device allocations
    SellDev E[8];              // apply_D: outputs = local rows of E_k = D_k R, sources = unknowns
    SellDev G;                 // gradient: outputs = unknowns, sources = gy entries (k * nloc + i)
    SellDev Hs;                // Hessian: outputs = upper-triangle entries, sources = V entries (pair * nloc + i)
    const int32_t* up_t = nullptr;   // per slot: position of the entry in the CSR value array (-1: padding slot,
                                     // <= -2: partial slot of a chunked entry)
    const int32_t* up_m = nullptr;   // per slot: position of the mirror entry (-1: diagonal / padding / partial)
    const int32_t* nat_t = nullptr;  // the two positions per upper entry in CSR order (chunked entries)
    const int32_t* nat_m = nullptr;
    double* Dz = nullptr;    // nloc x ND
    double* gy = nullptr;    // nloc x ND
    double* V = nullptr;     // nloc x npair
    double* part = nullptr;  // nblk x 4
    int64_t nblk = 0;
    ~CsrDev() { for (void* p : owned) cudaFree(p); }
};

template <class T>
static const T* csr_up(CsrDev& d, const std::vector<T>& h, cudaStream_t st, size_t& bytes) {
    T* p = nullptr;
    const size_t nb = std::max<size_t>(h.size(), 1) * sizeof(T);
    if (cudaMalloc(&p, nb) != cudaSuccess) throw std::runtime_error("cudaMalloc failed (csr plan)");
    d.owned.push_back(p);
    // the host vectors die before the stream is synchronised by the caller: copy synchronously
    if (!h.empty() && cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess)
        throw std::runtime_error("cudaMemcpy failed (csr plan)");
    (void)st;
    bytes += nb;
    return p;
}

static SellDev sell_upload(CsrDev& d, const SellHost& h, cudaStream_t st, size_t& bytes) {
    SellDev o;
    o.off = csr_up(d, h.off, st, bytes);
    o.code = csr_up(d, h.code, st, bytes);
    o.src = csr_up(d, h.src, st, bytes);
    o.coef = csr_up(d, h.coef, st, bytes);
    o.nslot = (int64_t)h.code.size();
    o.comb_ptr = csr_up(d, h.comb_ptr, st, bytes);
    o.comb_out = csr_up(d, h.comb_out, st, bytes);
    o.ncomb = (int64_t)h.comb_out.size();
    double* p = nullptr;
    const size_t np = (size_t)std::max<int32_t>(h.comb_ptr.back(), 1);
    if (cudaMalloc(&p, np * 8) != cudaSuccess) throw std::runtime_error("cudaMalloc failed (replay partials)");
    d.owned.push_back(p);
    bytes += np * 8;
    o.part = p;
    return o;
}

static size_t csr_upload(const CsrPlan& cp, const BarrierDesc& bar, CsrDev& d, cudaStream_t st) {
    size_t bytes = 0;
    d.ND = cp.ND; d.nloc = cp.nloc; d.m = cp.m; d.nnzH = (int64_t)cp.h_colidx.size();
    d.nup = (int64_t)cp.up_t.size(); d.bar = bar;
    d.npair = cp.npair;
    for (int c = 0; c < cp.npair; ++c) { d.pair_a[c] = cp.pair_a[c]; d.pair_b[c] = cp.pair_b[c]; }
    for (int k = 0; k < cp.ND; ++k) {
        SellHost h;
        build_sell(cp.nloc, cp.E[k].ptr.data(), cp.E[k].val.data(), cp.E[k].idx.data(), sell_sigma_apply(), sell_chunk(), h);
        d.E[k] = sell_upload(d, h, st, bytes);
    }
    {
        SellHost h;
        build_sell(cp.m, cp.gt_ptr.data(), cp.gt_coef.data(), cp.gt_src.data(), sell_sigma_grad(), sell_chunk(), h);
        d.G = sell_upload(d, h, st, bytes);
    }
    {
        SellHost h;
        build_sell(d.nup, cp.prod_ptr.data(), cp.prod_coef.data(), cp.prod_v.data(), sell_sigma(), sell_chunk(), h);
        d.Hs = sell_upload(d, h, st, bytes);
        std::vector<int32_t> t(h.code.size(), -1), mm(h.code.size(), -1);
        for (size_t q = 0; q < h.code.size(); ++q) {
            if (h.code[q] >= 0) { t[q] = cp.up_t[(size_t)h.code[q]]; mm[q] = cp.up_m[(size_t)h.code[q]]; }
            else t[q] = h.code[q];
        }
        d.up_t = csr_up(d, t, st, bytes);
        d.up_m = csr_up(d, mm, st, bytes);
        d.nat_t = csr_up(d, cp.up_t, st, bytes);
        d.nat_m = csr_up(d, cp.up_m, st, bytes);
    }
    auto scratch = [&](size_t count) {
        double* p = nullptr;
        if (cudaMalloc(&p, std::max<size_t>(count, 1) * 8) != cudaSuccess) throw std::runtime_error("cudaMalloc failed (csr scratch)");
        d.owned.push_back(p);
        bytes += count * 8;
        return p;
    };
    d.Dz = scratch((size_t)cp.nloc * cp.ND);
    d.gy = scratch((size_t)cp.nloc * cp.ND);
    d.V = scratch((size_t)cp.nloc * std::max(cp.npair, 1));
    d.nblk = (cp.nloc + 255) / 256;
    d.part = scratch((size_t)d.nblk * 4);
    return bytes;
}

// one slot's replay: sum of coef * x[src] over the slice's rows, in list order.  
__device__ __forceinline__ double sell_replay(const SellDev& L, const double* __restrict__ x, const int64_t slot) {
    const int64_t sl = slot >> 5;
    const uint32_t o0 = __ldg(&L.off[sl]), o1 = __ldg(&L.off[sl + 1]);
    int64_t q = (int64_t)o0 * 32 + (slot & 31);
    double acc = 0.0;
#pragma unroll 4
    for (uint32_t r = o0; r < o1; ++r, q += 32) {
        const int32_t v = __ldg(&L.src[q]);
        const double c = __ldg(&L.coef[q]);
        if (v >= 0) acc = fma(c, x[v], acc);
    }
    return acc;
}
// partials of chunked output q, added in list order
__device__ __forceinline__ double sell_combine(const SellDev& L, const int64_t q) {
    const int32_t r0 = __ldg(&L.comb_ptr[q]), r1 = __ldg(&L.comb_ptr[q + 1]);
    double acc = 0.0;
#pragma unroll 4
    for (int32_t r = r0; r < r1; ++r) acc += __ldcs(&L.part[r]);
    return acc;
}

struct CsrApplyParams {
    SellDev E[8];
    int ND;
    int64_t n;
    const double* s;
    const double* Dz0;
    double* Dz;
};

// apply_D: blockIdx.y = operator, one slot per local row or chunk of a row
__global__ void __launch_bounds__(256) csr_apply_kernel(const __grid_constant__ CsrApplyParams P) {
    pdl_launch_dependents();
    const int k = blockIdx.y;
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= P.E[k].nslot) return;
    const double dot = sell_replay(P.E[k], P.s, slot);
    const int32_t i = __ldg(&P.E[k].code[slot]);
    if (i >= 0) {
        const int64_t o = (int64_t)k * P.n + i;
        P.Dz[o] = P.Dz0 ? P.Dz0[o] + dot : dot;
    } else if (i <= -2) P.E[k].part[-(i + 2)] = dot;
}
__global__ void __launch_bounds__(256) csr_apply_combine_kernel(const __grid_constant__ CsrApplyParams P) {
    const int k = blockIdx.y;
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P.E[k].ncomb) return;
    const double dot = sell_combine(P.E[k], q);
    const int64_t o = (int64_t)k * P.n + __ldg(&P.E[k].comb_out[q]);
    P.Dz[o] = P.Dz0 ? P.Dz0[o] + dot : dot;
}

struct CsrBarrierParams {
    int ND, nq, slack;
    int idx[8];
    int nq2;       // second cone: number of q columns (-1: no second cone)
    int idx2[8];
    double p2;
    int npair;
    signed char pair_a[36], pair_b[36];   // operator pair of every V column (unique pairs coupled by a cone)
    signed char loc[2][8];                // per cone: operator column -> local slot (0..2 q, 3 s, 4 slack) or -1
    int64_t n;
    double p, t;
    const double* Dz;
    const double* c;
    const double* w;
    double* gy;
    double* V;
    double* part;
    int want_f, want_g, want_h;
};

// entry (la, lb) of the 5x5 local Hessian of one cone over (q0, q1, q2, s, slack); la, lb are warp-uniform (they come
// from the kernel parameters), so the switch is a short tree of uniform branches and the barrier derivatives stay
// in registers (a chain of selects over all 25 combinations made this kernel instruction bound)
__device__ __forceinline__ double cone_hess_pick(const BarrierOut& bo, const double tt, int la, int lb) {
    if (la > lb) { const int x = la; la = lb; lb = x; }
    switch (la * 5 + lb) {
        case 0: return bo.Hqq[0][0];
        case 1: return bo.Hqq[0][1];
        case 2: return bo.Hqq[0][2];
        case 3: case 4: return bo.Hqs[0];
        case 6: return bo.Hqq[1][1];
        case 7: return bo.Hqq[1][2];
        case 8: case 9: return bo.Hqs[1];
        case 12: return bo.Hqq[2][2];
        case 13: case 14: return bo.Hqs[2];
        case 18: case 19: return bo.Hss;
        case 24: return bo.Hss + tt;
        default: return 0.0;
    }
}
__device__ __forceinline__ double cone_grad_pick(const BarrierOut& bo, const double itau, const int la) {
    switch (la) {
        case 0: return bo.gq[0];
        case 1: return bo.gq[1];
        case 2: return bo.gq[2];
        case 3: return bo.gs;
        case 4: return bo.gs - itau;
        default: return 0.0;
    }
}

// map_rows of F / F1 / F2 over the Dz rows for one or two cones: everything of a point stays in registers, every
// output column (w.*F1: ND columns, w.*F2: one column per coupled operator pair) is stored exactly once.
template <int NCONES>
__global__ void __launch_bounds__(256) csr_barrier_kernel(const __grid_constant__ CsrBarrierParams P) {
    pdl_launch_dependents();   // the replay kernel may stage its list metadata while this grid drains
    pdl_wait_primary();        // Dz comes from the apply kernel
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool act = i < P.n;
    const int64_t n = P.n;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
    if (act) {
        const double wi = P.w[i];
        const int ND = P.ND;
        double cd = 0.0;
        for (int k = 0; k < ND; ++k) cd = fma(P.c[(int64_t)k * n + i], P.Dz[(int64_t)k * n + i], cd);
        v1 = wi * cd;
        BarrierOut bo[NCONES];
        double itau = 0.0, tt = 0.0, Fsum = 0.0;
        bool feas = true;
        constexpr int ncones = NCONES;
#pragma unroll
        for (int cone = 0; cone < NCONES; ++cone) {
            const int nq = cone ? P.nq2 : P.nq;
            const int* idx = cone ? P.idx2 : P.idx;
            const double pp = cone ? P.p2 : P.p;
            double q[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int j = 0; j < 3; ++j)
                if (j < nq) q[j] = P.Dz[(int64_t)idx[j] * n + i];
            double s = P.Dz[(int64_t)idx[nq] * n + i];
            const bool sl = P.slack && cone == 0;
            if (sl) s += P.Dz[(int64_t)(ND - 1) * n + i];
            barrier_eval<3, true, true>(q, s, pp, bo[cone]);
            if (sl) {  // slack bounded below by -log(1 + tau)
                const double tau1 = 1.0 + P.Dz[(int64_t)(ND - 1) * n + i];
                bo[cone].F = (tau1 > 0.0) ? bo[cone].F - log(tau1) : __longlong_as_double(0x7ff0000000000000LL);
                bo[cone].feasible = bo[cone].feasible && (tau1 > 0.0);
                itau = 1.0 / tau1;
                tt = itau * itau;
            }
            Fsum += bo[cone].F;
            feas = feas && bo[cone].feasible;
        }
        if (P.want_g) {
            for (int k = 0; k < ND; ++k) {
                double g = cone_grad_pick(bo[0], itau, P.loc[0][k]);
                if (NCONES == 2) g += cone_grad_pick(bo[NCONES - 1], 0.0, P.loc[1][k]);
                P.gy[(int64_t)k * n + i] = wi * (g + P.t * P.c[(int64_t)k * n + i]);
            }
        }
        if (P.want_h) {
            for (int cidx = 0; cidx < P.npair; ++cidx) {
                const int ka = P.pair_a[cidx], kb = P.pair_b[cidx];
                double h = 0.0;
                if (P.loc[0][ka] >= 0 && P.loc[0][kb] >= 0) h = cone_hess_pick(bo[0], tt, P.loc[0][ka], P.loc[0][kb]);
                if (NCONES == 2 && P.loc[1][ka] >= 0 && P.loc[1][kb] >= 0) h += cone_hess_pick(bo[NCONES - 1], 0.0, P.loc[1][ka], P.loc[1][kb]);
                P.V[(int64_t)cidx * n + i] = wi * h;
            }
        }
        v0 = P.want_f ? wi * Fsum : 0.0;
        v2 = feas ? 0.0 : 1.0;
    }
#pragma unroll
    for (int mk = 16; mk >= 1; mk >>= 1) {
        v0 += shfl_xor_d(v0, mk);
        v1 += shfl_xor_d(v1, mk);
        v2 += shfl_xor_d(v2, mk);
    }
    __shared__ double red[3][8];
    const int wid = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { red[0][wid] = v0; red[1][wid] = v1; red[2][wid] = v2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int r = 0; r < 8; ++r) { s0 += red[0][r]; s1 += red[1][r]; s2 += red[2][r]; }
        P.part[(int64_t)blockIdx.x * 4 + 0] = s0;
        P.part[(int64_t)blockIdx.x * 4 + 1] = s1;
        P.part[(int64_t)blockIdx.x * 4 + 2] = s2;
    }
}

// Gradient and Hessian replays share one launch (both depend on the barrier kernel only): the gradient blocks come
// first - their lists are longer - and their latency hides inside the Hessian replay, which is most of the step.
//   gradient: g[a] = sum_r coef[r] * gy[src[r]], the transposed operators merged into one list per unknown
//             (gather, no atomics, fixed order)
//   Hessian : numeric-only triple product on the frozen pattern: one lane per UPPER-triangle output entry (or chunk
//             of one); it sums its precomputed products coef * V in list order (no atomics, bit-reproducible) and
//             stores the value at (a,b) and at the mirror (b,a) - R'HR is symmetric, so half of the product lists
//             never has to be read.  coef = E_ka[i,a]*E_kb[i,b] is level data, V = w.*F2 changes every Newton step.
struct CsrReplayParams {
    SellDev G;
    const double* gy;
    double* grad;
    int64_t nblk_g;       // 0: no gradient
    SellDev H;
    const int32_t* up_t;
    const int32_t* up_m;
    const double* V;
    double* hval;
};

__global__ void __launch_bounds__(256) csr_replay_kernel(const __grid_constant__ CsrReplayParams P) {
    // programmatic dependent launch: this grid may start while the barrier kernel drains; everything loaded before
    // pdl_wait_primary() is level data (plan arrays), gy / V are read after it.  Every thread reaches the wait.
    pdl_launch_dependents();
    if ((int64_t)blockIdx.x < P.nblk_g) {
        const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const bool act = slot < P.G.nslot;
        const int32_t a = act ? __ldg(&P.G.code[slot]) : -1;
        pdl_wait_primary();
        if (!act) return;
        const double acc = sell_replay(P.G, P.gy, slot);
        if (a >= 0) P.grad[a] = acc;
        else if (a <= -2) P.G.part[-(a + 2)] = acc;
        return;
    }
    const int64_t slot = ((int64_t)blockIdx.x - P.nblk_g) * blockDim.x + threadIdx.x;
    const bool act = slot < P.H.nslot;
    const int32_t t = act ? __ldg(&P.up_t[slot]) : -1, tm = act ? __ldg(&P.up_m[slot]) : -1;
    pdl_wait_primary();
    if (!act) return;
    const double acc = sell_replay(P.H, P.V, slot);
    if (t >= 0) {
        P.hval[t] = acc;
        if (tm >= 0) P.hval[tm] = acc;
    } else if (t <= -2) P.H.part[-(t + 2)] = acc;
}

// Last launch of an assembly: partial sums of the chunked gradient lists, of the chunked Hessian lists, and the
// fold of the per-block scalars (last block) - one launch instead of three small ones.
struct CsrFinishParams {
    SellDev G;
    double* grad;
    int64_t nblk_g;
    SellDev H;
    const int32_t* nat_t;
    const int32_t* nat_m;
    double* hval;
    int64_t nblk_h;
    const double* part;
    int64_t nparts;
    double t;
    double* scal;
};

__global__ void __launch_bounds__(256) csr_finish_kernel(const __grid_constant__ CsrFinishParams P) {
    pdl_wait_primary();   // partial sums and scalar partials come from the replay / barrier kernels
    const int64_t b = blockIdx.x;
    if (b < P.nblk_g) {
        const int64_t q = b * blockDim.x + threadIdx.x;
        if (q < P.G.ncomb) P.grad[__ldg(&P.G.comb_out[q])] = sell_combine(P.G, q);
    } else if (b < P.nblk_g + P.nblk_h) {
        const int64_t q = (b - P.nblk_g) * blockDim.x + threadIdx.x;
        if (q >= P.H.ncomb) return;
        const int32_t e = __ldg(&P.H.comb_out[q]);
        const int32_t t = __ldg(&P.nat_t[e]), tm = __ldg(&P.nat_m[e]);
        const double acc = sell_combine(P.H, q);
        P.hval[t] = acc;
        if (tm >= 0) P.hval[tm] = acc;
    } else {
        fold_scalars_block(P.part, P.nparts, P.t, P.scal);
    }
}

static void csr_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

// launch as a programmatic dependent of the previous kernel in the stream (its blocks may be scheduled while the
// previous grid drains and block at griddepcontrol.wait); plain launch when the attribute is refused or MGB_NO_PDL is set
template <class Params>
static void csr_launch(void (*kernel)(Params), unsigned grid, cudaStream_t st, const Params& params, bool pdl) {
    static bool pdl_ok = std::getenv("MGB_NO_PDL") == nullptr;
    if (pdl && pdl_ok) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (cudaLaunchKernelEx(&cfg, kernel, params) == cudaSuccess) return;
        cudaGetLastError();
        pdl_ok = false;
    }
    kernel<<<grid, 256, 0, st>>>(params);
}

static unsigned sell_grid(int64_t count) { return (unsigned)((count + 255) / 256); }

// returns the number of kernels launched
static int csr_assemble(CsrDev& d, const double* w, const double* s, const double* Dz0, const double* c, double t,
                        int flags, double* scal, double* grad, double* hval, double* Dz_out, cudaStream_t st) {
    int launches = 0;
    const int64_t n = d.nloc;
    double* Dz = ((flags & 8) && Dz_out) ? Dz_out : d.Dz;
    if ((flags & 2) && !grad) throw std::runtime_error("MGB_WANT_GRAD without grad buffer");
    if ((flags & 4) && !hval) throw std::runtime_error("MGB_WANT_HESS without hval buffer");
    {
        CsrApplyParams P{};
        int64_t nslot = 0, ncomb = 0;
        for (int k = 0; k < d.ND; ++k) { P.E[k] = d.E[k]; nslot = std::max(nslot, d.E[k].nslot); ncomb = std::max(ncomb, d.E[k].ncomb); }
        P.ND = d.ND; P.n = n; P.s = s; P.Dz0 = Dz0; P.Dz = Dz;
        if (nslot > 0) {
            csr_apply_kernel<<<dim3(sell_grid(nslot), (unsigned)d.ND), 256, 0, st>>>(P);
            ++launches;
        }
        if (ncomb > 0) {
            csr_apply_combine_kernel<<<dim3(sell_grid(ncomb), (unsigned)d.ND), 256, 0, st>>>(P);
            ++launches;
        }
    }
    {
        CsrBarrierParams P{};
        P.ND = d.ND; P.nq = d.bar.nidx - 1; P.slack = d.bar.slack;
        for (int j = 0; j < d.bar.nidx; ++j) P.idx[j] = d.bar.idx[j];
        P.nq2 = d.bar.nidx2 > 0 ? d.bar.nidx2 - 1 : -1; P.p2 = d.bar.p2;
        for (int j = 0; j < d.bar.nidx2; ++j) P.idx2[j] = d.bar.idx2[j];
        P.npair = d.npair;
        for (int c2 = 0; c2 < d.npair; ++c2) { P.pair_a[c2] = (signed char)d.pair_a[c2]; P.pair_b[c2] = (signed char)d.pair_b[c2]; }
        for (int cone = 0; cone < 2; ++cone)
            for (int k = 0; k < 8; ++k) P.loc[cone][k] = -1;
        for (int j = 0; j < P.nq; ++j) P.loc[0][P.idx[j]] = (signed char)j;
        P.loc[0][P.idx[P.nq]] = 3;
        if (P.slack) P.loc[0][d.ND - 1] = 4;
        if (P.nq2 >= 0) {
            for (int j = 0; j < P.nq2; ++j) P.loc[1][P.idx2[j]] = (signed char)j;
            P.loc[1][P.idx2[P.nq2]] = 3;
        }
        P.n = n; P.p = d.bar.p; P.t = t; P.Dz = Dz; P.c = c; P.w = w; P.gy = d.gy; P.V = d.V; P.part = d.part;
        P.want_f = (flags & 1) ? 1 : 0; P.want_g = (flags & 2) ? 1 : 0; P.want_h = (flags & 4) ? 1 : 0;
        if (P.nq2 >= 0) csr_launch(csr_barrier_kernel<2>, (unsigned)d.nblk, st, P, true);
        else csr_launch(csr_barrier_kernel<1>, (unsigned)d.nblk, st, P, true);
        ++launches;
    }
    const bool want_g = (flags & 2) && d.G.nslot > 0, want_h = (flags & 4) && d.nup > 0;
    if (want_g || want_h) {
        CsrReplayParams P{};
        if (want_g) { P.G = d.G; P.gy = d.gy; P.grad = grad; P.nblk_g = sell_grid(d.G.nslot); }
        if (want_h) { P.H = d.Hs; P.up_t = d.up_t; P.up_m = d.up_m; P.V = d.V; P.hval = hval; }
        const int64_t nblk_h = want_h ? sell_grid(d.Hs.nslot) : 0;
        csr_launch(csr_replay_kernel, (unsigned)(P.nblk_g + nblk_h), st, P, true);
        ++launches;
    }
    {
        CsrFinishParams P{};
        if (want_g && d.G.ncomb > 0) { P.G = d.G; P.grad = grad; P.nblk_g = sell_grid(d.G.ncomb); }
        if (want_h && d.Hs.ncomb > 0) { P.H = d.Hs; P.nat_t = d.nat_t; P.nat_m = d.nat_m; P.hval = hval; P.nblk_h = sell_grid(d.Hs.ncomb); }
        P.part = d.part; P.nparts = d.nblk; P.t = t; P.scal = scal;
        csr_launch(csr_finish_kernel, (unsigned)(P.nblk_g + P.nblk_h + 1), st, P, true);
        ++launches;
    }
    csr_check(cudaGetLastError(), "csr_assemble launch");
    return launches;
}

static int csr_map_barrier(const BarrierDesc& bar, int ND, int64_t n, const double* Dz, int which, double* out,
                           cudaStream_t st) {
    // only the contiguous idx = (1..d+1) layout is exposed through this seam
    const int d = bar.nidx - 1;
    for (int j = 0; j < bar.nidx; ++j)
        if (bar.idx[j] != 1 + j) throw std::runtime_error("map_barrier: idx must be (1..d+1)");
    const unsigned nb = (unsigned)((n + 255) / 256);
    if (d == 1) map_barrier_kernel<1><<<nb, 256, 0, st>>>(Dz, n, ND, bar.slack, bar.p, which, out);
    else if (d == 2) map_barrier_kernel<2><<<nb, 256, 0, st>>>(Dz, n, ND, bar.slack, bar.p, which, out);
    else if (d == 3) map_barrier_kernel<3><<<nb, 256, 0, st>>>(Dz, n, ND, bar.slack, bar.p, which, out);
    else throw std::runtime_error("map_barrier: unsupported number of derivative columns");
    csr_check(cudaGetLastError(), "map_barrier launch");
    return 1;
}

}  // namespace mgb
