// Multi-GPU numeric phase (one process per GPU): the gather stage fused with the exchange.
//
// Every rank runs element_kernel on its own quadrature rows (whole elements, so apply_D needs no halo),
// then push_kernel replays its frozen contribution lists exactly like gather_kernel does on one GPU -
// but each result is stored straight into the exchange window of the rank that OWNS the output row
// (HPCSparseArrays row partition), through NVLink peer-mapped memory: entries fed by this
// rank alone land in their final position of the owner's CSR value array / gradient block, entries on the
// element-partition interface land in the owner's staging area.  The last CTA to retire publishes an epoch
// flag in every peer's window (release, system scope), then - fused mode, mgb_dist_assemble - acts as the
// owner: waits for all peers' flags (acquire), sums the staged values in source-rank order
// (bit-reproducible) and folds the scalars.  Split mode (mgb_dist_begin / mgb_dist_end) runs that owner
// step as finish_kernel instead, so several ranks can share one stream (tests on a single GPU).
// No NCCL on the data path; no atomics on values.  Windows are double-buffered by epoch parity: a rank can
// only reach epoch k+2 after every peer finished reading epoch k.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"
#include "plan_host.h"

namespace mgb {

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct FinishParams {
    double* win;                        // this rank's window (this epoch's parity)
    const unsigned long long* flag;     // this rank's flag array (written by the peers)
    int nranks;
    unsigned long long epoch, timeout_ns;
    int64_t n_fh, n_fg;
    const int32_t* fh_pos; const int32_t* fh_ptr;
    const int32_t* fg_pos; const int32_t* fg_ptr;
    int64_t off_h, off_g, off_scal, off_stg_h, off_stg_g, off_stg_scal;
    double t;
    int* err;                           // set to 1 when a peer's flag did not arrive in time
};

// out[pos[j]] = sum of stg[ptr[j] .. ptr[j+1]) in list (= source-rank) order.  FIN_U entries per thread with all
// index loads, then all value loads, issued together: the fused finish runs in ONE CTA, so memory-level
// parallelism per thread is what bounds it.
constexpr int FIN_U = 4;
constexpr int FIN_CTAS = 8;   // CTAs that share the fused owner-side finish
__device__ __forceinline__ void finish_family(const double* stg, double* out, const int32_t* __restrict__ pos,
                                              const int32_t* __restrict__ ptr, const int64_t n, const int bid, const int nb) {
    for (int64_t base = (int64_t)bid * blockDim.x * FIN_U; base < n; base += (int64_t)nb * blockDim.x * FIN_U) {
        int r0[FIN_U], r1[FIN_U], ps[FIN_U];
#pragma unroll
        for (int u = 0; u < FIN_U; ++u) {
            const int64_t j = base + (int64_t)u * blockDim.x + threadIdx.x;
            const bool ok = j < n;
            r0[u] = ok ? __ldg(&ptr[j]) : 0;
            r1[u] = ok ? __ldg(&ptr[j + 1]) : 0;
            ps[u] = ok ? __ldg(&pos[j]) : -1;
        }
        double a[FIN_U], b[FIN_U];
#pragma unroll
        for (int u = 0; u < FIN_U; ++u) {
            a[u] = (r0[u] < r1[u]) ? __ldcg(&stg[r0[u]]) : 0.0;
            b[u] = (r0[u] + 1 < r1[u]) ? __ldcg(&stg[r0[u] + 1]) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < FIN_U; ++u) {
            if (ps[u] < 0) continue;
            double acc = a[u] + b[u];
            for (int r = r0[u] + 2; r < r1[u]; ++r) acc += __ldcg(&stg[r]);  // more than two source ranks: mesh corners
            out[ps[u]] = acc;
        }
    }
}

// Owner side.  CTA `bid` of `nb`: wait until every rank (this one included) has published `epoch`, then sum
// the staged interface values in source-rank order and fold the scalars.
__device__ __forceinline__ void finish_body(const FinishParams& P, const int bid, const int nb, unsigned long long* dbg = nullptr) {
    // the index lists are level data: pull them into L2 while the peers' flags are still on their way
    for (int64_t j = ((int64_t)bid * blockDim.x + threadIdx.x) * 32; j < P.n_fh; j += (int64_t)nb * blockDim.x * 32) {
        prefetch_l2(P.fh_ptr + j); prefetch_l2(P.fh_pos + j);
    }
    for (int64_t j = ((int64_t)bid * blockDim.x + threadIdx.x) * 32; j < P.n_fg; j += (int64_t)nb * blockDim.x * 32) {
        prefetch_l2(P.fg_ptr + j); prefetch_l2(P.fg_pos + j);
    }
    if ((int)threadIdx.x < P.nranks) {
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(P.flag + threadIdx.x) < P.epoch) {
            if (global_timer_ns() - t0 > P.timeout_ns) { atomicExch(P.err, 1); break; }
        }
    }
    __syncthreads();
    if (dbg && threadIdx.x == 0) dbg[4] = global_timer_ns();
    // staged values were written by peers: L2-coherent loads, never the read-only (nc) path
    finish_family(P.win + P.off_stg_h, P.win + P.off_h, P.fh_pos, P.fh_ptr, P.n_fh, bid, nb);
    finish_family(P.win + P.off_stg_g, P.win + P.off_g, P.fg_pos, P.fg_ptr, P.n_fg, bid, nb);
    if (bid == 0 && threadIdx.x == 0) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int r = 0; r < P.nranks; ++r) {  // rank order: identical result on every rank
            const double* d = P.win + P.off_stg_scal + 4 * r;
            s0 += __ldcg(d + 0); s1 += __ldcg(d + 1); s2 += __ldcg(d + 2);
        }
        double* scal = P.win + P.off_scal;
        scal[0] = s0 + P.t * s1;
        scal[1] = (s2 == 0.0) ? 1.0 : 0.0;
        scal[2] = s1;
        scal[3] = s2;
    }
}

struct PushParams {
    GatherParams G;                        // local replay lists (hval / grad / scal members unused)
    const int32_t* h_dest;                 // per local Hessian entry: (owner << 27) | offset in the owner's window
    const int2* g_dest;                    // per unknown this rank contributes to: {unknown, destination}
    int64_t n_gtouch;
    double* win[DIST_MAX_RANKS];           // peers' windows (this epoch's parity), win[rank] = own
    unsigned long long* flag[DIST_MAX_RANKS];  // peers' flag arrays; this rank writes flag[p][rank]
    int64_t scal_off[DIST_MAX_RANKS];      // offset of this rank's 4 staged scalars inside window p
    int rank, nranks;
    unsigned long long epoch;
    unsigned int* counter;                 // CTA retirement counter (zero between launches)
    int64_t h_rot;                         // rotation of the Hessian block order: [higher ranks][lower ranks][own entries]
    int64_t h_loc_blk;                     // CTAs' worth of own entries: spread evenly between the remote CTAs so
                                           // NVLink-bound and HBM-bound CTAs are resident together
    int fused;                             // the last CTA also runs the owner-side finish (F)
    FinishParams F;
    unsigned long long* dbg;               // optional timeline slots of this epoch (MGB_DIST_DEBUG), else null
};

// timeline slots (globaltimer ns): 0 first CTA start (min), 1 last CTA's stores issued (max), 2 last ticket taken,
// 3 flags published, 4 all flags seen, 5 finish done, 6 stamp before the element kernel
static __global__ void stamp_kernel(unsigned long long* slot) { *slot = global_timer_ns(); }

__device__ __forceinline__ double* dist_dst(const PushParams& P, int32_t d) {
    return P.win[d >> DIST_RANK_SHIFT] + (d & DIST_OFF_MASK);
}

static __global__ void __launch_bounds__(256, 6) push_kernel(const __grid_constant__ PushParams P) {
    const GatherParams& G = P.G;
    const int64_t b = blockIdx.x;
    if (P.dbg && threadIdx.x == 0) atomicMin(&P.dbg[0], global_timer_ns());
    if (b < G.nblk_h) {
        const int64_t l0 = b * P.h_loc_blk / G.nblk_h, l1 = (b + 1) * P.h_loc_blk / G.nblk_h;
        const int64_t pos = (l1 > l0) ? (G.nblk_h - P.h_loc_blk + l0) : (b - l0);
        const int64_t bb = (pos + P.h_rot < G.nblk_h) ? pos + P.h_rot : pos + P.h_rot - G.nblk_h;
        const int64_t base = bb * (256 * GATHER_UNROLL) + threadIdx.x;
        int2 src[GATHER_UNROLL];
        int32_t dst[GATHER_UNROLL];
#pragma unroll
        for (int j = 0; j < GATHER_UNROLL; ++j) {
            const int64_t t = base + (int64_t)j * 256;
            src[j] = (t < G.nnzH) ? __ldg(&G.h_src2[t]) : make_int2(-1, -1);
            dst[j] = (t < G.nnzH) ? __ldg(&P.h_dest[t]) : 0;
        }
        pdl_wait_primary();
        double v0[GATHER_UNROLL], v1[GATHER_UNROLL];
#pragma unroll
        for (int j = 0; j < GATHER_UNROLL; ++j) {
            v0[j] = (src[j].x >= 0) ? G.sel[src[j].x] : 0.0;
            v1[j] = (src[j].y >= 0) ? G.sel[src[j].y] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < GATHER_UNROLL; ++j) {
            if (src[j].x < 0) continue;  // out of range, or a long entry (finished by its own block)
            *dist_dst(P, dst[j]) = v0[j] + v1[j];
        }
    } else if (b < G.nblk_h + G.nblk_l) {
        const int64_t li = (b - G.nblk_h) * 256 + threadIdx.x;
        pdl_wait_primary();
        if (li < G.n_long) {
            const int64_t c0 = __ldg(&G.h_lptr[li]), c1 = __ldg(&G.h_lptr[li + 1]);
            double acc = 0.0;
            for (int64_t cix = c0; cix < c1; ++cix) acc += G.sel[__ldg(&G.h_lidx[cix])];
            *dist_dst(P, __ldg(&P.h_dest[__ldg(&G.h_lt[li])])) = acc;
        }
    } else if (b < G.nblk_h + G.nblk_l + G.nblk_g) {
        const int64_t a = (b - G.nblk_h - G.nblk_l) * 256 + threadIdx.x;
        pdl_wait_primary();
        if (a < P.n_gtouch) {
            const int2 ad = __ldg(&P.g_dest[a]);
            const int64_t c0 = __ldg(&G.g_cptr[ad.x]), c1 = __ldg(&G.g_cptr[ad.x + 1]);
            double acc = 0.0;
            for (int64_t cix = c0; cix < c1; ++cix) acc += G.rel[__ldg(&G.g_cidx[cix])];
            *dist_dst(P, ad.y) = acc;
        }
    } else {
        // this rank's scalar partials {sum w F, <c,Dz>_w, infeasible count} -> every rank's staging row
        pdl_wait_primary();
        __shared__ double sh[3][256];
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int64_t r = threadIdx.x; r < G.nparts; r += blockDim.x) {
            s0 += G.part[r * 4 + 0];
            s1 += G.part[r * 4 + 1];
            s2 += G.part[r * 4 + 2];
        }
        sh[0][threadIdx.x] = s0; sh[1][threadIdx.x] = s1; sh[2][threadIdx.x] = s2;
        __syncthreads();
        for (int st = blockDim.x / 2; st >= 1; st >>= 1) {
            if ((int)threadIdx.x < st) {
                sh[0][threadIdx.x] += sh[0][threadIdx.x + st];
                sh[1][threadIdx.x] += sh[1][threadIdx.x + st];
                sh[2][threadIdx.x] += sh[2][threadIdx.x + st];
            }
            __syncthreads();
        }
        if ((int)threadIdx.x < P.nranks) {
            double* d = P.win[threadIdx.x] + P.scal_off[threadIdx.x];
            d[0] = sh[0][0]; d[1] = sh[1][0]; d[2] = sh[2][0]; d[3] = 0.0;
        }
    }
    // ---- publish: the last CTA to retire raises this rank's epoch flag in every peer's window
    __shared__ int s_slot;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (P.dbg) atomicMax(&P.dbg[1], global_timer_ns());
        // release at GPU scope: this CTA's stores (ordered before this point by the barrier above) happen-before the
        // ticket; the last CTA acquires every ticket and then fences ONCE at system scope before raising the flags.
        // (Causality is transitive across scopes in the PTX memory model; a system-scope fence in every CTA costs
        // 2-6 us each under store load and keeps CTAs resident - measured 2x on the whole kernel.)
        __threadfence();
        const unsigned int prev = atomicAdd(P.counter, 1u);
        const int slot = (int)(gridDim.x - 1 - prev);      // 0 for the last CTA to retire
        if (slot == 0) {
            if (P.dbg) P.dbg[2] = global_timer_ns();
            atomicExch(P.counter, 0u);
            __threadfence_system();  // one fence orders everything the grid stored before ALL flag stores below
            for (int p = 0; p < P.nranks; ++p) st_relaxed_sys(P.flag[p] + P.rank, P.epoch);
            if (P.dbg) P.dbg[3] = global_timer_ns();
        }
        s_slot = slot;
    }
    __syncthreads();
    // fused owner-side finish: the interface is a few thousand entries; the last FIN_CTAS CTAs to retire fold it
    // (they spin on this rank's flag array until every rank, this one included, has published the epoch)
    const int nfin = (int)min((unsigned)FIN_CTAS, gridDim.x);
    if (P.fused && s_slot < nfin) {
        finish_body(P.F, s_slot, nfin, s_slot == 0 ? P.dbg : nullptr);
        __syncthreads();
        if (P.dbg && s_slot == 0 && threadIdx.x == 0) P.dbg[5] = global_timer_ns();
    }
}

static __global__ void __launch_bounds__(256) finish_kernel(const FinishParams P) {
    finish_body(P, (int)blockIdx.x, (int)gridDim.x);
}

}  // namespace mgb
