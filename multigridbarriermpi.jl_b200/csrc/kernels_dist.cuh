// Multi-GPU numeric phase (one process per GPU): the gather stage fused with the exchange.
//
// Every rank runs element_kernel on its own quadrature rows (whole elements, so apply_D needs no halo),
// then push_kernel replays its frozen contribution lists exactly like gather_kernel does on one GPU -
// but each result is stored straight into the exchange window of the rank that OWNS the output row
// (HPCSparseArrays row partition), through NVLink peer-mapped memory: entries fed by this
// rank alone land in their final position of the owner's CSR value array / gradient block, entries on the
// element-partition interface land in the owner's staging area.  The last CTA to retire publishes an epoch
// flag in every peer's window (release, system scope).  finish_kernel on the owner waits for all peers'
// flags (acquire), sums the staged values in source-rank order (bit-reproducible) and folds the scalars.
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
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct PushParams {
    GatherParams G;                        // local replay lists (hval / grad / scal members unused)
    const int32_t* h_dest;                 // per local Hessian entry: (owner << 27) | offset in the owner's window
    const int32_t* g_dest;                 // per unknown: same, -1 = this rank has no contribution
    double* win[DIST_MAX_RANKS];           // peers' windows (this epoch's parity), win[rank] = own
    unsigned long long* flag[DIST_MAX_RANKS];  // peers' flag arrays; this rank writes flag[p][rank]
    int64_t scal_off[DIST_MAX_RANKS];      // offset of this rank's 4 staged scalars inside window p
    int rank, nranks;
    unsigned long long epoch;
    unsigned int* counter;                 // CTA retirement counter (zero between launches)
};

__device__ __forceinline__ double* dist_dst(const PushParams& P, int32_t d) {
    return P.win[d >> DIST_RANK_SHIFT] + (d & DIST_OFF_MASK);
}

static __global__ void __launch_bounds__(256) push_kernel(const __grid_constant__ PushParams P) {
    const GatherParams& G = P.G;
    const int64_t b = blockIdx.x;
    if (b < G.nblk_h) {
        const int64_t base = b * (256 * GATHER_UNROLL) + threadIdx.x;
        int2 src[GATHER_UNROLL];
        int32_t dst[GATHER_UNROLL];
#pragma unroll
        for (int j = 0; j < GATHER_UNROLL; ++j) {
            const int64_t t = base + (int64_t)j * 256;
            src[j] = (t < G.nnzH) ? __ldg(&G.h_src2[t]) : make_int2(-1, -1);
            dst[j] = (t < G.nnzH) ? __ldg(&P.h_dest[t]) : 0;
        }
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
        if (li < G.n_long) {
            const int64_t c0 = __ldg(&G.h_lptr[li]), c1 = __ldg(&G.h_lptr[li + 1]);
            double acc = 0.0;
            for (int64_t cix = c0; cix < c1; ++cix) acc += G.sel[__ldg(&G.h_lidx[cix])];
            *dist_dst(P, __ldg(&P.h_dest[__ldg(&G.h_lt[li])])) = acc;
        }
    } else if (b < G.nblk_h + G.nblk_l + G.nblk_g) {
        const int64_t a = (b - G.nblk_h - G.nblk_l) * 256 + threadIdx.x;
        if (a < G.m) {
            const int32_t d = __ldg(&P.g_dest[a]);
            if (d >= 0) {
                const int64_t c0 = __ldg(&G.g_cptr[a]), c1 = __ldg(&G.g_cptr[a + 1]);
                double acc = 0.0;
                for (int64_t cix = c0; cix < c1; ++cix) acc += G.rel[__ldg(&G.g_cidx[cix])];
                *dist_dst(P, d) = acc;
            }
        }
    } else {
        // this rank's scalar partials {sum w F, <c,Dz>_w, infeasible count} -> every rank's staging row
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
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int prev = atomicAdd(P.counter, 1u);
        if (prev == gridDim.x - 1) {
            atomicExch(P.counter, 0u);
            __threadfence_system();
            for (int p = 0; p < P.nranks; ++p) st_release_sys(P.flag[p] + P.rank, P.epoch);
        }
    }
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
    int64_t nblk_h, nblk_g;
    int* err;                           // set to 1 when a peer's flag did not arrive in time
};

static __global__ void __launch_bounds__(256) finish_kernel(const FinishParams P) {
    // every CTA waits until all ranks (this one included) have published epoch `epoch`
    if ((int)threadIdx.x < P.nranks) {
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(P.flag + threadIdx.x) < P.epoch) {
            if (global_timer_ns() - t0 > P.timeout_ns) { atomicExch(P.err, 1); break; }
        }
    }
    __syncthreads();
    const int64_t b = blockIdx.x;
    // staged values were written by peers: plain (coherent) loads, never the read-only path
    if (b < P.nblk_h) {
        const int64_t j = b * 256 + threadIdx.x;
        if (j < P.n_fh) {
            const double* __restrict__ stg = P.win + P.off_stg_h;
            double acc = 0.0;
            for (int r = __ldg(&P.fh_ptr[j]); r < __ldg(&P.fh_ptr[j + 1]); ++r) acc += __ldcg(&stg[r]);
            P.win[P.off_h + __ldg(&P.fh_pos[j])] = acc;
        }
        return;
    }
    if (b < P.nblk_h + P.nblk_g) {
        const int64_t j = (b - P.nblk_h) * 256 + threadIdx.x;
        if (j < P.n_fg) {
            const double* __restrict__ stg = P.win + P.off_stg_g;
            double acc = 0.0;
            for (int r = __ldg(&P.fg_ptr[j]); r < __ldg(&P.fg_ptr[j + 1]); ++r) acc += __ldcg(&stg[r]);
            P.win[P.off_g + __ldg(&P.fg_pos[j])] = acc;
        }
        return;
    }
    if (threadIdx.x == 0) {
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

}  // namespace mgb
