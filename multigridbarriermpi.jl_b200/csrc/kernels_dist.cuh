// Multi-GPU numeric phase (one process per GPU): owner-computes sharding.
//
// A rank's plan covers every broken element that touches one of its output rows (plan_host.h
// dist_select_elements), so element_kernel + gather_kernel complete the owned rows of R'HR and the owned block of
// the gradient locally - no Hessian or gradient value crosses NVLink.  What does cross is the sum of the three
// objective scalars: the gather kernel's scalar block stores this rank's partials straight into every peer's
// window (NVLink peer-mapped memory) as 64-bit words that carry their own epoch tag, then reads the peers' words
// from its own window and sums them in rank order - fused into the kernel that produces them, no fence, no flag,
// no collective, no host round trip (kernels.cuh dist_publish / dist_collect).
// Split mode (mgb_dist_begin / mgb_dist_end) publishes in the gather kernel and collects in dist_finish_kernel, so
// the ranks of a test can share one GPU and one stream without any kernel waiting on another.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace mgb {

// ---- all-gather of the Newton unknown (the halo exchange of apply_D: R*s needs entries of s other ranks own).
// In the reference s is an HPCVector: every rank holds its block only.  Each rank stores its block into EVERY rank's
// copy of the whole vector (NVLink peer stores, coalesced), the last CTA to retire fences once at system scope and
// raises this rank's epoch flag in every window; dist_s_wait_kernel (one CTA) then waits for all flags before the
// element kernel - which reads the local copy - starts.  Two parities (see DistScal).
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct DistGather {
    int rank, nranks;
    unsigned long long epoch;
    double* full[DIST_LL_RANKS];               // every rank's copy of the whole vector (this epoch's parity)
    unsigned long long* flag[DIST_LL_RANKS];   // every rank's flag array; this rank writes flag[p][rank]
    const double* own;                         // this rank's block
    int64_t off, count;                        // its position / length in the whole vector
    unsigned int* counter;                     // CTA retirement counter (zero between launches)
    unsigned long long timeout_ns;
    int* err;
};

static __global__ void __launch_bounds__(256) dist_s_scatter_kernel(const __grid_constant__ DistGather G) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < G.count) {
        const double v = G.own[i];
        for (int p = 0; p < G.nranks; ++p) G.full[p][G.off + i] = v;
    }
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(G.counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        atomicExch(G.counter, 0u);
        __threadfence_system();   // every store of the grid (ordered before the tickets) happens-before the flags
        for (int p = 0; p < G.nranks; ++p) st_release_sys(G.flag[p] + G.rank, G.epoch);
    }
}

static __global__ void __launch_bounds__(32) dist_s_wait_kernel(const __grid_constant__ DistGather G) {
    if ((int)threadIdx.x < G.nranks) {
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(G.flag[G.rank] + threadIdx.x) < G.epoch)
            if (global_timer_ns() - t0 > G.timeout_ns) { atomicExch(G.err, 1); break; }
    }
}

static __global__ void __launch_bounds__(128) dist_finish_kernel(const DistScal D, double t, double* __restrict__ scal) {
    double sum[3];
    const bool ok = dist_collect(D, sum);
    if (threadIdx.x == 0) write_scalars(sum[0], sum[1], sum[2], t, ok, scal);
}

}  // namespace mgb
