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

static __global__ void __launch_bounds__(128) dist_finish_kernel(const DistScal D, double t, double* __restrict__ scal) {
    double sum[3];
    const bool ok = dist_collect(D, sum);
    if (threadIdx.x == 0) write_scalars(sum[0], sum[1], sum[2], t, ok, scal);
}

}  // namespace mgb
