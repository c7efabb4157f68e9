// Launch entry points of the templated element kernels.  The instantiations live in
// inst_1d.cu / inst_2d.cu so the translation units compile in parallel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace mgb {

bool element_supported(int B, int dim);
// canonical flag sets the kernels are instantiated for: 1 (objective), 7 (objective+gradient+Hessian),
// 8 (apply_D only), 15 (everything + Dz); other requests run the next superset.
int canonical_flags(int flags);

// grid = nblk CTAs of MGB_ELEM_THREADS threads, smem = dynamic shared memory per CTA (opted in above 48 KB)
void launch_element(int B, int dim, int mode, bool fine, const ElemParams& P, int flags, int64_t nblk, size_t smem, cudaStream_t st);
// resident CTAs per SM of the full (flags = 7) instance with that much dynamic shared memory
int element_ctas_per_sm(int B, int dim, int mode, bool fine, size_t smem);

void launch_element_1d(int mode, bool fine, const ElemParams& P, int flags, int64_t nblk, size_t smem, cudaStream_t st);
void launch_element_2d(int mode, bool fine, const ElemParams& P, int flags, int64_t nblk, size_t smem, cudaStream_t st);
int element_ctas_per_sm_1d(int mode, bool fine, size_t smem);
int element_ctas_per_sm_2d(int mode, bool fine, size_t smem);

// thread-per-element kernel (kernels_te.cuh; fine levels, one cone): grid = nblk CTAs of one warp
void launch_element_te(int B, int dim, const ElemParams& P, int flags, int64_t nblk, cudaStream_t st);
int element_te_ctas_per_sm(int B, int dim);

}  // namespace mgb
