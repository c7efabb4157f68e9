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

void launch_element(int B, int dim, int mode, bool fine, const ElemParams& P, int flags, int64_t nblk, cudaStream_t st);

void launch_element_1d(int mode, bool fine, const ElemParams& P, int flags, int64_t nblk, cudaStream_t st);
void launch_element_2d(int mode, bool fine, const ElemParams& P, int flags, int64_t nblk, cudaStream_t st);

}  // namespace mgb
