// fem2d instantiations: 7-node P2+bubble broken triangles, two derivative operators.
#include "inst_common.cuh"

namespace mgb {
void launch_element_2d(int mode, bool fine, const ElemParams& P, int flags, int64_t nblk, size_t smem, cudaStream_t st) {
    launch_elem_bd<7, 2>(P, mode, fine, flags, nblk, smem, st);
}
int element_ctas_per_sm_2d(int mode, bool fine, size_t smem) { return elem_ctas_bd<7, 2>(mode, fine, smem); }
}  // namespace mgb
