// fem2d instantiations: 7-node P2+bubble broken triangles, two derivative operators.
#include "inst_common.cuh"

namespace mgb {
void launch_element_2d(int mode, bool fine, const ElemParams& P, int flags, int64_t nblk, cudaStream_t st) {
    launch_elem_bd<7, 2>(P, mode, fine, flags, nblk, st);
}
}  // namespace mgb
