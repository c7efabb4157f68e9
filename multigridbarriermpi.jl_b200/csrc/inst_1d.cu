// fem1d instantiations: 2-node broken elements, one derivative operator.
#include "inst_common.cuh"

namespace mgb {
void launch_element_1d(int mode, bool fine, const ElemParams& P, int flags, int64_t nblk, cudaStream_t st) {
    launch_elem_bd<2, 1>(P, mode, fine, flags, nblk, st);
}
}  // namespace mgb
