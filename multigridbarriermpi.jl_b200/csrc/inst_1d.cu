// fem1d instantiations: 2-node broken elements, one derivative operator.
#include "inst_common.cuh"

namespace mgb {
void launch_element_1d(int mode, bool fine, const ElemParams& P, int flags, int64_t nblk, cudaStream_t st) {
    launch_elem_bd<2, 1>(P, mode, fine, flags, nblk, st);
}
void launch_patch_1d(bool slack, bool fine, int patch, const ElemParams& P, const PatchParams& Q, int flags, int64_t nblk,
                     size_t smem, cudaStream_t st) {
    (void)patch;
    launch_patch_bd<2, 1, 64>(P, Q, slack, fine, flags, nblk, smem, st);
}
}  // namespace mgb
