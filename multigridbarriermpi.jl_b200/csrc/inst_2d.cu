// fem2d instantiations: 7-node P2+bubble broken triangles, two derivative operators.
#include "inst_common.cuh"

namespace mgb {
void launch_element_2d(int mode, bool fine, const ElemParams& P, int flags, int64_t nblk, cudaStream_t st) {
    launch_elem_bd<7, 2>(P, mode, fine, flags, nblk, st);
}
void launch_patch_2d(bool slack, bool fine, int patch, const ElemParams& P, const PatchParams& Q, int flags, int64_t nblk,
                     size_t smem, cudaStream_t st) {
    if (patch == 16) launch_patch_bd<7, 2, 16>(P, Q, slack, fine, flags, nblk, smem, st);
    else if (patch == 64) launch_patch_bd<7, 2, 64>(P, Q, slack, fine, flags, nblk, smem, st);
    else launch_patch_bd<7, 2, 32>(P, Q, slack, fine, flags, nblk, smem, st);
}
}  // namespace mgb
