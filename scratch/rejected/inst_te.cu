// thread-per-element instantiations (fem1d: 2-node elements, fem2d: 7-node elements; fine levels, one cone)
#include "inst_common.cuh"
#include "kernels_te.cuh"

namespace mgb {
namespace {
template <int B, int D, int FLAGS>
void launch_te_one(const ElemParams& P, int64_t nblk, cudaStream_t st) {
    auto kern = element_te_kernel<B, D, FLAGS>;
    constexpr int smem = TeShape<B, D>::SMEM;
    static bool opted = false;
    if (!opted && smem > 48 * 1024) {
        inst_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "cudaFuncSetAttribute");
        opted = true;
    }
    kern<<<dim3((unsigned)nblk), dim3(32), smem, st>>>(P);
}
template <int B, int D>
void launch_te_flags(const ElemParams& P, int flags, int64_t nblk, cudaStream_t st) {
    switch (canonical_flags(flags)) {
        case 1: launch_te_one<B, D, 1>(P, nblk, st); break;
        case 7: launch_te_one<B, D, 7>(P, nblk, st); break;
        case 8: launch_te_one<B, D, 8>(P, nblk, st); break;
        case 15: launch_te_one<B, D, 15>(P, nblk, st); break;
        default: throw std::runtime_error("assemble: empty flags");
    }
}
template <int B, int D>
int te_ctas() {
    auto kern = element_te_kernel<B, D, 15>;
    constexpr int smem = TeShape<B, D>::SMEM;
    if (smem > 48 * 1024) inst_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "cudaFuncSetAttribute");
    int nb = 0;
    inst_check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 32, smem), "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    return nb;
}
}  // namespace

void launch_element_te(int B, int dim, const ElemParams& P, int flags, int64_t nblk, cudaStream_t st) {
    if (B == 2 && dim == 1) launch_te_flags<2, 1>(P, flags, nblk, st);
    else if (B == 7 && dim == 2) launch_te_flags<7, 2>(P, flags, nblk, st);
    else throw std::runtime_error("thread-per-element kernel not instantiated for this element type");
}
int element_te_ctas_per_sm(int B, int dim) {
    if (B == 2 && dim == 1) return te_ctas<2, 1>();
    if (B == 7 && dim == 2) return te_ctas<7, 2>();
    throw std::runtime_error("thread-per-element kernel not instantiated for this element type");
}
}  // namespace mgb
