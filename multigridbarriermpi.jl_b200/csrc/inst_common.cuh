// Shared dispatch templates for the per-dimension instantiation units.
#pragma once
#include <stdexcept>
#include <string>

#include "launch.h"

namespace mgb {

inline void inst_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

template <int B, int D, int MODE, bool FINE>
void launch_elem_flags(const ElemParams& P, int flags, int64_t nblk, cudaStream_t st) {
    const dim3 g((unsigned)nblk), b(MGB_ELEM_THREADS);
    switch (canonical_flags(flags)) {
        case 1: element_kernel<B, D, MODE, FINE, 1><<<g, b, 0, st>>>(P); break;
        case 7: element_kernel<B, D, MODE, FINE, 7><<<g, b, 0, st>>>(P); break;
        case 8: element_kernel<B, D, MODE, FINE, 8><<<g, b, 0, st>>>(P); break;
        case 15: element_kernel<B, D, MODE, FINE, 15><<<g, b, 0, st>>>(P); break;
        default: throw std::runtime_error("assemble: empty flags");
    }
}

// mode: 0 one cone, 1 feasibility (slack), 2 two cones (parabolic)
template <int B, int D>
void launch_elem_bd(const ElemParams& P, int mode, bool fine, int flags, int64_t nblk, cudaStream_t st) {
    if (mode == 1) {
        if (fine) launch_elem_flags<B, D, 1, true>(P, flags, nblk, st);
        else launch_elem_flags<B, D, 1, false>(P, flags, nblk, st);
    } else if (mode == 2) {
        if (fine) launch_elem_flags<B, D, 2, true>(P, flags, nblk, st);
        else launch_elem_flags<B, D, 2, false>(P, flags, nblk, st);
    } else {
        if (fine) launch_elem_flags<B, D, 0, true>(P, flags, nblk, st);
        else launch_elem_flags<B, D, 0, false>(P, flags, nblk, st);
    }
}

template <int B, int D, bool SLACK, bool FINE, int FLAGS, int PATCH>
void launch_patch_one(const ElemParams& P, const PatchParams& Q, int64_t nblk, size_t smem, cudaStream_t st) {
    auto kern = patch_kernel<B, D, SLACK, FINE, FLAGS, PATCH>;
    if (smem > 40 * 1024) inst_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute");
    kern<<<dim3((unsigned)nblk), dim3(PATCH * Pow2Ceil<B>::value), smem, st>>>(P, Q);
}

template <int B, int D, bool SLACK, bool FINE, int PATCH>
void launch_patch_flags(const ElemParams& P, const PatchParams& Q, int flags, int64_t nblk, size_t smem, cudaStream_t st) {
    switch (canonical_flags(flags)) {
        case 1: launch_patch_one<B, D, SLACK, FINE, 1, PATCH>(P, Q, nblk, smem, st); break;
        case 7: launch_patch_one<B, D, SLACK, FINE, 7, PATCH>(P, Q, nblk, smem, st); break;
        case 8: launch_patch_one<B, D, SLACK, FINE, 8, PATCH>(P, Q, nblk, smem, st); break;
        case 15: launch_patch_one<B, D, SLACK, FINE, 15, PATCH>(P, Q, nblk, smem, st); break;
        default: throw std::runtime_error("assemble: empty flags");
    }
}

template <int B, int D, int PATCH>
void launch_patch_bd(const ElemParams& P, const PatchParams& Q, bool slack, bool fine, int flags, int64_t nblk, size_t smem,
                     cudaStream_t st) {
    if (slack) {
        if (fine) launch_patch_flags<B, D, true, true, PATCH>(P, Q, flags, nblk, smem, st);
        else launch_patch_flags<B, D, true, false, PATCH>(P, Q, flags, nblk, smem, st);
    } else {
        if (fine) launch_patch_flags<B, D, false, true, PATCH>(P, Q, flags, nblk, smem, st);
        else launch_patch_flags<B, D, false, false, PATCH>(P, Q, flags, nblk, smem, st);
    }
}

}  // namespace mgb
