// Shared dispatch templates for the per-dimension instantiation units.
#pragma once
#include <stdexcept>
#include <string>

#include "launch.h"

namespace mgb {

inline void inst_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

template <int B, int D, int MODE, bool FINE, int FLAGS>
void launch_elem_one(const ElemParams& P, int64_t nblk, size_t smem, cudaStream_t st) {
    auto kern = element_kernel<B, D, MODE, FINE, FLAGS>;
    static size_t opted = 0;   // per instance: largest dynamic shared memory size opted in so far
    if (smem > 48 * 1024 && smem > opted) {
        inst_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute");
        opted = smem;
    }
    kern<<<dim3((unsigned)nblk), dim3(MGB_ELEM_THREADS), smem, st>>>(P);
}

template <int B, int D, int MODE, bool FINE>
void launch_elem_flags(const ElemParams& P, int flags, int64_t nblk, size_t smem, cudaStream_t st) {
    switch (canonical_flags(flags)) {
        case 1: launch_elem_one<B, D, MODE, FINE, 1>(P, nblk, smem, st); break;
        case 7: launch_elem_one<B, D, MODE, FINE, 7>(P, nblk, smem, st); break;
        case 8: launch_elem_one<B, D, MODE, FINE, 8>(P, nblk, smem, st); break;
        case 15: launch_elem_one<B, D, MODE, FINE, 15>(P, nblk, smem, st); break;
        default: throw std::runtime_error("assemble: empty flags");
    }
}

template <int B, int D, int MODE, bool FINE>
int elem_ctas_per_sm(size_t smem) {
    auto kern = element_kernel<B, D, MODE, FINE, 15>;   // the largest instance bounds them all
    if (smem > 48 * 1024) inst_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute");
    int nb = 0;
    inst_check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, MGB_ELEM_THREADS, smem), "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    return nb;
}

// mode: 0 one cone, 1 feasibility (slack), 2 two cones (parabolic)
template <int B, int D>
void launch_elem_bd(const ElemParams& P, int mode, bool fine, int flags, int64_t nblk, size_t smem, cudaStream_t st) {
    if (mode == 1) {
        if (fine) launch_elem_flags<B, D, 1, true>(P, flags, nblk, smem, st);
        else launch_elem_flags<B, D, 1, false>(P, flags, nblk, smem, st);
    } else if (mode == 2) {
        if (fine) launch_elem_flags<B, D, 2, true>(P, flags, nblk, smem, st);
        else launch_elem_flags<B, D, 2, false>(P, flags, nblk, smem, st);
    } else {
        if (fine) launch_elem_flags<B, D, 0, true>(P, flags, nblk, smem, st);
        else launch_elem_flags<B, D, 0, false>(P, flags, nblk, smem, st);
    }
}

template <int B, int D>
int elem_ctas_bd(int mode, bool fine, size_t smem) {
    if (mode == 1) return fine ? elem_ctas_per_sm<B, D, 1, true>(smem) : elem_ctas_per_sm<B, D, 1, false>(smem);
    if (mode == 2) return fine ? elem_ctas_per_sm<B, D, 2, true>(smem) : elem_ctas_per_sm<B, D, 2, false>(smem);
    return fine ? elem_ctas_per_sm<B, D, 0, true>(smem) : elem_ctas_per_sm<B, D, 0, false>(smem);
}

}  // namespace mgb
