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

}  // namespace mgb
