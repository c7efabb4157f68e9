#include "launch.h"

#include <stdexcept>

namespace mgb {

bool element_supported(int B, int dim) { return (B == 2 && dim == 1) || (B == 7 && dim == 2); }

int canonical_flags(int flags) {
    const int f = flags & 15;
    if (f == 0) return 0;
    if (f == 1 || f == 8) return f;
    if (f == 9) return 15;
    return (f & 8) ? 15 : 7;
}

void launch_element(int B, int dim, int mode, bool fine, const ElemParams& P, int flags, int64_t nblk, cudaStream_t st) {
    if (B == 2 && dim == 1) launch_element_1d(mode, fine, P, flags, nblk, st);
    else if (B == 7 && dim == 2) launch_element_2d(mode, fine, P, flags, nblk, st);
    else throw std::runtime_error("element kernel not instantiated for this element type");
}

}  // namespace mgb
