// Known-answer check of the host sparse algebra the symbolic phase is built on (csrc/plan_host.cpp: spgemm,
// transpose) against the literals of the reference's own test (test/test_basic_ops.jl:27-66):
//   A = [1 0; 2 3; 0 4] (3x2), B = [1 2 3; 4 5 6] (2x3):  A*B = [1 2 3; 14 19 24; 16 20 24],  A'A = [5 6; 6 25]
// plus the Julia spmatmul convention the pattern logic relies on: structural zeros are kept.
#include <cmath>
#include <cstdio>
#include <vector>

#include "plan_host.h"

static mgb::HostCSR dense(int nr, int nc, const std::vector<double>& a, bool keep_zeros = false) {
    mgb::HostCSR M;
    M.nrows = nr; M.ncols = nc; M.ptr.assign(1, 0);
    for (int i = 0; i < nr; ++i) {
        for (int j = 0; j < nc; ++j)
            if (keep_zeros || a[(size_t)i * nc + j] != 0.0) { M.idx.push_back(j); M.val.push_back(a[(size_t)i * nc + j]); }
        M.ptr.push_back((int64_t)M.idx.size());
    }
    return M;
}
static int expect(const mgb::HostCSR& M, int nr, int nc, const std::vector<double>& want, const char* what) {
    int bad = (M.nrows != nr || M.ncols != nc) ? 1 : 0;
    std::vector<double> got((size_t)nr * nc, 0.0);
    for (int64_t i = 0; i < M.nrows && !bad; ++i)
        for (int64_t p = M.ptr[i]; p < M.ptr[i + 1]; ++p) {
            if (p > M.ptr[i] && M.idx[p] <= M.idx[p - 1]) bad++;   // sorted, no duplicates
            got[(size_t)i * nc + M.idx[p]] += M.val[p];
        }
    for (size_t k = 0; k < want.size(); ++k) bad += std::fabs(got[k] - want[k]) > 1e-14 ? 1 : 0;
    if (bad) std::printf("%s: %d mismatches\n", what, bad);
    return bad;
}
int main() {
    int bad = 0;
    const mgb::HostCSR A = dense(3, 2, {1, 0, 2, 3, 0, 4}), B = dense(2, 3, {1, 2, 3, 4, 5, 6});
    bad += expect(mgb::spgemm(A, B), 3, 3, {1, 2, 3, 14, 19, 24, 16, 20, 24}, "A*B");
    const mgb::HostCSR At = mgb::transpose(A);
    bad += expect(At, 2, 3, {1, 2, 0, 0, 3, 4}, "A'");
    bad += expect(mgb::spgemm(At, A), 2, 2, {5, 6, 6, 25}, "A'A");
    // structural zeros survive a product (Julia spmatmul keeps them; the frozen pattern depends on it)
    const mgb::HostCSR Z = dense(2, 2, {1, 0, 0, 1}, /*keep_zeros=*/true);
    const mgb::HostCSR ZZ = mgb::spgemm(Z, Z);
    bad += (ZZ.nnz() == 4) ? 0 : 1;
    bad += expect(ZZ, 2, 2, {1, 0, 0, 1}, "Z*Z");
    // cancellation does not remove an entry either: [1 -1] * [1; 1] = structural 1x1 zero
    const mgb::HostCSR R1 = dense(1, 2, {1, -1}), C1 = dense(2, 1, {1, 1});
    const mgb::HostCSR P = mgb::spgemm(R1, C1);
    bad += (P.nnz() == 1 && P.val[0] == 0.0) ? 0 : 1;
    std::printf("HOST_ALGEBRA bad %d\n", bad);
    return bad ? 1 : 0;
}
