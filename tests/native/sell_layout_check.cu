// Host-only check of the replay-list layouts of the CSR path (csrc/kernels_csr.cuh: build_sell with sorting
// windows and chunks).  The layout code is host C++; this program replays the lists exactly as the kernels do
// (slice by slice, lane by lane, partial slots, second-stage sums) and compares with the sequential sum of every
// list.  Built and run by tests/test_host.py on the CPU; needs no GPU.
#include "kernels_csr.cuh"

#include <cmath>
#include <cstdio>
#include <random>

int main() {
    std::mt19937 rng(1);
    int totalbad = 0, cases = 0;
    for (int chunk : {1, 4, 16, 1000000})
        for (int sigma : {1, 256, 16384})
            for (int nent : {0, 1, 31, 32, 33, 1000, 5000}) {
                std::vector<int32_t> ptr(nent + 1, 0), src;
                std::vector<double> coef;
                for (int e = 0; e < nent; ++e) {
                    const int len = rng() % 7 == 0 ? rng() % 150 : rng() % 4;   // mostly short lists, a heavy tail
                    for (int r = 0; r < len; ++r) { coef.push_back((rng() % 1000) / 7.0 - 60.0); src.push_back(rng() % 100); }
                    ptr[e + 1] = (int32_t)coef.size();
                }
                std::vector<double> x(100);
                for (auto& v : x) v = (rng() % 1000) / 3.0 - 100.0;
                mgb::SellHost h;
                mgb::build_sell<int32_t>(nent, ptr.data(), coef.data(), src.data(), sigma, chunk, h);
                const size_t nsl = h.off.size() - 1;
                std::vector<double> out(nent, -1.0), part(std::max(1, h.comb_ptr.back()), -5.0);
                std::vector<int> seen(nent, 0);
                int bad = 0;
                uint32_t maxlen = 0;
                for (size_t sl = 0; sl < nsl; ++sl) {
                    maxlen = std::max(maxlen, h.off[sl + 1] - h.off[sl]);
                    for (int l = 0; l < 32; ++l) {
                        double acc = 0.0;
                        size_t q = (size_t)h.off[sl] * 32 + l;
                        for (uint32_t r = h.off[sl]; r < h.off[sl + 1]; ++r, q += 32)
                            if (h.src[q] >= 0) acc = std::fma(h.coef[q], x[h.src[q]], acc);
                        const int c = h.code[sl * 32 + l];
                        if (c >= 0) { out[c] = acc; seen[c]++; }
                        else if (c <= -2) part[-(c + 2)] = acc;
                    }
                }
                for (size_t q = 0; q < h.comb_out.size(); ++q) {
                    double acc = 0.0;
                    for (int r = h.comb_ptr[q]; r < h.comb_ptr[q + 1]; ++r) acc += part[r];
                    out[h.comb_out[q]] = acc;
                    seen[h.comb_out[q]]++;
                }
                if ((int64_t)maxlen > chunk) bad++;   // no lane replays more than one chunk
                for (int e = 0; e < nent; ++e) {
                    double acc = 0.0, mag = 0.0;
                    for (int r = ptr[e]; r < ptr[e + 1]; ++r) { acc = std::fma(coef[r], x[src[r]], acc); mag += std::fabs(coef[r] * x[src[r]]); }
                    if (seen[e] != 1) bad++;
                    else if ((ptr[e + 1] - ptr[e]) <= chunk ? acc != out[e] : std::fabs(acc - out[e]) > 1e-13 * (mag + 1e-300)) bad++;
                }
                if (bad) std::printf("chunk %d sigma %d nent %d: %d bad\n", chunk, sigma, nent, bad);
                totalbad += bad;
                ++cases;
            }
    std::printf("SELL_LAYOUT cases %d bad %d\n", cases, totalbad);
    return totalbad ? 1 : 0;
}
