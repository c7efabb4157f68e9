// Host symbolic phase.  See plan_host.h.
#include "plan_host.h"

#include <cstdlib>

#include <algorithm>
#include <cstring>
#include <iterator>
#include <numeric>
#include <stdexcept>

namespace mgb {

HostCSR spgemm(const HostCSR& A, const HostCSR& B) {
    if (A.ncols != B.nrows) throw std::runtime_error("spgemm: inner dimensions differ");
    HostCSR C;
    C.nrows = A.nrows;
    C.ncols = B.ncols;
    C.ptr.assign(A.nrows + 1, 0);
    std::vector<int64_t> mark(B.ncols, -1);
    std::vector<double> acc(B.ncols, 0.0);
    std::vector<int32_t> cols;
    for (int64_t i = 0; i < A.nrows; ++i) {
        cols.clear();
        for (int64_t pa = A.ptr[i]; pa < A.ptr[i + 1]; ++pa) {
            const int32_t k = A.idx[pa];
            const double av = A.val[pa];
            for (int64_t pb = B.ptr[k]; pb < B.ptr[k + 1]; ++pb) {
                const int32_t j = B.idx[pb];
                if (mark[j] != i) {
                    mark[j] = i;
                    acc[j] = 0.0;
                    cols.push_back(j);
                }
                acc[j] += av * B.val[pb];
            }
        }
        std::sort(cols.begin(), cols.end());
        for (int32_t j : cols) {
            C.idx.push_back(j);
            C.val.push_back(acc[j]);
        }
        C.ptr[i + 1] = (int64_t)C.idx.size();
    }
    return C;
}

HostCSR transpose(const HostCSR& A) {
    HostCSR T;
    T.nrows = A.ncols;
    T.ncols = A.nrows;
    T.ptr.assign(A.ncols + 1, 0);
    for (int32_t j : A.idx) T.ptr[j + 1]++;
    for (int64_t j = 0; j < A.ncols; ++j) T.ptr[j + 1] += T.ptr[j];
    T.idx.resize(A.nnz());
    T.val.resize(A.nnz());
    std::vector<int64_t> pos(T.ptr.begin(), T.ptr.end() - 1);
    for (int64_t i = 0; i < A.nrows; ++i)
        for (int64_t p = A.ptr[i]; p < A.ptr[i + 1]; ++p) {
            const int64_t d = pos[A.idx[p]]++;
            T.idx[d] = (int32_t)i;
            T.val[d] = A.val[p];
        }
    return T;
}

int64_t count_gram_pattern(const std::vector<HostCSR>& D) {
    const int64_t n = D[0].nrows, N = D[0].ncols;
    // union pattern U (n x N) of all operators, and its transpose
    HostCSR U;
    U.nrows = n; U.ncols = N; U.ptr.assign(n + 1, 0);
    std::vector<int32_t> cols;
    for (int64_t i = 0; i < n; ++i) {
        cols.clear();
        for (const HostCSR& A : D)
            for (int64_t p = A.ptr[i]; p < A.ptr[i + 1]; ++p) cols.push_back(A.idx[p]);
        std::sort(cols.begin(), cols.end());
        cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
        U.idx.insert(U.idx.end(), cols.begin(), cols.end());
        U.ptr[i + 1] = (int64_t)U.idx.size();
    }
    U.val.assign(U.idx.size(), 1.0);
    HostCSR Ut = transpose(U);
    std::vector<int64_t> mark(N, -1);
    int64_t total = 0;
    for (int64_t p = 0; p < N; ++p)
        for (int64_t r = Ut.ptr[p]; r < Ut.ptr[p + 1]; ++r) {
            const int64_t i = Ut.idx[r];
            for (int64_t c = U.ptr[i]; c < U.ptr[i + 1]; ++c)
                if (mark[U.idx[c]] != p) { mark[U.idx[c]] = p; ++total; }
        }
    return total;
}

static int pow2ceil(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}
static int round_up(int v, int a) { return (v + a - 1) / a * a; }

void SlotLayout::build(int B_, int dim_, bool slack_, bool fine_, bool mma_) {
    B = B_;
    dim = dim_;
    slack = slack_;
    fine = fine_;
    mma = mma_;
    LPE = pow2ceil(B);
    NU = 2 + (slack ? 1 : 0);
    if (mma) { off_uu = 0; off_us = 64; off_ss = 128; NS = 192; return; }
    const int ntri = round_up(B * (B + 1) / 2, LPE);
    const int nfull = fine ? B * LPE : round_up(B * B, LPE);
    const int ndiag = fine ? LPE : ntri;
    int o = 0;
    off_uu = o, o += ntri;
    off_us = o, o += nfull;
    off_ss = o, o += ndiag;
    if (slack) {
        off_ut = o, o += nfull;
        off_st = o, o += fine ? LPE : round_up(B * B, LPE);
        off_tt = o, o += ndiag;
    }
    NS = o;
}

int SlotLayout::tri(int q, int q2) const { return q * B - q * (q - 1) / 2 + (q2 - q); }
int SlotLayout::ntri_pad() const { return round_up(B * (B + 1) / 2, LPE); }
int SlotLayout::nfull_pad() const { return round_up(B * B, LPE); }

namespace {
struct UF {
    std::vector<int64_t> p;
    explicit UF(int64_t n) : p(n) { std::iota(p.begin(), p.end(), 0); }
    int64_t find(int64_t a) {
        while (p[a] != a) {
            p[a] = p[p[a]];
            a = p[a];
        }
        return a;
    }
    void unite(int64_t a, int64_t b) {
        a = find(a), b = find(b);
        if (a != b) p[std::max(a, b)] = std::min(a, b);
    }
};
}  // namespace


namespace {
// See ElementPlan::dense.  Fills the pattern and the replay lists exactly like the element path does, so the
// upload / gather machinery of the element path is reused unchanged; only stage 1 is a different kernel.
void build_dense_plan(const std::vector<HostCSR>& D, const HostCSR& R, int64_t n_global, const BarrierDesc& bar, ElementPlan& P,
                      bool want_hessian, int64_t B, const std::vector<int>& var, int dim, int nu, int64_t out0, int64_t out1) {
    const int ND = (int)D.size(), NB = DENSE_NB;
    const int64_t nloc = D[0].nrows, m = R.ncols;
    (void)n_global;
    if (P.mode != 0 || nu != 2 || dim != 3) { P.why = "element block size not supported by the fused kernels (1..8)"; return; }
    if (out0 != 0 || out1 != m) { P.why = "dense path does not shard"; return; }
    if (nloc % B) { P.why = "rows are not whole elements"; return; }
    const int64_t E = nloc / B;
    std::vector<HostCSR> Ek(ND);
    for (int k = 0; k < ND; ++k) Ek[k] = spgemm(D[k], R);
    // worth it?  products of the list replay ~ sum_i |U_i|^2 / 2 vs ~12 k multiply-adds per point of the dense contraction
    {
        double prod = 0.0;
        for (int64_t i = 0; i < nloc; ++i) {
            double len = 0.0;
            for (int k = 0; k < ND; ++k) len += (double)(Ek[k].ptr[i + 1] - Ek[k].ptr[i]);
            prod += 0.5 * len * len;
        }
        const char* ev = getenv("MGB_DENSE_MIN_PRODUCTS");   // tests force the dense path on small elements with 0
        if (prod < (ev ? atof(ev) : 1536.0) * (double)nloc) { P.why = "short operator rows: the list replay of the CSR path is cheaper than a dense contraction"; return; }
    }
    // dof sets of the fine elements, greedy grouping of consecutive elements
    std::vector<std::vector<int32_t>> eset((size_t)E * 2);
    for (int64_t e = 0; e < E; ++e)
        for (int v = 0; v < 2; ++v) {
            auto& t = eset[(size_t)e * 2 + v];
            for (int k = 0; k < ND; ++k) {
                if (var[k] != v) continue;
                for (int64_t i = e * B; i < (e + 1) * B; ++i)
                    for (int64_t q = Ek[k].ptr[i]; q < Ek[k].ptr[i + 1]; ++q) t.push_back(Ek[k].idx[q]);
            }
            std::sort(t.begin(), t.end());
            t.erase(std::unique(t.begin(), t.end()), t.end());
            if ((int)t.size() > NB) { P.why = "an element touches more than 64 unknowns of one variable"; return; }
        }
    std::vector<int64_t> gstart;   // first element of every group
    std::vector<std::vector<int32_t>> gset;
    {
        std::vector<int32_t> cur[2], uni;
        for (int64_t e = 0; e < E; ++e) {
            bool fits = e > 0;
            std::vector<int32_t> nxt[2];
            for (int v = 0; v < 2 && fits; ++v) {
                uni.clear();
                std::set_union(cur[v].begin(), cur[v].end(), eset[(size_t)e * 2 + v].begin(), eset[(size_t)e * 2 + v].end(), std::back_inserter(uni));
                fits = (int)uni.size() <= NB;
                nxt[v] = uni;
            }
            if (!fits) {
                if (e > 0) { gset.push_back(cur[0]); gset.push_back(cur[1]); }
                gstart.push_back(e);
                cur[0] = eset[(size_t)e * 2]; cur[1] = eset[(size_t)e * 2 + 1];
            } else { cur[0] = nxt[0]; cur[1] = nxt[1]; }
        }
        gset.push_back(cur[0]); gset.push_back(cur[1]);
        gstart.push_back(E);
    }
    const int64_t ng = (int64_t)gstart.size() - 1;
    P.dense = true; P.fine = false; P.agg = 1;
    P.B = NB; P.LPE = NB; P.dim = dim; P.NU = nu; P.ND = ND; P.slack = false;
    P.nloc = nloc; P.m = m; P.out0 = 0; P.m_out = m;
    P.lay = SlotLayout(); P.lay.B = NB; P.lay.LPE = NB; P.lay.NU = 2; P.lay.dim = dim; P.lay.NS = 3 * NB * NB;
    P.lay.off_uu = 0; P.lay.off_us = NB * NB; P.lay.off_ss = 2 * NB * NB;
    P.d_ngroups = ng;
    P.d_gdof.assign((size_t)ng * 2 * NB, -1);
    for (int64_t g = 0; g < ng; ++g)
        for (int v = 0; v < 2; ++v) std::copy(gset[(size_t)g * 2 + v].begin(), gset[(size_t)g * 2 + v].end(), P.d_gdof.begin() + ((size_t)g * 2 + v) * NB);
    // chunks: one CTA each, one CTA per SM at a time.  Every group is cut into equal chunks of whole 16-point tiles, at
    // most DENSE_CHUNK points; the target size is the one whose chunk count fills whole waves of a 148-SM part best
    // (fem3d L=5 level 2: 512-point chunks make 512 CTAs = 3.46 waves of 32 tiles, 464-point chunks 576 CTAs = 3.89
    // waves of 29 tiles), small levels get at least ~2 chunks per SM.  Fewer, larger chunks win ties (fewer records).
    constexpr int64_t SMS = 148;
    auto split = [&](int64_t npts, int64_t target, int64_t& nchg, int64_t& size) {
        nchg = std::max<int64_t>(1, (npts + target - 1) / target);
        size = ((npts + nchg - 1) / nchg + 15) / 16 * 16;
        nchg = (npts + size - 1) / size;
    };
    const int64_t cap = std::min<int64_t>(DENSE_CHUNK, std::max<int64_t>(64, (nloc / (2 * SMS) + 15) / 16 * 16));
    int64_t best_target = cap, best_cost = INT64_MAX;
    for (int64_t target = cap; target >= std::max<int64_t>(64, cap / 2); target -= 16) {
        int64_t total = 0, maxtiles = 0;
        for (int64_t g = 0; g < ng; ++g) {
            int64_t nchg, size;
            split((gstart[g + 1] - gstart[g]) * B, target, nchg, size);
            total += nchg; maxtiles = std::max(maxtiles, size / 16);
        }
        const int64_t cost = (total + SMS - 1) / SMS * (maxtiles + 3);   // + pipeline fill / record write-out per chunk
        if (cost < best_cost) { best_cost = cost; best_target = target; }
    }
    std::vector<int64_t> cgroup;
    for (int64_t g = 0; g < ng; ++g) {
        int64_t nchg, size;
        split((gstart[g + 1] - gstart[g]) * B, best_target, nchg, size);
        for (int64_t p0 = gstart[g] * B; p0 < gstart[g + 1] * B; p0 += size) {
            P.d_chunk.push_back(g); P.d_chunk.push_back(p0); P.d_chunk.push_back(std::min<int64_t>(p0 + size, gstart[g + 1] * B));
            cgroup.push_back(g);
        }
    }
    const int64_t nch = (int64_t)cgroup.size();
    P.d_nchunks = nch; P.E = nch;
    if (nch * (int64_t)P.lay.NS > INT32_MAX) { P.dense = false; P.why = "dense slot buffer exceeds int32 indexing"; return; }
    // dense operator rows over the group's dofs; per-point support masks (two 64-bit words: u dofs | s dofs)
    const int NR = dim + 2;   // rows per point: derivative ops (1..dim), u.id (op 0), s.id (op dim+1)
    P.d_rows.assign((size_t)nloc * NR * DENSE_STRIDE, 0.0);
    std::vector<uint64_t> pmask((size_t)nloc * 2, 0);
    for (int64_t g = 0; g < ng; ++g) {
        const std::vector<int32_t>* gs[2] = {&gset[(size_t)g * 2], &gset[(size_t)g * 2 + 1]};
        for (int64_t i = gstart[g] * B; i < gstart[g + 1] * B; ++i)
            for (int k = 0; k < ND; ++k) {
                const int v = var[k];
                const int r = (k == 0) ? dim : (k <= dim ? k - 1 : dim + 1);
                for (int64_t q = Ek[k].ptr[i]; q < Ek[k].ptr[i + 1]; ++q) {
                    const int la = (int)(std::lower_bound(gs[v]->begin(), gs[v]->end(), Ek[k].idx[q]) - gs[v]->begin());
                    P.d_rows[((size_t)i * NR + r) * DENSE_STRIDE + la] = Ek[k].val[q];
                    pmask[(size_t)i * 2 + v] |= 1ull << la;
                }
            }
    }
    // pattern: union over points of support x support, per group through row masks
    P.h_rowptr.assign(m + 1, 0);
    P.h_cptr.assign(1, 0);
    std::vector<uint64_t> rmask((size_t)ng * 2 * NB * 2, 0);   // [group][row: v*NB + la][word: u | s]
    if (want_hessian) {
        for (int64_t g = 0; g < ng; ++g)
            for (int64_t i = gstart[g] * B; i < gstart[g + 1] * B; ++i)
                for (int v = 0; v < 2; ++v) {
                    uint64_t mk = pmask[(size_t)i * 2 + v];
                    while (mk) {
                        const int la = __builtin_ctzll(mk); mk &= mk - 1;
                        uint64_t* rm = &rmask[(((size_t)g * 2 + v) * NB + la) * 2];
                        rm[0] |= pmask[(size_t)i * 2]; rm[1] |= pmask[(size_t)i * 2 + 1];
                    }
                }
        // (row, col, contribution) triples: one per chunk of every group that holds the pair
        std::vector<int64_t> rowcnt(m + 1, 0);
        std::vector<int64_t> nchg(ng, 0), ch0(ng, 0);
        for (int64_t c = 0; c < nch; ++c) { if (nchg[cgroup[c]]++ == 0) ch0[cgroup[c]] = c; }
        auto for_pairs = [&](auto&& fn) {
            for (int64_t g = 0; g < ng; ++g)
                for (int v1 = 0; v1 < 2; ++v1)
                    for (int la = 0; la < NB; ++la) {
                        const int32_t ga = P.d_gdof[((size_t)g * 2 + v1) * NB + la];
                        if (ga < 0) continue;
                        for (int v2 = 0; v2 < 2; ++v2) {
                            uint64_t mk = rmask[(((size_t)g * 2 + v1) * NB + la) * 2 + v2];
                            while (mk) {
                                const int lb = __builtin_ctzll(mk); mk &= mk - 1;
                                const int32_t gb = P.d_gdof[((size_t)g * 2 + v2) * NB + lb];
                                // slot inside a chunk record: uu[la][lb] | us[u la][s lb] | ss[la][lb]; (s, u) reads us transposed
                                const int slot = (v1 == 0 && v2 == 0) ? la * NB + lb
                                               : (v1 == 0 && v2 == 1) ? NB * NB + la * NB + lb
                                               : (v1 == 1 && v2 == 0) ? NB * NB + lb * NB + la
                                                                      : 2 * NB * NB + la * NB + lb;
                                fn(g, ga, gb, slot);
                            }
                        }
                    }
        };
        for_pairs([&](int64_t g, int32_t ga, int32_t, int) { rowcnt[ga + 1] += nchg[g]; });
        for (int64_t a = 0; a < m; ++a) rowcnt[a + 1] += rowcnt[a];
        std::vector<int32_t> tb(rowcnt[m]), ts(rowcnt[m]);
        std::vector<int64_t> fillpos(rowcnt.begin(), rowcnt.end() - 1);
        for_pairs([&](int64_t g, int32_t ga, int32_t gb, int slot) {
            for (int64_t c = ch0[g]; c < ch0[g] + nchg[g]; ++c) {
                const int64_t d = fillpos[ga]++;
                tb[d] = gb; ts[d] = (int32_t)(c * P.lay.NS + slot);
            }
        });
        P.h_cidx.resize(tb.size());
        P.h_cptr.clear();
        std::vector<std::pair<int32_t, int32_t>> rowbuf;
        int64_t outc = 0;
        for (int64_t a = 0; a < m; ++a) {
            rowbuf.clear();
            for (int64_t d = rowcnt[a]; d < rowcnt[a + 1]; ++d) rowbuf.emplace_back(tb[d], ts[d]);
            std::sort(rowbuf.begin(), rowbuf.end());
            int32_t last = -1;
            for (auto& pr : rowbuf) {
                if (pr.first != last) { P.h_colidx.push_back(pr.first); P.h_cptr.push_back(outc); last = pr.first; }
                P.h_cidx[outc++] = pr.second;
            }
            if ((int64_t)P.h_colidx.size() > INT32_MAX) throw std::runtime_error("nnz(H) exceeds int32 indexing");
            P.h_rowptr[a + 1] = (int32_t)P.h_colidx.size();
        }
        P.h_cptr.push_back(outc);
    }
    // gradient lists: rel record of a chunk = [u dofs (NB) | s dofs (NB)]
    std::vector<int64_t> gcnt(m + 1, 0);
    for (int64_t c = 0; c < nch; ++c)
        for (int k = 0; k < 2 * NB; ++k) { const int32_t a = P.d_gdof[(size_t)cgroup[c] * 2 * NB + k]; if (a >= 0) gcnt[a + 1]++; }
    for (int64_t a = 0; a < m; ++a) gcnt[a + 1] += gcnt[a];
    P.g_cptr = gcnt;
    P.g_cidx.resize(gcnt[m]);
    std::vector<int64_t> gpos(gcnt.begin(), gcnt.end() - 1);
    for (int64_t c = 0; c < nch; ++c)
        for (int k = 0; k < 2 * NB; ++k) { const int32_t a = P.d_gdof[(size_t)cgroup[c] * 2 * NB + k]; if (a >= 0) P.g_cidx[gpos[a]++] = (int32_t)(c * 2 * NB + k); }
    P.ok = true;
}
}  // namespace

void build_element_plan(const std::vector<HostCSR>& D, const HostCSR& R, int64_t n_global, const double* w_local,
                        const BarrierDesc& bar, ElementPlan& P, bool want_hessian, int64_t out0, int64_t out1, bool allow_agg) {
    P.ok = false;
    const int ND = (int)D.size();
    const int64_t nloc = D[0].nrows, N = D[0].ncols, m = R.ncols;
    if (out1 < 0) out1 = m;
    if (out0 < 0 || out0 > out1 || out1 > m) { P.why = "output row range outside 0..m"; return; }
    const int64_t mo = out1 - out0;
    P.out0 = out0; P.m_out = mo;
    if (N % n_global) { P.why = "N is not a multiple of n"; return; }
    const int nu = (int)(N / n_global);
    const bool two = bar.nidx2 > 0;
    const bool slack = bar.slack != 0 || two;   // three state variables
    const int dim = ND - 2 - (slack ? 1 : 0);
    if (bar.kind != 1 || dim < 1 || dim > 3 || nu != 2 + (slack ? 1 : 0)) { P.why = "not the p-Laplace operator table"; return; }
    if (bar.nidx != dim + 1) { P.why = "barrier idx does not select (derivatives, s)"; return; }
    for (int j = 0; j < dim; ++j)
        if (bar.idx[j] != 1 + j) { P.why = "barrier idx does not select (derivatives, s)"; return; }
    if (bar.idx[dim] != (two ? dim + 2 : dim + 1)) { P.why = "barrier idx does not select (derivatives, s)"; return; }
    if (two && (bar.slack != 0 || bar.nidx2 != 2 || bar.idx2[0] != 0 || bar.idx2[1] != dim + 1)) {
        P.why = "second cone is not (u.id, s1.id): general CSR kernels";
        return;
    }
    P.mode = two ? 2 : (bar.slack != 0 ? 1 : 0);
    // operator -> state variable
    std::vector<int> var(ND, 0);
    for (int k = 0; k < ND; ++k) {
        int64_t lo = N, hi = -1;
        for (int32_t j : D[k].idx) lo = std::min<int64_t>(lo, j), hi = std::max<int64_t>(hi, j);
        if (hi >= 0) {
            if (lo / n_global != hi / n_global) { P.why = "operator spans several state variables"; return; }
            var[k] = (int)(lo / n_global);
        } else {
            var[k] = (k <= dim) ? 0 : (k - dim);
        }
        const int expect = (k <= dim) ? 0 : (k - dim);
        if (var[k] != expect) { P.why = "operator table is not [u.id; u.d*; s.id; (slack.id)]"; return; }
    }
    // elements = connected components of the row/column graph of the fine operators
    UF uf(nloc);
    {
        std::vector<int64_t> first(N, -1);
        for (int k = 0; k < ND; ++k)
            for (int64_t i = 0; i < nloc; ++i)
                for (int64_t p = D[k].ptr[i]; p < D[k].ptr[i + 1]; ++p) {
                    int64_t& f = first[D[k].idx[p]];
                    if (f < 0) f = i; else uf.unite(i, f);
                }
    }
    int64_t B = 0;
    {
        int64_t run = 0, root = -1;
        for (int64_t i = 0; i <= nloc; ++i) {
            const int64_t r = (i < nloc) ? uf.find(i) : -2;
            if (r != root) {
                if (root >= 0) {
                    if (B == 0) B = run;
                    if (run != B) { P.why = "elements are not uniform consecutive row blocks"; return; }
                }
                if (i < nloc && r != i) { P.why = "element rows are not consecutive"; return; }
                root = r;
                run = 0;
            }
            ++run;
        }
    }
    if (B > 8) {   // large elements (fem3d Q3): the dense path where it pays, else the general CSR kernels
        build_dense_plan(D, R, n_global, bar, P, want_hessian, B, var, dim, nu, out0, out1);
        return;
    }
    if (B < 1) { P.why = "element block size not supported by the fused kernels (1..8)"; return; }
    const int64_t E = nloc / B;
    P.B = (int)B; P.dim = dim; P.NU = nu; P.ND = ND; P.slack = slack;
    P.E = E; P.nloc = nloc; P.m = m;
    const int LPE = pow2ceil((int)B);
    P.LPE = LPE;

    std::vector<HostCSR> Ek(ND);
    for (int k = 0; k < ND; ++k) Ek[k] = spgemm(D[k], R);

    // local column lists
    P.lcols.assign((size_t)nu * E * LPE, -1);
    std::vector<int32_t> tmp;
    for (int v = 0; v < nu; ++v)
        for (int64_t e = 0; e < E; ++e) {
            tmp.clear();
            for (int k = 0; k < ND; ++k) {
                if (var[k] != v) continue;
                for (int64_t i = e * B; i < (e + 1) * B; ++i)
                    for (int64_t p = Ek[k].ptr[i]; p < Ek[k].ptr[i + 1]; ++p) tmp.push_back(Ek[k].idx[p]);
            }
            std::sort(tmp.begin(), tmp.end());
            tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
            if ((int64_t)tmp.size() > B) { P.why = "element touches more dofs per variable than it has nodes"; return; }
            for (size_t q = 0; q < tmp.size(); ++q) P.lcols[((size_t)e * nu + v) * LPE + q] = tmp[q];
        }
    auto local_of = [&](int v, int64_t e, int32_t col) -> int {
        const int32_t* lc = &P.lcols[((size_t)e * nu + v) * LPE];
        for (int q = 0; q < (int)B; ++q)
            if (lc[q] == col) return q;
        return -1;
    };
    // id-like operators: fine (one owned column per row) or dense
    const int idop[3] = {0, dim + 1, dim + 2};
    bool fine = true;
    for (int v = 0; v < nu && fine; ++v) {
        const HostCSR& A = Ek[idop[v]];
        for (int64_t e = 0; e < E && fine; ++e) {
            unsigned used = 0;
            for (int64_t i = e * B; i < (e + 1) * B; ++i) {
                const int64_t cnt = A.ptr[i + 1] - A.ptr[i];
                if (cnt > 1) { fine = false; break; }
                if (cnt == 1) {
                    const int q = local_of(v, e, A.idx[A.ptr[i]]);
                    if (used & (1u << q)) { fine = false; break; }
                    used |= 1u << q;
                }
            }
        }
    }
    P.fine = fine;
    // coarse levels: the children of one coarse element are consecutive elements with identical dofs; the largest
    // power of two (up to the elements of one warp) for which every aligned group agrees is summed in the kernel
    P.agg = 1;
    if (!fine && allow_agg) {
        const int epw = 32 / LPE;
        for (int g = 2; g <= epw; g *= 2) {
            bool same = true;
            for (int64_t e = 0; e < E && same; ++e) {
                const int64_t e0 = e / g * g;
                if (e == e0) continue;
                same = std::equal(P.lcols.begin() + (size_t)e * nu * LPE, P.lcols.begin() + (size_t)(e + 1) * nu * LPE,
                                  P.lcols.begin() + (size_t)e0 * nu * LPE);
            }
            if (!same) break;
            P.agg = g;
        }
    }
    // coarse fem2d levels with whole warps of siblings: 8 x 8 blocks on the tensor cores (kernels.cuh, ElemParams::mma)
    P.mma = !fine && B == 7 && dim == 2 && P.mode == 0 && P.agg == 4 && getenv("MGB_NO_MMA2D") == nullptr;
    P.lay.build((int)B, dim, slack, fine, P.mma);
    const SlotLayout& lay = P.lay;

    // per-point records in element-local columns (one contiguous, 16-byte aligned record per point)
    {
        const int rwf = dim * (int)B + 1 + nu + 1, rwc = dim * (int)B + 1 + nu * (int)B;
        const int RW = ((fine ? rwf : rwc) + 1) / 2 * 2;
        P.RW = RW;
        P.prec.assign((size_t)nloc * RW, 0.0);
        for (int kd = 0; kd < dim; ++kd) {
            const HostCSR& A = Ek[1 + kd];
            for (int64_t i = 0; i < nloc; ++i)
                for (int64_t p = A.ptr[i]; p < A.ptr[i + 1]; ++p)
                    P.prec[(size_t)i * RW + kd * B + local_of(0, i / B, A.idx[p])] = A.val[p];
        }
        for (int64_t i = 0; i < nloc; ++i) P.prec[(size_t)i * RW + dim * B] = w_local[i];
        if (fine) {
            for (int64_t i = 0; i < nloc; ++i) {
                unsigned long long bits = 0;
                for (int v = 0; v < nu; ++v) {
                    const HostCSR& A = Ek[idop[v]];
                    unsigned lq = 255;
                    if (A.ptr[i + 1] > A.ptr[i]) {
                        P.prec[(size_t)i * RW + dim * B + 1 + v] = A.val[A.ptr[i]];
                        lq = (unsigned)local_of(v, i / B, A.idx[A.ptr[i]]);
                    }
                    bits |= (unsigned long long)lq << (8 * v);
                }
                double packed;
                std::memcpy(&packed, &bits, sizeof(double));
                P.prec[(size_t)i * RW + dim * B + 1 + nu] = packed;
            }
        } else {
            for (int v = 0; v < nu; ++v) {
                const HostCSR& A = Ek[idop[v]];
                for (int64_t i = 0; i < nloc; ++i)
                    for (int64_t p = A.ptr[i]; p < A.ptr[i + 1]; ++p)
                        P.prec[(size_t)i * RW + dim * B + 1 + v * B + local_of(v, i / B, A.idx[p])] = A.val[p];
            }
        }
    }

    // structural pattern of R'(sum_jk D_j' diag D_k)R = union_i U_i x U_i, U_i = support of point i
    // over all operators (products keep structural zeros; see DESIGN.md "explicit-zero policy").
    const int NL = nu * (int)B;  // local index alpha = v*B + q
    auto slot_of = [&](int a1, int a2, const std::vector<int>& own /*[nu][B] local col per point*/) -> int {
        int v1 = a1 / (int)B, q1 = a1 % (int)B, v2 = a2 / (int)B, q2 = a2 % (int)B;
        if (v1 > v2 || (v1 == v2 && q1 > q2)) { std::swap(v1, v2); std::swap(q1, q2); }
        const int NT = lay.ntri_pad(), NF = lay.nfull_pad();
        if (lay.mma) return ((v1 == 0 && v2 == 0) ? lay.off_uu : (v1 == 0 ? lay.off_us : lay.off_ss)) + q1 * 8 + q2;
        if (v1 == 0 && v2 == 0) return lay.packed(lay.off_uu, lay.tri(q1, q2), NT);
        if (v1 == 0 && v2 == 1) return lay.fine ? lay.off_us + q1 * lay.LPE + q2 : lay.packed(lay.off_us, q1 * lay.B + q2, NF);
        if (v1 == 0 && v2 == 2) return lay.fine ? lay.off_ut + q1 * lay.LPE + q2 : lay.packed(lay.off_ut, q1 * lay.B + q2, NF);
        if (v1 == 1 && v2 == 1) return lay.fine ? (q1 == q2 ? lay.off_ss + q1 : -1) : lay.packed(lay.off_ss, lay.tri(q1, q2), NT);
        if (v1 == 2 && v2 == 2) return lay.fine ? (q1 == q2 ? lay.off_tt + q1 : -1) : lay.packed(lay.off_tt, lay.tri(q1, q2), NT);
        if (v1 == 1 && v2 == 2) {
            if (!lay.fine) return lay.packed(lay.off_st, q1 * lay.B + q2, NF);
            for (int l = 0; l < (int)B; ++l)
                if (own[1 * B + l] == q1 && own[2 * B + l] == q2) return lay.off_st + l;
            return -1;
        }
        return -1;
    };

    P.h_rowptr.assign(mo + 1, 0);
    P.h_cptr.assign(1, 0);
    if (want_hessian) {
    std::vector<int64_t> rowcnt(mo + 1, 0);
    std::vector<uint8_t> pres((size_t)NL * NL);
    std::vector<uint8_t> U((size_t)B * NL);
    std::vector<int> own((size_t)nu * B, -1);
    // two passes: count, fill
    std::vector<int32_t> tb, ts;
    std::vector<int64_t> fillpos;
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            for (int64_t a = 0; a < mo; ++a) rowcnt[a + 1] += rowcnt[a];
            tb.resize(rowcnt[mo]);
            ts.resize(rowcnt[mo]);
            fillpos.assign(rowcnt.begin(), rowcnt.end() - 1);
        }
        for (int64_t e = 0; e < E; e += P.agg) {   // one record per aggregation group (agg = 1: per element)
            std::fill(pres.begin(), pres.end(), 0);
            for (int64_t ee = e; ee < std::min<int64_t>(e + P.agg, E); ++ee) {
                std::fill(U.begin(), U.end(), 0);
                if (ee == e) std::fill(own.begin(), own.end(), -1);
                for (int k = 0; k < ND; ++k) {
                    const HostCSR& A = Ek[k];
                    const int v = var[k];
                    for (int l = 0; l < (int)B; ++l) {
                        const int64_t i = ee * B + l;
                        for (int64_t p = A.ptr[i]; p < A.ptr[i + 1]; ++p) {
                            const int q = local_of(v, ee, A.idx[p]);
                            U[(size_t)l * NL + v * B + q] = 1;
                            if (k == idop[v] && ee == e) own[v * B + l] = q;
                        }
                    }
                }
                for (int l = 0; l < (int)B; ++l)
                    for (int a1 = 0; a1 < NL; ++a1)
                        if (U[(size_t)l * NL + a1])
                            for (int a2 = 0; a2 < NL; ++a2)
                                if (U[(size_t)l * NL + a2]) pres[(size_t)a1 * NL + a2] = 1;
            }
            for (int a1 = 0; a1 < NL; ++a1) {
                const int32_t ga = P.lcols[((size_t)e * nu + a1 / B) * LPE + a1 % B];
                if (ga < out0 || ga >= out1) continue;   // eliminated dof (-1) or a row another rank owns
                for (int a2 = 0; a2 < NL; ++a2) {
                    if (!pres[(size_t)a1 * NL + a2]) continue;
                    const int32_t gb = P.lcols[((size_t)e * nu + a2 / B) * LPE + a2 % B];
                    if (gb < 0) continue;
                    if (pass == 0) { rowcnt[ga - out0 + 1]++; continue; }
                    const int sl = slot_of(a1, a2, own);
                    if (sl < 0) throw std::runtime_error("internal: structurally present pair without a slot");
                    const int64_t abs_slot = (e / P.agg) * lay.NS + sl;
                    if (abs_slot > INT32_MAX) throw std::runtime_error("element slot buffer exceeds int32 indexing");
                    const int64_t d = fillpos[ga - out0]++;
                    tb[d] = gb;
                    ts[d] = (int32_t)abs_slot;
                }
            }
        }
    }
    if (((int64_t)E + P.agg - 1) / P.agg * lay.NS > INT32_MAX) throw std::runtime_error("element slot buffer exceeds int32 indexing");
    // sort each row by (column, source) and compress
    P.h_rowptr.assign(mo + 1, 0);
    P.h_colidx.clear();
    P.h_cptr.clear();
    P.h_cidx.resize(tb.size());
    std::vector<std::pair<int32_t, int32_t>> rowbuf;
    int64_t outc = 0;
    for (int64_t a = 0; a < mo; ++a) {
        rowbuf.clear();
        for (int64_t d = rowcnt[a]; d < rowcnt[a + 1]; ++d) rowbuf.emplace_back(tb[d], ts[d]);
        std::sort(rowbuf.begin(), rowbuf.end());
        int32_t last = -1;
        for (auto& pr : rowbuf) {
            if (pr.first != last) {
                P.h_colidx.push_back(pr.first);
                P.h_cptr.push_back(outc);
                last = pr.first;
            }
            P.h_cidx[outc++] = pr.second;
        }
        if ((int64_t)P.h_colidx.size() > INT32_MAX) throw std::runtime_error("nnz(H) exceeds int32 indexing");
        P.h_rowptr[a + 1] = (int32_t)P.h_colidx.size();
    }
    P.h_cptr.push_back(outc);
    }  // want_hessian

    // gradient replay lists
    std::vector<int64_t> gcnt(mo + 1, 0);
    for (int v = 0; v < nu; ++v)
        for (int64_t e = 0; e < E; e += P.agg)
            for (int q = 0; q < (int)B; ++q) {
                const int32_t a = P.lcols[((size_t)e * nu + v) * LPE + q];
                if (a >= out0 && a < out1) gcnt[a - out0 + 1]++;
            }
    for (int64_t a = 0; a < mo; ++a) gcnt[a + 1] += gcnt[a];
    P.g_cptr = gcnt;
    P.g_cidx.resize(gcnt[mo]);
    std::vector<int64_t> gpos(gcnt.begin(), gcnt.end() - 1);
    for (int64_t e = 0; e < E; e += P.agg)  // element-major so every list is ordered by element
        for (int v = 0; v < nu; ++v)
            for (int q = 0; q < (int)B; ++q) {
                const int32_t a = P.lcols[((size_t)e * nu + v) * LPE + q];
                if (a >= out0 && a < out1) P.g_cidx[gpos[a - out0]++] = (int32_t)(((e / P.agg) * nu + v) * LPE + q);
            }
    P.ok = true;
}

void dist_select_elements(const std::vector<int32_t>& lcols, int64_t E, int NU, int LPE, int B, int rank, int nranks,
                          const int64_t* out_part, std::vector<int64_t>& rows_sel, int64_t& n_primary_rows) {
    const int64_t out0 = out_part[rank], out1 = out_part[rank + 1];
    auto owner = [&](int32_t a) { return (int)(std::upper_bound(out_part, out_part + nranks + 1, (int64_t)a) - out_part) - 1; };
    std::vector<int64_t> prim, halo;
    for (int64_t e = 0; e < E; ++e) {
        bool touch = false;
        int32_t first_any = -1, first_last = -1;
        for (int v = 0; v < NU; ++v)
            for (int q = 0; q < LPE; ++q) {
                const int32_t a = lcols[((size_t)e * NU + v) * LPE + q];
                if (a < 0) continue;
                touch = touch || (a >= out0 && a < out1);
                if (first_any < 0) first_any = a;
                if (v == NU - 1 && first_last < 0) first_last = a;
            }
        const int32_t key = first_last >= 0 ? first_last : first_any;
        const int prim_rank = key >= 0 ? owner(key) : 0;
        if (prim_rank == rank) prim.push_back(e);          // (an element without any dof still counts in the objective)
        else if (touch) halo.push_back(e);
    }
    rows_sel.clear();
    for (int64_t e : prim) for (int l = 0; l < B; ++l) rows_sel.push_back(e * B + l);
    n_primary_rows = (int64_t)rows_sel.size();
    for (int64_t e : halo) for (int l = 0; l < B; ++l) rows_sel.push_back(e * B + l);
}

void build_csr_plan(const std::vector<HostCSR>& D, const HostCSR& R, const BarrierDesc& bar, CsrPlan& P,
                    bool want_hessian) {
    const int ND = (int)D.size();
    P.ND = ND;
    P.nloc = D[0].nrows;
    P.m = R.ncols;
    const int64_t m = P.m;
    P.E.resize(ND);
    std::vector<HostCSR> Et(ND);
    for (int k = 0; k < ND; ++k) {
        P.E[k] = spgemm(D[k], R);
        Et[k] = transpose(P.E[k]);
    }
    if ((int64_t)ND * P.nloc > INT32_MAX) throw std::runtime_error("csr path: gradient source index exceeds int32");
    // merged transposed lists (gradient)
    P.gt_ptr.assign(m + 1, 0);
    for (int64_t a = 0; a < m; ++a) {
        for (int k = 0; k < ND; ++k)
            for (int64_t p = Et[k].ptr[a]; p < Et[k].ptr[a + 1]; ++p) {
                P.gt_coef.push_back(Et[k].val[p]);
                P.gt_src.push_back((int32_t)((int64_t)k * P.nloc + Et[k].idx[p]));
            }
        if ((int64_t)P.gt_coef.size() > INT32_MAX) throw std::runtime_error("csr path: gradient list exceeds int32");
        P.gt_ptr[a + 1] = (int32_t)P.gt_coef.size();
    }
    // V columns: unique operator pairs coupled by at least one cone
    bool act[8][8] = {{false}};
    auto mark_cone = [&](const int* idx, int cnt, bool with_slack) {
        int cols[10], nc = 0;
        for (int j = 0; j < cnt; ++j) cols[nc++] = idx[j];
        if (with_slack) cols[nc++] = ND - 1;
        for (int x = 0; x < nc; ++x)
            for (int y = 0; y < nc; ++y) act[cols[x]][cols[y]] = true;
    };
    mark_cone(bar.idx, bar.nidx, bar.slack != 0);
    if (bar.nidx2 > 0) mark_cone(bar.idx2, bar.nidx2, false);
    P.npair = 0;
    for (int a = 0; a < 8; ++a)
        for (int b = 0; b < 8; ++b) P.pair_col[a][b] = -1;
    for (int a = 0; a < ND; ++a)
        for (int b = a; b < ND; ++b)
            if (act[a][b]) {
                P.pair_a[P.npair] = a; P.pair_b[P.npair] = b;
                P.pair_col[a][b] = P.pair_col[b][a] = P.npair++;
            }
    if ((int64_t)std::max(P.npair, 1) * P.nloc > INT32_MAX) throw std::runtime_error("csr path: V index exceeds int32");
    // pattern: row a = union over (ka, i in column a of E_ka) of union_kb supp(E_kb[i,:])  (structural: every
    // operator pair, also the ones whose F2 block is identically zero - Julia products keep structural zeros)
    P.h_rowptr.assign(m + 1, 0);
    P.prod_ptr.assign(1, 0);
    std::vector<int32_t> mark(m, -1), pos(m, 0), cols;
    std::vector<std::vector<std::pair<double, int32_t>>> rowprod;
    std::vector<int32_t> ent_row;   // row of every pattern entry (for the mirror lookup)
    for (int64_t a = 0; a < m && want_hessian; ++a) {
        cols.clear();
        for (int ka = 0; ka < ND; ++ka)
            for (int64_t p = Et[ka].ptr[a]; p < Et[ka].ptr[a + 1]; ++p) {
                const int32_t i = Et[ka].idx[p];
                for (int kb = 0; kb < ND; ++kb)
                    for (int64_t r = P.E[kb].ptr[i]; r < P.E[kb].ptr[i + 1]; ++r) {
                        const int32_t b = P.E[kb].idx[r];
                        if (mark[b] != (int32_t)a) { mark[b] = (int32_t)a; cols.push_back(b); }
                    }
            }
        std::sort(cols.begin(), cols.end());
        const int64_t row0 = (int64_t)P.h_colidx.size();
        for (size_t t = 0; t < cols.size(); ++t) { pos[cols[t]] = (int32_t)t; P.h_colidx.push_back(cols[t]); }
        if ((int64_t)P.h_colidx.size() > INT32_MAX) throw std::runtime_error("nnz(H) exceeds int32 indexing");
        P.h_rowptr[a + 1] = (int32_t)P.h_colidx.size();
        P.max_row = std::max<int32_t>(P.max_row, (int32_t)cols.size());
        if (rowprod.size() < cols.size()) rowprod.resize(cols.size());
        for (size_t t = 0; t < cols.size(); ++t) rowprod[t].clear();
        for (int ka = 0; ka < ND; ++ka)
            for (int64_t p = Et[ka].ptr[a]; p < Et[ka].ptr[a + 1]; ++p) {
                const int32_t i = Et[ka].idx[p];
                const double alpha = Et[ka].val[p];
                for (int kb = 0; kb < ND; ++kb) {
                    const int col = P.pair_col[ka][kb];
                    if (col < 0) continue;   // F2 block identically zero for this operator pair
                    for (int64_t r = P.E[kb].ptr[i]; r < P.E[kb].ptr[i + 1]; ++r) {
                        const int32_t b = P.E[kb].idx[r];
                        if (b < (int32_t)a) continue;   // lower triangle: filled from the mirror entry
                        rowprod[pos[b]].emplace_back(alpha * P.E[kb].val[r], (int32_t)((int64_t)col * P.nloc + i));
                    }
                }
            }
        for (size_t t = 0; t < cols.size(); ++t) {
            if (cols[t] < (int32_t)a) continue;
            P.up_t.push_back((int32_t)(row0 + (int64_t)t));
            P.up_m.push_back(cols[t] == (int32_t)a ? -1 : -2 - cols[t]);   // resolved below: -2-b = "row b, column a"
            ent_row.push_back((int32_t)a);
            for (auto& pr : rowprod[t]) { P.prod_coef.push_back(pr.first); P.prod_v.push_back(pr.second); }
            if ((int64_t)P.prod_coef.size() > INT32_MAX) throw std::runtime_error("csr path: product list exceeds int32");
            P.prod_ptr.push_back((int32_t)P.prod_coef.size());
        }
    }
    // mirror positions: entry (b, a) of the (structurally symmetric) pattern
    for (size_t j = 0; j < P.up_t.size(); ++j) {
        if (P.up_m[j] == -1) continue;
        const int32_t b = -2 - P.up_m[j], a = ent_row[j];
        const int32_t* lo = P.h_colidx.data() + P.h_rowptr[b];
        const int32_t* hi = P.h_colidx.data() + P.h_rowptr[b + 1];
        const int32_t* it = std::lower_bound(lo, hi, a);
        if (it == hi || *it != a) throw std::runtime_error("internal: pattern of sum_jk E_j' diag E_k is not symmetric");
        P.up_m[j] = (int32_t)(it - P.h_colidx.data());
    }
}

}  // namespace mgb
