p='multigridbarriermpi.jl_b200/csrc/kernels.cuh'
s=open(p).read()
s=s.replace("    int off_uu, off_us, off_ss, off_ut, off_st, off_tt, NS;","    int off[3][3];           // slot-record offset of block (row variable, column variable)\n    int NS;")
# generalized group reduce with explicit width
s=s.replace("struct BarrierOut {","""// Same butterfly restricted to the lower W lanes-bits (W = 1 leaves v untouched).
template <int NV, int W>
__device__ __forceinline__ void group_reduce_w(double (&v)[NV], int lane) {
    int len = NV;
#pragma unroll
    for (int M = W / 2; M >= 1; M >>= 1) {
        const int half = len / 2;
        const bool up = (lane & M) != 0;
#pragma unroll
        for (int r = 0; r < NV / 2; ++r) {
            if (r < half) {
                const double lo = v[r], hi = v[r + half];
                const double send = up ? lo : hi;
                const double keep = up ? hi : lo;
                v[r] = keep + shfl_xor_d(send, M);
            }
        }
        len = half;
    }
}

struct BarrierOut {""")
a=s.index("    if (WH) {\n    // ---- Hessian: element-local blocks of sum_jk a_j' (w Y_jk) a_k")
b=s.index("    }  // WH\n}")
new='''    if (WH) {
    // ---- Hessian: element-local blocks of sum_jk a_j' (w Y_jk) a_k.  Every block is stored with full
    // rows (row = local dof of the row variable): lane l ends up with row l, so the gather kernel reads
    // contiguous runs.  rows_of() folds the first butterfly step into the evaluation of X(q,q2): the
    // lower and upper half of the rows are formed on the fly, only HALF*B values are ever live.
    constexpr int HALF = (LPE > 1) ? LPE / 2 : 1;
    constexpr int NVH = HALF * B;
    const bool up = (l & HALF) != 0;
#define MGB_ROWS_OF(VARR, XEXPR)                                                                     \\
    {                                                                                                \\
        _Pragma("unroll") for (int r = 0; r < NVH; ++r) {                                            \\
            const int q = r / B, q2 = r % B;                                                         \\
            double lo, hi = 0.0;                                                                     \\
            { const int qq = q; lo = (XEXPR); }                                                      \\
            if (q + HALF < B) { const int qq = (q + HALF < B) ? q + HALF : 0; hi = (XEXPR); }        \\
            (void)q2;                                                                                \\
            if (LPE > 1) {                                                                           \\
                const double send = up ? lo : hi, keep = up ? hi : lo;                               \\
                VARR[r] = keep + shfl_xor_d(send, HALF);                                             \\
            } else {                                                                                 \\
                VARR[r] = lo;                                                                        \\
            }                                                                                        \\
        }                                                                                            \\
        group_reduce_w<NVH, HALF>(VARR, l);                                                          \\
    }
    // u-u block (derivative operators only: the u.id row of F2 is identically zero)
    {
        double T[D][B];
#pragma unroll
        for (int j = 0; j < D; ++j)
#pragma unroll
            for (int q = 0; q < B; ++q) {
                double tacc = 0.0;
#pragma unroll
                for (int j2 = 0; j2 < D; ++j2) tacc = fma(wi * bo.Hqq[j][j2], a[j2][q], tacc);
                T[j][q] = tacc;
            }
        double v[NVH];
#define MGB_SUU(QQ, Q2) suu_val<D, B>(a, T, QQ, Q2)
        MGB_ROWS_OF(v, MGB_SUU(qq, q2))
#undef MGB_SUU
        if (act_e && l < B) {
#pragma unroll
            for (int q2 = 0; q2 < B; ++q2) sel[P.off[0][0] + l * LPE + q2] = v[q2];
        }
    }
    double bs[B];
#pragma unroll
    for (int q = 0; q < B; ++q) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) acc = fma(a[j][q], wi * bo.Hqs[j], acc);
        bs[q] = acc;
    }
    const double vss = wi * bo.Hss;
    const double vtt = SLACK ? vss + wi * itau * itau : vss;  // slack-slack curvature incl. -log(1+tau)
    if (FINE) {
#pragma unroll
        for (int v2 = 1; v2 < NU; ++v2) {
            if (act && oh[v2]) {
#pragma unroll
                for (int q = 0; q < B; ++q) {
                    const double val = bs[q] * oval[v2];
                    sel[P.off[0][v2] + q * LPE + olq[v2]] = val;   // column of the u x v2 block
                    sel[P.off[v2][0] + olq[v2] * LPE + q] = val;   // row of the v2 x u block
                }
                sel[P.off[v2][v2] + olq[v2]] = (v2 == 2 ? vtt : vss) * oval[v2] * oval[v2];
            }
        }
        if (SLACK && act && oh[1] && oh[SLACK ? 2 : 1]) {
            constexpr int VT = SLACK ? 2 : 1;
            const double val = vss * oval[1] * oval[VT];
            sel[P.off[1][VT] + olq[1]] = val;
            sel[P.off[VT][1] + olq[VT]] = val;
        }
    } else {
#pragma unroll
        for (int v2 = 1; v2 < NU; ++v2) {  // u x {s, slack} and transposes
            double v[NVH];
            MGB_ROWS_OF(v, bs[qq] * aid[FINE ? 0 : v2][FINE ? 0 : q2])
            if (act_e && l < B) {
#pragma unroll
                for (int q2 = 0; q2 < B; ++q2) {
                    sel[P.off[0][v2] + l * LPE + q2] = v[q2];
                    sel[P.off[v2][0] + q2 * LPE + l] = v[q2];
                }
            }
        }
#pragma unroll
        for (int v1 = 1; v1 < NU; ++v1)
#pragma unroll
            for (int v2 = v1; v2 < NU; ++v2) {  // {s,slack} x {s,slack}
                const double c12 = (v1 == 2 && v2 == 2) ? vtt : vss;
                double v[NVH];
                MGB_ROWS_OF(v, c12 * aid[FINE ? 0 : v1][FINE ? 0 : qq] * aid[FINE ? 0 : v2][FINE ? 0 : q2])
                if (act_e && l < B) {
#pragma unroll
                    for (int q2 = 0; q2 < B; ++q2) {
                        sel[P.off[v1][v2] + l * LPE + q2] = v[q2];
                        if (v1 != v2) sel[P.off[v2][v1] + q2 * LPE + l] = v[q2];
                    }
                }
            }
    }
#undef MGB_ROWS_OF
'''
s=s[:a]+new+s[b:]
# helper for S_uu value
s=s.replace("// FLAGS bits: 1 objective, 2 gradient, 4 Hessian, 8 store Dz\n// Per-point work","""template <int D, int B>
__device__ __forceinline__ double suu_val(const double (&a)[D][B], const double (&T)[D][B], int q, int q2) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) acc = fma(a[j][q], T[j][q2], acc);
    return acc;
}

// FLAGS bits: 1 objective, 2 gradient, 4 Hessian, 8 store Dz
// Per-point work""")
s=s.replace("    constexpr int NTRI = (B * (B + 1) / 2 + LPE - 1) / LPE * LPE;\n    constexpr int NFULL = (B * B + LPE - 1) / LPE * LPE;\n\n    const bool act_e","\n    const bool act_e")
open(p,'w').write(s)

p='multigridbarriermpi.jl_b200/csrc/plan_host.h'
h=open(p).read()
a=h.index("struct SlotLayout {")
b=h.index("// One output family (Hessian values or gradient entries)")
h=h[:a]+'''struct SlotLayout {
    int B = 0, LPE = 0, NU = 0, dim = 0;
    bool slack = false, fine = false;
    // block (v1,v2) of the element Hessian, rows = local dofs of variable v1:
    //   full blocks:  off[v1][v2] + q1*LPE + q2
    //   fine, v1,v2 >= 1 (one owned column per point): LPE entries indexed by q1
    int off[3][3] = {{0}};
    int NS = 0;  // doubles per element
    void build(int B_, int dim_, bool slack_, bool fine_);
    bool is_full(int v1, int v2) const { return !fine || v1 == 0 || v2 == 0; }
};

'''+h[b:]
open(p,'w').write(h)
p='multigridbarriermpi.jl_b200/csrc/plan_host.cpp'
c=open(p).read()
a=c.index("void SlotLayout::build(int B_, int dim_, bool slack_, bool fine_) {")
b=c.index("namespace {\nstruct UF {")
c=c[:a]+'''void SlotLayout::build(int B_, int dim_, bool slack_, bool fine_) {
    B = B_;
    dim = dim_;
    slack = slack_;
    fine = fine_;
    LPE = pow2ceil(B);
    NU = 2 + (slack ? 1 : 0);
    int o = 0;
    for (int v1 = 0; v1 < NU; ++v1)
        for (int v2 = 0; v2 < NU; ++v2) {
            off[v1][v2] = o;
            o += is_full(v1, v2) ? B * LPE : LPE;
        }
    NS = round_up(o, 2);
}

'''+c[b:]
a=c.index("    auto slot_of = [&](int a1, int a2, const std::vector<int>& own")
b=c.index("    P.h_rowptr.assign(m + 1, 0);\n    P.h_cptr.assign(1, 0);")
c=c[:a]+'''    auto slot_of = [&](int a1, int a2, const std::vector<int>& own /*[nu][B] local col per point*/) -> int {
        const int v1 = a1 / (int)B, q1 = a1 % (int)B, v2 = a2 / (int)B, q2 = a2 % (int)B;
        if (lay.is_full(v1, v2)) return lay.off[v1][v2] + q1 * lay.LPE + q2;
        if (v1 == v2) return q1 == q2 ? lay.off[v1][v1] + q1 : -1;
        for (int l = 0; l < (int)B; ++l)
            if (own[v1 * B + l] == q1 && own[v2 * B + l] == q2) return lay.off[v1][v2] + q1;
        return -1;
    };

'''+c[b:]
open(p,'w').write(c)
p='multigridbarriermpi.jl_b200/csrc/mgb_b200.cu'
m=open(p).read()
m=m.replace("""    P.off_uu = ep.lay.off_uu; P.off_us = ep.lay.off_us; P.off_ss = ep.lay.off_ss;
    P.off_ut = ep.lay.off_ut; P.off_st = ep.lay.off_st; P.off_tt = ep.lay.off_tt; P.NS = ep.lay.NS;""","""    for (int v1 = 0; v1 < 3; ++v1)
        for (int v2 = 0; v2 < 3; ++v2) P.off[v1][v2] = ep.lay.off[v1][v2];
    P.NS = ep.lay.NS;""")
# default: two-stage unless MGB_PATCH is set
m=m.replace("""            if ((force_flags & MGB_PLAN_TWO_STAGE) == 0) {
                want_patch = (ep.B == 2) ? 64 : 32;
                if (const char* ev = getenv("MGB_PATCH")) want_patch = atoi(ev);
                if (ep.B == 7 && want_patch != 16 && want_patch != 32 && want_patch != 64) want_patch = 32;
                if (ep.B == 2) want_patch = 64;
            }""","""            if (want_patch_early(force_flags)) {
                want_patch = atoi(getenv("MGB_PATCH"));
                if (ep.B == 7 && want_patch != 16 && want_patch != 32 && want_patch != 64) want_patch = 16;
                if (ep.B == 2) want_patch = 64;
            }""")
m=m.replace("bool want_patch_early(int force_flags) { return (force_flags & MGB_PLAN_TWO_STAGE) == 0; }","""// The patch-fused kernel is opt-in (MGB_PATCH=16|32|64): measured slower than the two-stage pair at L=8.
bool want_patch_early(int force_flags) {
    const char* ev = getenv("MGB_PATCH");
    return (force_flags & MGB_PLAN_TWO_STAGE) == 0 && ev && atoi(ev) > 0;
}""")
open(p,'w').write(m)
