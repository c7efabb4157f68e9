// Numeric phase kernels (sm_100a).  Float64 throughout; no atomics on the value arrays; every
// reduction has a fixed tree so results are bit-reproducible run to run.
//
//   element_kernel : one LPE-lane group per broken element.  Fuses, per quadrature point kept in
//                    registers:  z gather (R), apply_D (reference test/test_apply_d.jl:44), the
//                    barrier map F/F1/F2 (src/MultiGridBarrierMPI.jl:161-170), w.*y scaling
//                    (amgb_diag, src:137-147), the element-local part of D_j' diag D_k and of
//                    R'(.)R (test/test_map_rows_compare.jl:102-123,165-171) and of the gradient.
//                    Cross-point sums use a transposing butterfly over warp shuffles.
//   gather_kernel  : replays the frozen contribution lists into the preallocated CSR value array
//                    and the gradient; finishes the scalar reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef MGB_ELEM_MINBLOCKS
#define MGB_ELEM_MINBLOCKS 5
#endif
#ifndef MGB_ELEM_MINBLOCKS_F0   // objective-only instances (line-search points) need fewer registers: 7 CTAs/SM, 22.5 -> 20.6 us at L=8
#define MGB_ELEM_MINBLOCKS_F0 7
#endif
#ifndef MGB_ELEM_THREADS
#define MGB_ELEM_THREADS 128
#endif

namespace mgb {

struct ElemParams {
    // geometry / plan (device)
    int64_t E, nloc;
    int agg;                 // coarse levels: `agg` (power of two <= elements per warp) consecutive elements share their dofs
                             // (children of one coarse element): their records are summed in the warp and stored once
    int64_t Eprim;           // elements [0, Eprim) count in the objective / <c,Dz> / feasibility scalars (a sharded plan
                             // also evaluates halo elements whose scalars another rank owns); E unless sharded
    const int32_t* lcols;    // [E][NU][LPE] element -> dof (-1: eliminated)
    const double* prec;      // [nloc][RW] per-point record: derivative rows (D*B), w, then
                             //   fine:   own_val[NU], own_lq bytes packed in one 8-byte slot
                             //   coarse: dense id-like rows [NU][B]
                             // RW even -> every record is 16-byte aligned (128-bit loads)
    // per call
    const double* s;         // m
    const double* Dz0;       // nloc x ND or null
    const double* c;         // nloc x ND
    double t, p, p2;         // p2: exponent of the second cone (MODE 2)
    // outputs
    double* sel;             // E*NS
    double* rel;             // E*NU*LPE
    double* part;            // gridDim.x * 4  {f0, cdot, nonfinite count, -}
    double* Dz;              // nloc x ND or null
    int off_uu, off_us, off_ss, off_ut, off_st, off_tt, NS;
    int mma;                 // coarse fem2d levels (B = 7, one cone, four sibling elements per warp): the warp's 8 x 8 blocks
                             // are contracted on the FP64 tensor cores; record = full uu | us | ss blocks of 64 doubles
};

template <int V>
struct Pow2Ceil { static constexpr int value = (V <= 1) ? 1 : 2 * Pow2Ceil<(V + 1) / 2>::value; };

// Programmatic dependent launch (PDL): the element kernel lets its dependent (gather / push) start while its
// last wave drains; the dependent loads its index lists, then waits for the element records.  Both are
// no-ops when the kernel is launched without the programmatic-serialization attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_primary() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ double shfl_d(double v, int src, int width) {
    return __shfl_sync(0xffffffffu, v, src, width);
}
__device__ __forceinline__ double shfl_xor_d(double v, int mask) {
    return __shfl_xor_sync(0xffffffffu, v, mask);
}

// Transposing butterfly: every lane of an LPE-lane group holds NV partial values; afterwards lane l
// holds the group sums of entries [l*NV/LPE, (l+1)*NV/LPE) in v[0 .. NV/LPE).
template <int NV, int LPE>
__device__ __forceinline__ void group_reduce(double (&v)[NV], int lane) {
    static_assert(NV % LPE == 0, "NV must be a multiple of the group width");
    int len = NV;
#pragma unroll
    for (int M = LPE / 2; M >= 1; M >>= 1) {
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

// Coarse levels: the `agg` consecutive elements of an aligned group (lane groups of one warp) share their dofs, so their
// butterfly-reduced block entries are summed across the groups before the store: `agg` times fewer records to write and
// `agg` times shorter contribution lists to replay.  Uniform loop bounds; every lane of the warp takes part.
template <int K, int LPE>
__device__ __forceinline__ void agg_reduce(double (&v)[K * LPE], const int agg) {
    for (int mk = LPE; mk < LPE * agg; mk <<= 1) {
#pragma unroll
        for (int r = 0; r < K; ++r) v[r] += shfl_xor_d(v[r], mk);
    }
}
__device__ __forceinline__ double agg_reduce1(double v, const int agg, const int LPE) {
    for (int mk = LPE; mk < LPE * agg; mk <<= 1) v += shfl_xor_d(v, mk);
    return v;
}

struct BarrierOut {
    double F, gq[3], gs, Hqq[3][3], Hqs[3], Hss;
    bool feasible;
};

// mu(p): weight of the extra -log s term.  [U] (upstream convex_Euclidian_power is not vendored): 0 for p = 1
// (-log(s^2 - |q|^2), the Lorentz-cone barrier) and for p = 2 (-log(s - |q|^2), the paraboloid epigraph - both are
// self-concordant as they stand), 1 for 1 < p < 2, 2 for p > 2.  The ONE place the kernels define it; the oracle's
// _mu (oracle/mgb_oracle.py) is its twin.
__device__ __forceinline__ double barrier_mu(double p) { return (p == 1.0 || p == 2.0) ? 0.0 : ((p < 2.0) ? 1.0 : 2.0); }

// Euclidian power cone barrier on (q_0..q_{D-1}, s): F = -log(s^a - |q|^2) - mu log s, a = 2/p.
template <int D, bool WANT_F, bool WANT_D>
__device__ __forceinline__ void barrier_eval(const double (&q)[D], double s, double p, BarrierOut& o) {
    const double a = 2.0 / p;
    const double mu = barrier_mu(p);
    // p is uniform over the grid: the branches below are uniform.  The common cases p = 1 and p = 2 need neither pow
    // nor (mu = 0) log s nor - for p = 1 - the reciprocal of s; a division / a log cost 40-80 instructions each.
    const bool need_is = (mu != 0.0) || (p != 1.0);
    double sa, sa1, sa2;  // s^a, s^(a-1), s^(a-2)
    const double is = need_is ? 1.0 / s : 0.0;
    if (p == 1.0) { sa = s * s; sa1 = s; sa2 = 1.0; }
    else if (p == 2.0) { sa = s; sa1 = 1.0; sa2 = is; }
    else { sa = pow(s, a); sa1 = sa * is; sa2 = sa1 * is; }
    double qq = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) qq += q[j] * q[j];
    const double phi = sa - qq;
    o.feasible = (s > 0.0) && (phi > 0.0);
    if (WANT_F) {
        double f = -log(phi);
        if (mu != 0.0) f -= mu * log(s);
        o.F = o.feasible ? f : __longlong_as_double(0x7ff0000000000000LL);
    }
    if (WANT_D) {
        const double ip = 1.0 / phi, ip2 = ip * ip;
        const double ds = a * sa1;
#pragma unroll
        for (int j = 0; j < D; ++j) {
            o.gq[j] = 2.0 * q[j] * ip;
            o.Hqs[j] = -2.0 * q[j] * ds * ip2;
#pragma unroll
            for (int j2 = 0; j2 < D; ++j2) o.Hqq[j][j2] = 4.0 * q[j] * q[j2] * ip2 + (j == j2 ? 2.0 * ip : 0.0);
        }
        o.gs = -ds * ip - mu * is;
        o.Hss = -a * (a - 1.0) * sa2 * ip + ds * ds * ip2 + mu * is * is;
    }
}

// FLAGS bits: 1 objective, 2 gradient, 4 Hessian, 8 store Dz
// Per-point work of one element group (LPE lanes).  Writes the element's gradient record to `rel`
// and its slot record to `sel` (entry r of a butterfly-reduced block is stored at  off + r*LPE + lane).  Returns this
// thread's objective / <c,Dz> / infeasibility partials.  Every lane of the group must call it.
// MODE: 0 = one cone on (grad u, s); 1 = feasibility phase, cone on (grad u, s + tau) plus -log(1 + tau);
//       2 = two cones (upstream parabolic_solve): A on (u, s1) with exponent p2, B on (grad u, s2) with p.
//       Modes 1 and 2 share the three-variable operator table [u.id; u.d*; v1.id; v2.id] and slot layout.
template <int B, int D, int MODE, bool FINE, int FLAGS>
__device__ __forceinline__ void element_body(const ElemParams& P, const int64_t e, const int l, double* __restrict__ sel,
                                             double* __restrict__ rel, double& v0, double& v1, double& v2) {
    constexpr bool SLACK = MODE == 1, TWO = MODE == 2, THREE = MODE != 0;
    constexpr int LPE = Pow2Ceil<B>::value;
    constexpr int ND = D + 2 + (THREE ? 1 : 0);
    constexpr int NU = 2 + (THREE ? 1 : 0);
    constexpr bool WF = (FLAGS & 1) != 0, WG = (FLAGS & 2) != 0, WH = (FLAGS & 4) != 0, WDZ = (FLAGS & 8) != 0;
    constexpr int NTRI = (B * (B + 1) / 2 + LPE - 1) / LPE * LPE;
    constexpr int NFULL = (B * B + LPE - 1) / LPE * LPE;

    const bool act_e = e < P.E;
    const bool act = act_e && (l < B);
    const int agg = FINE ? 1 : P.agg;                                   // see agg_reduce
    const bool wr = act_e && (FINE || ((threadIdx.x & 31) / LPE) % agg == 0);   // this lane group stores the (summed) record
    const int64_t i = act ? e * B + l : 0;
    const int64_t n = P.nloc;

    // ---- loads.  Order matters (in-order issue): first the dof indices (their consumer, the gather
    // of s, comes last), then every independent stream, so one memory latency covers all of them.
    int32_t col[NU];
#pragma unroll
    for (int v = 0; v < NU; ++v) col[v] = act_e ? __ldg(&P.lcols[(e * NU + v) * LPE + l]) : -1;
    constexpr int RWF = D * B + 1 + NU + 1, RWC = D * B + 1 + NU * B;
    constexpr int RW = ((FINE ? RWF : RWC) + 1) / 2 * 2;
    double rec[RW];
    {
        const double2* __restrict__ rp = reinterpret_cast<const double2*>(P.prec + i * RW);
#pragma unroll
        for (int j = 0; j < RW / 2; ++j) {
            const double2 t2 = act ? __ldg(rp + j) : make_double2(0.0, 0.0);
            rec[2 * j] = t2.x;
            rec[2 * j + 1] = t2.y;
        }
    }
    double cc[ND], dz[ND];
#pragma unroll
    for (int k = 0; k < ND; ++k) {
        cc[k] = act ? __ldg(&P.c[(int64_t)k * n + i]) : 0.0;
        dz[k] = (act && P.Dz0) ? __ldg(&P.Dz0[(int64_t)k * n + i]) : 0.0;
    }
    // ---- gather the element's unknowns: lane q holds z[var][q]
    double zl[NU];
#pragma unroll
    for (int v = 0; v < NU; ++v) zl[v] = (col[v] >= 0) ? __ldg(&P.s[col[v]]) : 0.0;
    // ---- unpack the record
    double a[D][B];
#pragma unroll
    for (int k = 0; k < D; ++k)
#pragma unroll
        for (int q = 0; q < B; ++q) a[k][q] = rec[k * B + q];
    const double wi = rec[D * B];
    double aid[FINE ? 1 : NU][FINE ? 1 : B];
    double oval[NU];
    int olq[NU];
    bool oh[NU];  // this point owns a column of variable v (false: eliminated dof, e.g. Dirichlet)
    if (FINE) {
        const unsigned long long lqbits = (unsigned long long)__double_as_longlong(rec[FINE ? D * B + 1 + NU : 0]);
#pragma unroll
        for (int v = 0; v < NU; ++v) {
            oval[v] = rec[FINE ? D * B + 1 + v : 0];
            olq[v] = act ? (int)((lqbits >> (8 * v)) & 0xFFull) : 255;
            oh[v] = olq[v] != 255;
            if (!oh[v]) { oval[v] = 0.0; olq[v] = 0; }
        }
    } else {
#pragma unroll
        for (int v = 0; v < NU; ++v)
#pragma unroll
            for (int q = 0; q < B; ++q) aid[FINE ? 0 : v][FINE ? 0 : q] = rec[FINE ? 0 : D * B + 1 + v * B + q];
    }
    // ---- apply_D: Dz = Dz0 + (D R) s
#pragma unroll
    for (int q = 0; q < B; ++q) {
        const double zq = shfl_d(zl[0], q, LPE);
#pragma unroll
        for (int k = 0; k < D; ++k) dz[1 + k] = fma(a[k][q], zq, dz[1 + k]);
        if (!FINE) dz[0] = fma(aid[0][FINE ? 0 : q], zq, dz[0]);
    }
    if (FINE) {
#pragma unroll
        for (int v = 0; v < NU; ++v) {
            const double zo = shfl_d(zl[v], olq[v], LPE);
            const int k = (v == 0) ? 0 : D + v;
            dz[k] = fma(oval[v], zo, dz[k]);
        }
    } else {
#pragma unroll
        for (int v = 1; v < NU; ++v)
#pragma unroll
            for (int q = 0; q < B; ++q) dz[D + v] = fma(aid[FINE ? 0 : v][FINE ? 0 : q], shfl_d(zl[v], q, LPE), dz[D + v]);
    }
    if (WDZ && act && P.Dz) {
#pragma unroll
        for (int k = 0; k < ND; ++k) P.Dz[(int64_t)k * n + i] = dz[k];
    }
    // ---- barrier at this point
    double qv[D];
#pragma unroll
    for (int j = 0; j < D; ++j) qv[j] = dz[1 + j];
    double sv = TWO ? dz[D + (THREE ? 2 : 1)] : dz[D + 1];
    if (SLACK) sv += dz[D + (THREE ? 2 : 1)];
    if (!act) { sv = 1.0;
#pragma unroll
        for (int j = 0; j < D; ++j) qv[j] = 0.0; }
    BarrierOut bo;
    barrier_eval<D, WF, (WG || WH)>(qv, sv, P.p, bo);
    // feasibility phase: the slack tau is bounded below by the extra barrier -log(1 + tau)
    double tau1 = 1.0, itau = 1.0;
    if (SLACK) {
        tau1 = act ? 1.0 + dz[D + (THREE ? 2 : 1)] : 1.0;
        itau = 1.0 / tau1;
        if (WF) bo.F = (tau1 > 0.0) ? bo.F - log(tau1) : __longlong_as_double(0x7ff0000000000000LL);
        bo.feasible = bo.feasible && (tau1 > 0.0);
    }
    // second cone A on (u, s1)
    BarrierOut ba;
    if (TWO) {
        double qa[1] = {act ? dz[0] : 0.0};
        const double sa = act ? dz[D + 1] : 1.0;
        barrier_eval<1, WF, (WG || WH)>(qa, sa, P.p2, ba);
        if (WF) bo.F += ba.F;
        bo.feasible = bo.feasible && ba.feasible;
    }

    // ---- objective / feasibility partials of this point (reduced by the caller in a fixed order)
    {
        double cd = 0.0;
#pragma unroll
        for (int k = 0; k < ND; ++k) cd = fma(cc[k], dz[k], cd);
        const bool act_s = act && e < P.Eprim;
        v0 = (WF && act_s) ? wi * bo.F : 0.0;
        v1 = act_s ? wi * cd : 0.0;
        v2 = (act_s && !bo.feasible) ? 1.0 : 0.0;
    }
    // an infeasible point produces NaN/Inf values; they flow to the outputs as data (the caller
    // reads all_finite, like amgb_all_isfinite in the reference, src:121-133)

    // ---- gradient: r[var][q] = sum_points sum_k a_k[q] * w (F1_k + t c_k)
    if (WG) {
        double gy[ND];
        gy[0] = wi * ((TWO ? ba.gq[0] : 0.0) + P.t * cc[0]);
#pragma unroll
        for (int j = 0; j < D; ++j) gy[1 + j] = wi * (bo.gq[j] + P.t * cc[1 + j]);
        gy[D + 1] = wi * ((TWO ? ba.gs : bo.gs) + P.t * cc[D + 1]);
        if (SLACK) gy[D + (THREE ? 2 : 1)] = wi * (bo.gs - itau + P.t * cc[D + (THREE ? 2 : 1)]);
        if (TWO) gy[D + (THREE ? 2 : 1)] = wi * (bo.gs + P.t * cc[D + (THREE ? 2 : 1)]);
        double ru[LPE];
#pragma unroll
        for (int q = 0; q < LPE; ++q) {
            double r = 0.0;
            if (q < B) {
#pragma unroll
                for (int k = 0; k < D; ++k) r = fma(a[k][q < B ? q : 0], gy[1 + k], r);
                if (FINE) r += (q == olq[0]) ? oval[0] * gy[0] : 0.0;
                else r = fma(aid[0][FINE ? 0 : (q < B ? q : 0)], gy[0], r);
            }
            ru[q] = r;
        }
        group_reduce<LPE, LPE>(ru, l);
        if (!FINE) ru[0] = agg_reduce1(ru[0], agg, LPE);
        if (wr) rel[l] = ru[0];
        if (FINE) {
#pragma unroll
            for (int v = 1; v < NU; ++v)
                if (act && oh[v]) rel[v * LPE + olq[v]] = oval[v] * gy[D + v];
        } else {
#pragma unroll
            for (int v = 1; v < NU; ++v) {
                double rs[LPE];
#pragma unroll
                for (int q = 0; q < LPE; ++q) rs[q] = (q < B) ? aid[FINE ? 0 : v][FINE ? 0 : (q < B ? q : 0)] * gy[D + v] : 0.0;
                group_reduce<LPE, LPE>(rs, l);
                rs[0] = agg_reduce1(rs[0], agg, LPE);
                if (wr) rel[v * LPE + l] = rs[0];
            }
        }
    }
    constexpr bool CAN_MMA = !FINE && MODE == 0 && B == 7 && D == 2;
    if (WH && CAN_MMA && P.mma) {
        // ---- Hessian of a coarse fem2d level on the FP64 tensor cores.  The 28 points of the warp's four sibling elements
        // share their 7 + 7 coarse dofs, so  uu = sum_pt sum_jj' a_j' (w F2_jj') a_j',  us = sum_pt (a' w F2_qs) i_s',
        // ss = sum_pt i_s (w F2_ss) i_s'  are three 8 x 8 products with K = the warp's points: 8 steps of
        // mma.m8n8k4 over four lanes' points each.  Lane (g, t) of a step needs "dof g of point 4 step + t": the operator
        // values come from that point's record (an L1 hit - its own lane loaded the record a moment ago), the point's six
        // barrier values by shuffle from its lane.  Replaces 112 products per lane and a 105-value shuffle butterfly.
        const int lane = threadIdx.x & 31, fg = lane >> 2, ft = lane & 3;
        const int64_t ew = e - (lane >> 3);   // first element of this warp
        constexpr int RWC = D * B + 1 + NU * B, RWm = (RWC + 1) / 2 * 2;
        const double h00 = wi * bo.Hqq[0][0], h01 = wi * bo.Hqq[0][1], h11 = wi * bo.Hqq[1][1];
        const double hq0 = wi * bo.Hqs[0], hq1 = wi * bo.Hqs[1], hss = wi * bo.Hss;
        double cuu[2] = {0.0, 0.0}, cus[2] = {0.0, 0.0}, css[2] = {0.0, 0.0};
#pragma unroll
        for (int st = 0; st < 8; ++st) {
            const int src = 4 * st + ft;                       // lane of the point this k index stands for
            const int64_t pe = ew + (src >> 3);
            const int pl = src & 7;
            const bool ok = pl < B && pe < P.E && fg < B;
            const double* rp = P.prec + (pe * B + pl) * RWm;
            const double a0 = ok ? __ldg(rp + fg) : 0.0, a1 = ok ? __ldg(rp + B + fg) : 0.0;
            const double as = ok ? __ldg(rp + D * B + 1 + B + fg) : 0.0;
            const double H00 = shfl_d(h00, src, 32), H01 = shfl_d(h01, src, 32), H11 = shfl_d(h11, src, 32);
            const double Q0 = shfl_d(hq0, src, 32), Q1 = shfl_d(hq1, src, 32), SS = shfl_d(hss, src, 32);
            const double T0 = H00 * a0 + H01 * a1, T1 = H01 * a0 + H11 * a1;
            const double bsv = Q0 * a0 + Q1 * a1, hsv = SS * as;
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(cuu[0]), "+d"(cuu[1]) : "d"(a0), "d"(T0));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(cuu[0]), "+d"(cuu[1]) : "d"(a1), "d"(T1));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(cus[0]), "+d"(cus[1]) : "d"(bsv), "d"(as));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(css[0]), "+d"(css[1]) : "d"(hsv), "d"(as));
        }
        if (ew < P.E) {   // lane holds C[g][2t], C[g][2t + 1] of each block
            double* rec = sel + fg * 8 + 2 * ft;
            *reinterpret_cast<double2*>(rec + P.off_uu) = make_double2(cuu[0], cuu[1]);
            *reinterpret_cast<double2*>(rec + P.off_us) = make_double2(cus[0], cus[1]);
            *reinterpret_cast<double2*>(rec + P.off_ss) = make_double2(css[0], css[1]);
        }
    } else if (WH) {
    // ---- Hessian: element-local blocks of sum_jk a_j' (w Y_jk) a_k
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
        double v[NTRI];
#pragma unroll
        for (int r = 0; r < NTRI; ++r) v[r] = 0.0;
#pragma unroll
        for (int q = 0; q < B; ++q)
#pragma unroll
            for (int q2 = q; q2 < B; ++q2) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < D; ++j) acc = fma(a[j][q], T[j][q2], acc);
                v[q * B - q * (q - 1) / 2 + (q2 - q)] = acc;
            }
        if (TWO) {  // cone A couples u with itself through u.id
            const double huu = wi * ba.Hqq[0][0];
#pragma unroll
            for (int q = 0; q < B; ++q) {
                if (FINE) {
                    v[q * B - q * (q - 1) / 2] += (oh[0] && q == olq[0]) ? huu * oval[0] * oval[0] : 0.0;
                } else {
#pragma unroll
                    for (int q2 = q; q2 < B; ++q2)
                        v[q * B - q * (q - 1) / 2 + (q2 - q)] += huu * aid[0][FINE ? 0 : q] * aid[0][FINE ? 0 : q2];
                }
            }
        }
        group_reduce<NTRI, LPE>(v, l);
        if (!FINE) agg_reduce<NTRI / LPE, LPE>(v, agg);
        if (wr) {
#pragma unroll
            for (int r = 0; r < NTRI / LPE; ++r) sel[P.off_uu + r * LPE + l] = v[r];
        }
    }
    // u-s (and u-slack) blocks
    double bs[B];
#pragma unroll
    for (int q = 0; q < B; ++q) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) acc = fma(a[j][q], wi * bo.Hqs[j], acc);
        bs[q] = acc;
    }
    const double vss = wi * bo.Hss;
    const double vtt = SLACK ? vss + wi * itau * itau : vss;  // last variable's own curvature (slack: incl. -log(1+tau))
    const double haus = TWO ? wi * ba.Hqs[0] : 0.0;           // cone A: u x s1 and s1 x s1
    const double hass = TWO ? wi * ba.Hss : 0.0;
    constexpr int VT = THREE ? 2 : 0;
    if (FINE) {
        if (act && oh[1]) {
#pragma unroll
            for (int q = 0; q < B; ++q)
                sel[P.off_us + q * LPE + olq[1]] = TWO ? ((oh[0] && q == olq[0]) ? haus * oval[0] * oval[1] : 0.0) : bs[q] * oval[1];
            sel[P.off_ss + olq[1]] = (TWO ? hass : vss) * oval[1] * oval[1];
        }
        if (THREE && act && oh[VT]) {
#pragma unroll
            for (int q = 0; q < B; ++q) sel[P.off_ut + q * LPE + olq[VT]] = bs[q] * oval[VT];
            if (oh[1]) sel[P.off_st + l] = TWO ? 0.0 : vss * oval[1] * oval[VT];
            sel[P.off_tt + olq[VT]] = vtt * oval[VT] * oval[VT];
        }
    } else {
#pragma unroll
        for (int v2 = 1; v2 < NU; ++v2) {  // u x {v1, v2}
            double v[NFULL];
#pragma unroll
            for (int r = 0; r < NFULL; ++r) v[r] = 0.0;
#pragma unroll
            for (int q = 0; q < B; ++q)
#pragma unroll
                for (int q2 = 0; q2 < B; ++q2)
                    v[q * B + q2] = ((TWO && v2 == 1) ? haus * aid[0][FINE ? 0 : q] : bs[q]) * aid[FINE ? 0 : v2][FINE ? 0 : q2];
            group_reduce<NFULL, LPE>(v, l);
            agg_reduce<NFULL / LPE, LPE>(v, agg);
            const int off = (v2 == 1) ? P.off_us : P.off_ut;
            if (wr) {
#pragma unroll
                for (int r = 0; r < NFULL / LPE; ++r) sel[off + r * LPE + l] = v[r];
            }
        }
#pragma unroll
        for (int v1 = 1; v1 < NU; ++v1) {  // symmetric diagonal blocks of v1, v2
            double v[NTRI];
#pragma unroll
            for (int r = 0; r < NTRI; ++r) v[r] = 0.0;
            const double hd = (v1 == 2) ? vtt : (TWO ? hass : vss);
#pragma unroll
            for (int q = 0; q < B; ++q)
#pragma unroll
                for (int q2 = q; q2 < B; ++q2)
                    v[q * B - q * (q - 1) / 2 + (q2 - q)] = hd * aid[FINE ? 0 : v1][FINE ? 0 : q] * aid[FINE ? 0 : v1][FINE ? 0 : q2];
            group_reduce<NTRI, LPE>(v, l);
            agg_reduce<NTRI / LPE, LPE>(v, agg);
            const int off = (v1 == 1) ? P.off_ss : P.off_tt;
            if (wr) {
#pragma unroll
                for (int r = 0; r < NTRI / LPE; ++r) sel[off + r * LPE + l] = v[r];
            }
        }
        if (THREE) {  // v1 x v2 full block (two cones: structurally present, identically zero)
            double v[NFULL];
#pragma unroll
            for (int r = 0; r < NFULL; ++r) v[r] = 0.0;
            if (!TWO) {
#pragma unroll
                for (int q = 0; q < B; ++q)
#pragma unroll
                    for (int q2 = 0; q2 < B; ++q2)
                        v[q * B + q2] = vss * aid[FINE ? 0 : 1][FINE ? 0 : q] * aid[FINE ? 0 : VT][FINE ? 0 : q2];
                group_reduce<NFULL, LPE>(v, l);
                agg_reduce<NFULL / LPE, LPE>(v, agg);
            }
            if (wr) {
#pragma unroll
                for (int r = 0; r < NFULL / LPE; ++r) sel[P.off_st + r * LPE + l] = v[r];
            }
        }
    }
    }  // WH
}

// fixed-order block reduction of the three scalar partials -> part[blockIdx]
__device__ __forceinline__ void block_scalars(double v0, double v1, double v2, double* __restrict__ part) {
#pragma unroll
    for (int mk = 16; mk >= 1; mk >>= 1) {
        v0 += shfl_xor_d(v0, mk);
        v1 += shfl_xor_d(v1, mk);
        v2 += shfl_xor_d(v2, mk);
    }
    __shared__ double red[3][32];
    const int wid = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { red[0][wid] = v0; red[1][wid] = v1; red[2][wid] = v2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        const int nw = (blockDim.x + 31) >> 5;
        for (int r = 0; r < nw; ++r) { s0 += red[0][r]; s1 += red[1][r]; s2 += red[2][r]; }
        part[(int64_t)blockIdx.x * 4 + 0] = s0;
        part[(int64_t)blockIdx.x * 4 + 1] = s1;
        part[(int64_t)blockIdx.x * 4 + 2] = s2;
    }
}

// Two-stage path, stage 1: slot / gradient records to global memory (replayed by gather_kernel).
template <int B, int D, int MODE, bool FINE, int FLAGS>
__global__ void __launch_bounds__(MGB_ELEM_THREADS, (FLAGS & 6) ? MGB_ELEM_MINBLOCKS : MGB_ELEM_MINBLOCKS_F0) element_kernel(const ElemParams P) {
    constexpr int LPE = Pow2Ceil<B>::value;
    constexpr int NU = 2 + (MODE != 0 ? 1 : 0);
    pdl_launch_dependents();
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = tid / LPE;
    const int l = (int)(tid % LPE);
    double v0, v1, v2;
    const int64_t er = FINE ? e : e / P.agg;   // record of the element's aggregation group
    element_body<B, D, MODE, FINE, FLAGS>(P, e, l, P.sel + er * (int64_t)P.NS, P.rel + er * (NU * LPE), v0, v1, v2);
    block_scalars(v0, v1, v2, P.part);
}

// Cross-rank sum of the three objective scalars without a fence or a collective: every 64-bit word that crosses
// NVLink carries half a double and the 32-bit epoch of the assembly it belongs to, so a single relaxed system-scope
// store is self-validating (the protocol idea of NCCL's LL mode).  Word w of source rank q of epoch parity par sits at
// win[(par * DIST_LL_RANKS + q) * 8 + w] in every rank's window.  Two parities: a rank can run at most one epoch
// ahead of a peer (it cannot finish epoch k+1 without the peer's k+1 words, which the peer sends after reading k).
constexpr int DIST_LL_RANKS = 16;
struct DistScal {
    int rank, nranks;                           // nranks <= 1: single rank, nothing crosses
    unsigned epoch;                             // never 0
    unsigned long long* win[DIST_LL_RANKS];     // every rank's window (own included), peer-mapped
    int publish_only;                           // split mode (mgb_dist_begin): store the partials, do not wait
    unsigned long long timeout_ns;
    int* err;                                   // set to 1 when a peer's words did not arrive in time
};

__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// this rank's three partial sums -> every rank's window (threads 0 .. 6*nranks-1 of the block; one word each)
__device__ __forceinline__ void dist_publish(const DistScal& D, const double* __restrict__ tot /* 3, shared */) {
    const int tid = threadIdx.x;
    if (tid < 6 * D.nranks) {
        const int p = tid / 6, w = tid % 6;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(tot[w >> 1]);
        const unsigned long long half = (w & 1) ? (bits >> 32) : (bits & 0xFFFFFFFFull);
        st_relaxed_sys(D.win[p] + (((D.epoch & 1u) * DIST_LL_RANKS + D.rank) * 8 + w), ((unsigned long long)D.epoch << 32) | half);
    }
}
// every rank's partial sums of this epoch -> sums in rank order (identical on every rank).  Returns false on time-out.
__device__ __forceinline__ bool dist_collect(const DistScal& D, double* __restrict__ sum /* 3, thread 0 only */) {
    __shared__ unsigned int half_s[DIST_LL_RANKS * 6];
    __shared__ int bad_s;
    const int tid = threadIdx.x;
    if (tid == 0) bad_s = 0;
    __syncthreads();
    if (tid < 6 * D.nranks) {
        const int q = tid / 6, w = tid % 6;
        const unsigned long long* src = D.win[D.rank] + (((D.epoch & 1u) * DIST_LL_RANKS + q) * 8 + w);
        const unsigned long long t0 = global_timer_ns();
        unsigned long long word = ld_relaxed_sys(src);
        while ((unsigned)(word >> 32) != D.epoch) {
            if (global_timer_ns() - t0 > D.timeout_ns) { bad_s = 1; break; }
            word = ld_relaxed_sys(src);
        }
        half_s[tid] = (unsigned)(word & 0xFFFFFFFFull);
    }
    __syncthreads();
    if (tid == 0) {
        sum[0] = sum[1] = sum[2] = 0.0;
        for (int q = 0; q < D.nranks; ++q)
            for (int k = 0; k < 3; ++k)
                sum[k] += __longlong_as_double((long long)(((unsigned long long)half_s[q * 6 + 2 * k + 1] << 32) | half_s[q * 6 + 2 * k]));
        if (bad_s && D.err) atomicExch(D.err, 1);
    }
    return bad_s == 0;   // valid on thread 0 (it wrote and read bad_s around barriers); others do not use it
}

__device__ __forceinline__ void write_scalars(double s0, double s1, double s2, double t, bool ok, double* __restrict__ scal) {
    if (!scal) return;
    if (ok) {
        scal[0] = s0 + t * s1;
        scal[1] = (s2 == 0.0) ? 1.0 : 0.0;
        scal[2] = s1;
        scal[3] = s2;
    } else {   // a peer never delivered: poison the result instead of returning a partial sum as if it were the objective
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        scal[0] = nan; scal[1] = 0.0; scal[2] = nan; scal[3] = -1.0;
    }
}

// one block of 256 threads folds the per-block scalar partials in a fixed order -> {f0, all_finite, <c,Dz>_w, count};
// with several ranks (dist.nranks > 1) the three sums first cross the ranks (dist_publish / dist_collect)
__device__ __forceinline__ void fold_scalars_block(const double* __restrict__ part, int64_t nparts, double t, double* __restrict__ scal,
                                                   const DistScal* dist = nullptr) {
    __shared__ double sh[3][256];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int64_t r = threadIdx.x; r < nparts; r += blockDim.x) {
        s0 += part[r * 4 + 0];
        s1 += part[r * 4 + 1];
        s2 += part[r * 4 + 2];
    }
    sh[0][threadIdx.x] = s0; sh[1][threadIdx.x] = s1; sh[2][threadIdx.x] = s2;
    __syncthreads();
    for (int st = blockDim.x / 2; st >= 1; st >>= 1) {
        if ((int)threadIdx.x < st) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + st];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + st];
            sh[2][threadIdx.x] += sh[2][threadIdx.x + st];
        }
        __syncthreads();
    }
    if (dist && dist->nranks > 1) {
        __shared__ double tot[3];
        if (threadIdx.x == 0) { tot[0] = sh[0][0]; tot[1] = sh[1][0]; tot[2] = sh[2][0]; }
        __syncthreads();
        dist_publish(*dist, tot);
        if (dist->publish_only) return;
        double sum[3];
        const bool ok = dist_collect(*dist, sum);
        if (threadIdx.x == 0) write_scalars(sum[0], sum[1], sum[2], t, ok, scal);
        return;
    }
    if (threadIdx.x == 0) write_scalars(sh[0][0], sh[1][0], sh[2][0], t, true, scal);
}

struct GatherParams {
    DistScal dist;           // cross-rank sum of the scalars (nranks <= 1: none)
    int64_t nnzH, m;
    const int2* h_src2;       // per entry: {first, second} contribution slot; second = -1 if none;
                              // first < 0 : entry -1-first of the long list (more than two contributions)
    const int64_t* h_lptr;    // long list pointers
    const int32_t* h_lidx;    // long list slots
    const int32_t* h_lt;      // long list: output entry of each list
    int64_t n_long, nblk_l;
    const int64_t* g_cptr;
    const int32_t* g_cidx;
    const double* sel;
    const double* rel;
    double* hval;
    double* grad;
    const double* part;
    int64_t nparts;
    double* scal;  // {f0, all_finite, cdot, nonfinite count}
    double t;
    int want_h, want_g;
    int64_t nblk_h, nblk_g;
};

#ifndef MGB_GATHER_UNROLL
#define MGB_GATHER_UNROLL 4
#endif
constexpr int GATHER_UNROLL = MGB_GATHER_UNROLL;

// Replays the frozen contribution lists: blocks [0,nblk_h) produce Hessian values (GATHER_UNROLL
// entries per thread, two-deep dependent loads), blocks [nblk_h, nblk_h+nblk_g) the gradient, the
// last block folds the scalar partials in a fixed order.
// sum of src[idx[c]] for c in [c0, c1) in list order; four index loads, then four value loads in flight
// (the plain loop is a chain of 2 * (c1 - c0) dependent loads)
__device__ __forceinline__ double list_sum(const double* __restrict__ src, const int32_t* __restrict__ idx, int64_t c0, int64_t c1) {
    double acc = 0.0;
    for (int64_t c = c0; c < c1; c += 4) {
        int32_t k[4];
        double v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) k[j] = (c + j < c1) ? __ldg(&idx[c + j]) : -1;
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (k[j] >= 0) ? src[k[j]] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (k[j] >= 0) acc += v[j];
    }
    return acc;
}

static __global__ void __launch_bounds__(256) gather_kernel(const GatherParams P) {
    // launch order [long lists | gradient | two-source Hessian entries | scalars]: the list-walking blocks have the
    // longest dependent-load chains, so they start first and overlap the bulk instead of forming the kernel's tail
    // Several ranks: the scalar block goes FIRST instead of last - it publishes this rank's partial sums the moment the
    // element kernel has finished and then waits for the peers' words while the rest of the grid gathers, so the
    // NVLink latency and up to a whole gather of skew between the ranks are hidden behind the kernel's own work.
    int64_t b = blockIdx.x;
    if (P.dist.nranks > 1) b = (b == 0) ? (int64_t)gridDim.x - 1 : b - 1;
    {
        const int64_t nlg = P.nblk_l + P.nblk_g;
        if (b < nlg) b += P.nblk_h;
        else if (b < nlg + P.nblk_h) b -= nlg;
    }
    if (b < P.nblk_h) {
        const int64_t base = b * (256 * GATHER_UNROLL) + threadIdx.x;
        int2 src[GATHER_UNROLL];
#pragma unroll
        for (int j = 0; j < GATHER_UNROLL; ++j) {
            const int64_t t = base + (int64_t)j * 256;
            src[j] = (t < P.nnzH) ? __ldg(&P.h_src2[t]) : make_int2(0, -1);
        }
        pdl_wait_primary();
        double v0[GATHER_UNROLL], v1[GATHER_UNROLL];
#pragma unroll
        for (int j = 0; j < GATHER_UNROLL; ++j) {
            v0[j] = (src[j].x >= 0) ? P.sel[src[j].x] : 0.0;
            v1[j] = (src[j].y >= 0) ? P.sel[src[j].y] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < GATHER_UNROLL; ++j) {
            const int64_t t = base + (int64_t)j * 256;
            if (t >= P.nnzH || src[j].x < 0) continue;  // long entries are finished by their own blocks
            P.hval[t] = v0[j] + v1[j];
        }
        return;
    }
    if (b < P.nblk_h + P.nblk_l) {
        // entries with more than two contributions (vertex diagonals ...): thread per entry
        const int64_t li = (b - P.nblk_h) * 256 + threadIdx.x;
        int64_t c0 = 0, c1 = 0;
        int32_t dst = 0;
        if (li < P.n_long) { c0 = __ldg(&P.h_lptr[li]); c1 = __ldg(&P.h_lptr[li + 1]); dst = __ldg(&P.h_lt[li]); }
        pdl_wait_primary();
        if (li < P.n_long) P.hval[dst] = list_sum(P.sel, P.h_lidx, c0, c1);
        return;
    }
    if (b < P.nblk_h + P.nblk_l + P.nblk_g) {
        const int64_t a = (b - P.nblk_h - P.nblk_l) * 256 + threadIdx.x;
        int64_t c0 = 0, c1 = 0;
        if (a < P.m) { c0 = __ldg(&P.g_cptr[a]); c1 = __ldg(&P.g_cptr[a + 1]); }
        pdl_wait_primary();
        if (a < P.m) P.grad[a] = list_sum(P.rel, P.g_cidx, c0, c1);
        return;
    }
    // scalar block
    pdl_wait_primary();
    fold_scalars_block(P.part, P.nparts, P.t, P.scal, &P.dist);
}

// Coarse multigrid levels: few outputs with long contribution lists.  One launch serves up to two lists (Hessian
// values and gradient) plus the scalar fold, so a coarse assembly is element kernel + one or two of these launches
// instead of five small ones.  Per list: LPE lanes per output - 32 for lists of hundreds of contributions, 8 when a
// list has a few dozen (a full warp per entry would leave most lanes idle and quadruple the number of warps that
// cycle through the SMs); idx == nullptr sums the contiguous range src[ptr[o] .. ptr[o+1]) (stage 2 of a chunked
// list, see ChunkedList in mgb_b200.cu).
struct WarpList {
    int64_t nout;          // 0: unused
    const int64_t* ptr;
    const int32_t* idx;    // nullptr: contiguous
    const double* src;
    double* dst;
    int lpe;               // 8 or 32
    int64_t nblk;          // blocks of 256 threads
};
struct WarpGatherParams {
    WarpList a, b;
    const double* part;    // scalar partials (folded by the last block when scal != nullptr)
    int64_t nparts;
    double t;
    double* scal;
};

template <int LPE>
__device__ __forceinline__ void warp_list_sum(const WarpList& W, const int64_t blk) {
    const int64_t o = (blk * 256 + threadIdx.x) / LPE;
    const int lane = threadIdx.x % LPE;
    const bool act = o < W.nout;   // whole groups: every lane of a warp reaches the shuffles
    double acc = 0.0;
    if (act) {
        const int64_t c0 = W.ptr[o], c1 = W.ptr[o + 1];
        if (W.idx) for (int64_t c = c0 + lane; c < c1; c += LPE) acc += W.src[W.idx[c]];
        else for (int64_t c = c0 + lane; c < c1; c += LPE) acc += W.src[c];
    }
#pragma unroll
    for (int mk = LPE / 2; mk >= 1; mk >>= 1) acc += shfl_xor_d(acc, mk);
    if (act && lane == 0) W.dst[o] = acc;
}

static __global__ void __launch_bounds__(256) warp_gather_kernel(const WarpGatherParams P) {
    pdl_launch_dependents();   // a second stage may be scheduled while this grid drains
    pdl_wait_primary();        // records / partial sums of the previous kernel (no-op without the PDL attribute)
    const int64_t b = blockIdx.x;
    if (b < P.a.nblk) {
        if (P.a.lpe == 8) warp_list_sum<8>(P.a, b); else warp_list_sum<32>(P.a, b);
    } else if (b < P.a.nblk + P.b.nblk) {
        if (P.b.lpe == 8) warp_list_sum<8>(P.b, b - P.a.nblk); else warp_list_sum<32>(P.b, b - P.a.nblk);
    } else {
        fold_scalars_block(P.part, P.nparts, P.t, P.scal);
    }
}

// ---------------------------------------------------------------- small utilities
static __global__ void isfinite_kernel(const double* __restrict__ v, int64_t len, int* __restrict__ flag) {
    int bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x)
        bad |= !isfinite(v[i]);
    if (__syncthreads_or(bad) && threadIdx.x == 0) *flag = 0;  // idempotent store, not an accumulation
}

// dot / sum / squared 2-norm / max |.| of device vectors (HPCVector dot, sum, norm: reference
// tools/profile_scaling.jl:89-109).  Deterministic: a fixed grid of REDUCE_BLOCKS blocks, each thread
// strides the vector in a fixed pattern, shuffle tree inside the block, and the block that retires last folds the
// REDUCE_BLOCKS partials in index order - the same bits on every run and for every launch configuration.
constexpr int REDUCE_BLOCKS = 296;   // 2 per SM on B200
static __global__ void __launch_bounds__(256) reduce_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t len,
                                                            int op, double* __restrict__ part, unsigned* __restrict__ ticket,
                                                            double* __restrict__ out) {
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        const double a = x[i];
        if (op == 0) acc = fma(a, y[i], acc);          // dot
        else if (op == 1) acc += a;                     // sum
        else if (op == 2) acc = fma(a, a, acc);         // |x|^2
        else acc = fmax(acc, fabs(a));                  // max |x|
    }
    __shared__ double sh[8];
    __shared__ bool last;
#pragma unroll
    for (int mk = 16; mk >= 1; mk >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, acc, mk);
        acc = (op == 3) ? fmax(acc, o) : acc + o;
    }
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = sh[0];
        for (int r = 1; r < 8; ++r) v = (op == 3) ? fmax(v, sh[r]) : v + sh[r];
        part[blockIdx.x] = v;
        __threadfence();
        last = (atomicAdd(ticket, 1u) == gridDim.x - 1);   // a ticket, not an accumulation of values
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double v = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) {
            const double pb = *((volatile double*)&part[b]);
            v = (op == 3) ? fmax(v, pb) : v + pb;
        }
        out[0] = v;
        *ticket = 0u;
    }
}

static __global__ void diag_scale_kernel(const double* __restrict__ w, const double* __restrict__ y, int64_t n,
                                  double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = w[i] * y[i];
}

// L2 flush between timed steps: mode 0 writes a buffer larger than L2 (leaves L2 full of DIRTY lines whose
// write-back then competes with the timed kernels), mode 1 reads it (evicts everything, leaves clean lines).
static __global__ void l2_flush_kernel(double* __restrict__ buf, int64_t len, double v, int mode) {
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        if (mode == 0) buf[i] = v; else acc += buf[i];
    }
    if (mode != 0 && acc == 1.2345e300) buf[0] = acc;  // keeps the loads alive
}

// y = alpha * A x + beta * y0, thread per row (operator / restriction rows are short)
static __global__ void __launch_bounds__(256) spmv_kernel(int64_t nrows, const int64_t* __restrict__ ptr,
                                                   const int32_t* __restrict__ idx, const double* __restrict__ val,
                                                   double alpha, const double* __restrict__ x, double beta,
                                                   const double* __restrict__ y0, double* __restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    double acc = 0.0;
    for (int64_t p = ptr[i]; p < ptr[i + 1]; ++p) acc = fma(val[p], __ldg(&x[idx[p]]), acc);
    y[i] = alpha * acc + ((y0 && beta != 0.0) ? beta * y0[i] : 0.0);
}

// rows of hundreds of entries and more (the restriction R' of a coarse level: a few rows fed by every fine unknown):
// one warp per chunk of a row (stage 1), then one warp per row over the partial sums (stage 2, with alpha / beta)
static __global__ void __launch_bounds__(256) spmv_chunk_kernel(int64_t nchunks, const int64_t* __restrict__ kptr,
                                                         const int32_t* __restrict__ idx, const double* __restrict__ val,
                                                         const double* __restrict__ x, double* __restrict__ part) {
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const bool act = wid < nchunks;
    double acc = 0.0;
    if (act)
        for (int64_t p = kptr[wid] + lane; p < kptr[wid + 1]; p += 32) acc = fma(val[p], __ldg(&x[idx[p]]), acc);
#pragma unroll
    for (int mk = 16; mk >= 1; mk >>= 1) acc += shfl_xor_d(acc, mk);
    if (act && lane == 0) part[wid] = acc;
}
static __global__ void __launch_bounds__(256) spmv_fold_kernel(int64_t nrows, const int64_t* __restrict__ pptr,
                                                        const double* __restrict__ part, double alpha, double beta,
                                                        const double* __restrict__ y0, double* __restrict__ y) {
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const bool act = wid < nrows;
    double acc = 0.0;
    if (act)
        for (int64_t p = pptr[wid] + lane; p < pptr[wid + 1]; p += 32) acc += part[p];
#pragma unroll
    for (int mk = 16; mk >= 1; mk >>= 1) acc += shfl_xor_d(acc, mk);
    if (act && lane == 0) y[wid] = alpha * acc + ((y0 && beta != 0.0) ? beta * y0[wid] : 0.0);
}

static __global__ void __launch_bounds__(256) gather_idx_kernel(const double* __restrict__ src, const int32_t* __restrict__ idx,
                                                         int64_t count, double* __restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) out[k] = src[idx[k]];
}

static __global__ void __launch_bounds__(256) scatter_add_idx_kernel(const double* __restrict__ src, const int32_t* __restrict__ idx,
                                                              int64_t count, double* __restrict__ dst) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) dst[idx[k]] += src[k];  // idx unique within a call: plain read-modify-write
}

// dst[k] = sum_{r in [ptr[k], ptr[k+1])} src[idx[r]]  (fixed order; thread per output)
static __global__ void __launch_bounds__(256) segsum_idx_kernel(const double* __restrict__ src, const int32_t* __restrict__ ptr,
                                                               const int32_t* __restrict__ idx, int64_t nout,
                                                               double* __restrict__ dst) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nout) return;
    double acc = 0.0;
    for (int r = __ldg(&ptr[k]); r < __ldg(&ptr[k + 1]); ++r) acc += src[__ldg(&idx[r])];
    dst[k] = acc;
}

// map_rows of the barrier over Dz rows (separately callable seam; reference src:161-170)
template <int D>
__global__ void map_barrier_kernel(const double* __restrict__ Dz, int64_t n, int ND, int slack, double p, int which,
                                   double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double q[D];
#pragma unroll
    for (int j = 0; j < D; ++j) q[j] = Dz[(int64_t)(1 + j) * n + i];
    double s = Dz[(int64_t)(D + 1) * n + i];
    if (slack) s += Dz[(int64_t)(D + 2) * n + i];
    BarrierOut bo;
    barrier_eval<D, true, true>(q, s, p, bo);
    double tau1 = 1.0;
    if (slack) {
        tau1 = 1.0 + Dz[(int64_t)(D + 2) * n + i];
        bo.F = (tau1 > 0.0) ? bo.F - log(tau1) : __longlong_as_double(0x7ff0000000000000LL);
    }
    if (which == 0) { out[i] = bo.F; return; }
    const int ns = slack ? 2 : 1;
    if (which == 1) {
        out[i] = 0.0;
        for (int j = 0; j < D; ++j) out[(int64_t)(1 + j) * n + i] = bo.gq[j];
        for (int r = 0; r < ns; ++r) out[(int64_t)(D + 1 + r) * n + i] = bo.gs - (r == 1 ? 1.0 / tau1 : 0.0);
        return;
    }
    for (int c = 0; c < ND * ND; ++c) out[(int64_t)c * n + i] = 0.0;
    for (int j = 0; j < D; ++j) {
        for (int j2 = 0; j2 < D; ++j2) out[(int64_t)((1 + j) * ND + 1 + j2) * n + i] = bo.Hqq[j][j2];
        for (int r = 0; r < ns; ++r) {
            out[(int64_t)((1 + j) * ND + D + 1 + r) * n + i] = bo.Hqs[j];
            out[(int64_t)((D + 1 + r) * ND + 1 + j) * n + i] = bo.Hqs[j];
        }
    }
    for (int r = 0; r < ns; ++r)
        for (int r2 = 0; r2 < ns; ++r2)
            out[(int64_t)((D + 1 + r) * ND + D + 1 + r2) * n + i] = bo.Hss + ((r == 1 && r2 == 1) ? 1.0 / (tau1 * tau1) : 0.0);
}

}  // namespace mgb
