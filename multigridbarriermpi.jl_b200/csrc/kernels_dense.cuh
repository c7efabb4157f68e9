// Stage 1 of an assembly on the DENSE path (plan_host.h ElementPlan::dense): coarse levels of large elements (fem3d
// Q3 hexahedra), where every fine quadrature point touches up to 64 unknowns per variable and a product-list replay
// would need ~1 G products per level.  Per chunk of <= 512 points of one group (one CTA of 256 threads):
//   * the group's 2 x 64 unknowns are gathered once;
//   * the points' dense operator rows ([point][dx dy dz u.id s.id][64 dofs], 2.5 KB per point) stream in tiles of 16
//     points through a two-stage shared-memory ring filled by 1-D bulk copies (TMA: cp.async.bulk + mbarrier);
//   * phase 1 (two points per warp): apply_D as 64-long dot products (lane = 2 dofs, shuffle reduction), the barrier
//     at the two points on lanes 0 and 1 side by side, w .* F1 / F2 into shared memory, objective partials;
//   * phase 2: T = (w F2_qq) A, bs = A' (w F2_qs), hs = (w F2_ss) I_s per point, then the three contractions
//     uu += A' T,  us += bs' I_s,  ss += I_s' hs  on the FP64 tensor cores (mma.sync m8n8k4 = SASS DMMA.8x8x4): warp w
//     owns rows [8w, 8w+8) of each 64 x 64 block (8 column tiles, 48 accumulators per lane); the one dense contraction
//     of this path, and the one place where tensor cores pay.  The gradient rides along (threads 0..127);
//   * the chunk's full blocks (3 x 64 x 64 doubles) and gradient record go to `sel` / `rel`; the ordinary gather
//     kernels replay them into the CSR values of R'HR and into g (contribution lists built at plan time).
// Fixed summation order (points in order inside a chunk, chunks in order in the gather): bit-reproducible.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"
#include "plan_host.h"

namespace mgb {

constexpr int DN = DENSE_NB;   // 64
constexpr int DS = DENSE_STRIDE;   // 68: row stride (doubles) of every operand row in shared AND global memory - rows k, k+1,
                                   // k+2, k+3 of an MMA fragment then fall on distinct 8-byte bank pairs (4t + g mod 16)
constexpr int DPT = 16;        // points per tile
constexpr int DNR = 5;         // rows per point: dx dy dz u.id s.id  (dim = 3)
constexpr int D_TILE_BYTES = DPT * DNR * DS * 8;   // 42.5 KB
constexpr int D_SMEM = 2 * D_TILE_BYTES            // record ring
                     + 2 * DN * 8                  // unknowns
                     + DPT * 16 * 8                // per point: w F2 (10), w (F1 + t c) (5), pad
                     + DPT * 3 * DS * 8            // T
                     + 2 * DPT * DS * 8            // bs, hs
                     + 64;                         // mbarriers

// D(8x8) += A(8x4) * B(4x8) in FP64 on the tensor cores (SASS DMMA.8x8x4).  Fragment layout (g = lane / 4, t = lane % 4):
// a = A[g][t], b = B[t][g], c0 / c1 = C[g][2t], C[g][2t + 1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct DenseParams {
    int64_t nchunks, nloc;
    const int64_t* chunk;   // [nchunks][3] = {group, p0, p1}
    const int32_t* gdof;    // [ngroups][2][DN]
    const double* rows;     // [nloc][DNR][DS]
    const double* w;        // nloc
    const double* s;        // m
    const double* Dz0;      // nloc x 5 or null
    const double* c;        // nloc x 5
    double t, p;
    double* sel;            // nchunks * 3 * DN * DN
    double* rel;            // nchunks * 2 * DN
    double* part;           // nchunks * 4
    double* Dz;             // nloc x 5 or null
};

__device__ __forceinline__ uint32_t d_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void d_mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(d_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void d_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(d_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool d_mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(d_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void d_bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d_smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(d_smem_u32(bar)) : "memory");
}

// FLAGS bits as in the element kernels: 1 objective, 2 gradient, 4 Hessian, 8 store Dz
template <int FLAGS>
__global__ void __launch_bounds__(256, 1) dense_element_kernel(const __grid_constant__ DenseParams P) {
    constexpr bool WF = (FLAGS & 1) != 0, WG = (FLAGS & 2) != 0, WH = (FLAGS & 4) != 0, WDZ = (FLAGS & 8) != 0;
    extern __shared__ __align__(128) unsigned char dsm[];
    double* tile0 = reinterpret_cast<double*>(dsm);
    double* zu = reinterpret_cast<double*>(dsm + 2 * D_TILE_BYTES);
    double* zs = zu + DN;
    double* pw = zs + DN;                  // [DPT][16]
    double* Tt = pw + DPT * 16;            // [DPT * 3][DS]
    double* bs = Tt + DPT * 3 * DS;        // [DPT][DS]
    double* hs = bs + DPT * DS;            // [DPT][DS]
    uint64_t* bars = reinterpret_cast<uint64_t*>(hs + DPT * DS);
    pdl_launch_dependents();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t ch = blockIdx.x;
    const int64_t g = P.chunk[3 * ch], p0 = P.chunk[3 * ch + 1], p1 = P.chunk[3 * ch + 2];
    const int64_t npts = p1 - p0, n = P.nloc;
    const int ntile = (int)((npts + DPT - 1) / DPT);
    if (tid < 2 * DN) {
        const int32_t a = __ldg(&P.gdof[g * (2 * DN) + tid]);
        zu[tid] = a >= 0 ? __ldg(&P.s[a]) : 0.0;   // zu, zs are contiguous
    }
    if (tid == 0) {
        d_mbar_init(bars + 0, 1); d_mbar_init(bars + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int tl) {   // tile tl -> stage tl & 1 (thread 0)
        const int64_t q0 = p0 + (int64_t)tl * DPT;
        const unsigned bytes = (unsigned)(min((int64_t)DPT, p1 - q0) * DNR * DS * 8);
        d_mbar_expect_tx(bars + (tl & 1), bytes);
        d_bulk_g2s(tile0 + (size_t)(tl & 1) * (DPT * DNR * DS), P.rows + q0 * (DNR * DS), bytes, bars + (tl & 1));
    };
    if (tid == 0 && ntile > 0) issue(0);

    // warp w owns output rows [8w, 8w + 8) of the three 64 x 64 blocks: 8 column tiles of 8, two accumulators per lane each
    const int fg = lane >> 2, ft = lane & 3;
    double uu[8][2], us[8][2], ss[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) uu[i][0] = uu[i][1] = us[i][0] = us[i][1] = ss[i][0] = ss[i][1] = 0.0;
    double gacc = 0.0, sc0 = 0.0, sc1 = 0.0, sc2 = 0.0;

    for (int tl = 0; tl < ntile; ++tl) {
        __syncthreads();   // the other stage and T / bs / hs / pw are free again
        if (tid == 0 && tl + 1 < ntile) issue(tl + 1);
        double* tile = tile0 + (size_t)(tl & 1) * (DPT * DNR * DS);
        const int npt = (int)min((int64_t)DPT, npts - (int64_t)tl * DPT);
        while (!d_mbar_try_wait(bars + (tl & 1), (unsigned)((tl >> 1) & 1))) {}
        if (npt < DPT)   // the tail of a partial tile was not loaded: it must not contribute (0 * stale NaN)
            for (int k = npt * DNR * DS + tid; k < DPT * DNR * DS; k += 256) tile[k] = 0.0;
        // ---- phase 1: apply_D + barrier, two points per warp.  All lanes form the dot products of both points (lane = 2
        // dofs, xor-shuffle sums leave the totals on every lane); then lane 0 evaluates the first point and lane 1 the
        // second side by side - the barrier is a ~400-instruction dependent chain, running the two as SIMD lanes
        // instead of one after the other halves the phase.
        {
            double d[2][5];
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
                const int pt = 2 * warp + pp;
                const bool actp = pt < npt;
                const double* tp = tile + pt * (DNR * DS);
#pragma unroll
                for (int r = 0; r < 4; ++r) d[pp][r] = actp ? tp[r * DS + lane] * zu[lane] + tp[r * DS + lane + 32] * zu[lane + 32] : 0.0;
                d[pp][4] = actp ? tp[4 * DS + lane] * zs[lane] + tp[4 * DS + lane + 32] * zs[lane + 32] : 0.0;
            }
#pragma unroll
            for (int mk = 16; mk >= 1; mk >>= 1)
#pragma unroll
                for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                    for (int r = 0; r < 5; ++r) d[pp][r] += shfl_xor_d(d[pp][r], mk);
            if (lane < 2) {
                const int pt = 2 * warp + lane;
                const bool actp = pt < npt;
                const int64_t i = p0 + (int64_t)tl * DPT + pt;
                double dd[5];
#pragma unroll
                for (int r = 0; r < 5; ++r) dd[r] = lane == 0 ? d[0][r] : d[1][r];
                double* o = pw + pt * 16;
                if (actp) {
                    double cc[5], dz[5];
#pragma unroll
                    for (int k = 0; k < 5; ++k) { cc[k] = __ldg(&P.c[(int64_t)k * n + i]); dz[k] = P.Dz0 ? __ldg(&P.Dz0[(int64_t)k * n + i]) : 0.0; }
                    dz[0] += dd[3]; dz[1] += dd[0]; dz[2] += dd[1]; dz[3] += dd[2]; dz[4] += dd[4];
                    if (WDZ && P.Dz) {
#pragma unroll
                        for (int k = 0; k < 5; ++k) P.Dz[(int64_t)k * n + i] = dz[k];
                    }
                    const double wi = __ldg(&P.w[i]);
                    const double qv[3] = {dz[1], dz[2], dz[3]};
                    BarrierOut bo;
                    barrier_eval<3, WF, (WG || WH)>(qv, dz[4], P.p, bo);
                    double cd = 0.0;
#pragma unroll
                    for (int k = 0; k < 5; ++k) cd = fma(cc[k], dz[k], cd);
                    sc0 += WF ? wi * bo.F : 0.0; sc1 += wi * cd; sc2 += bo.feasible ? 0.0 : 1.0;
                    if (WG || WH) {
                        o[0] = wi * bo.Hqq[0][0]; o[1] = wi * bo.Hqq[0][1]; o[2] = wi * bo.Hqq[0][2];
                        o[3] = wi * bo.Hqq[1][1]; o[4] = wi * bo.Hqq[1][2]; o[5] = wi * bo.Hqq[2][2];
                        o[6] = wi * bo.Hqs[0]; o[7] = wi * bo.Hqs[1]; o[8] = wi * bo.Hqs[2]; o[9] = wi * bo.Hss;
                        o[10] = wi * (P.t * cc[0]);
                        o[11] = wi * (bo.gq[0] + P.t * cc[1]); o[12] = wi * (bo.gq[1] + P.t * cc[2]); o[13] = wi * (bo.gq[2] + P.t * cc[3]);
                        o[14] = wi * (bo.gs + P.t * cc[4]);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 15; ++k) o[k] = 0.0;
                }
            }
        }
        if (!(WG || WH)) continue;   // objective only: no contraction
        __syncthreads();
        // ---- phase 2a: per point T = (w F2_qq) A, bs = A' (w F2_qs), hs = (w F2_ss) I_s
        if (WH) {
            for (int idx = tid; idx < DPT * DN; idx += 256) {
                const int pt = idx / DN, b = idx % DN;
                const double* o = pw + pt * 16;
                const double* tp = tile + pt * (DNR * DS);
                const double A0 = tp[b], A1 = tp[DS + b], A2 = tp[2 * DS + b];
                Tt[(pt * 3 + 0) * DS + b] = o[0] * A0 + o[1] * A1 + o[2] * A2;
                Tt[(pt * 3 + 1) * DS + b] = o[1] * A0 + o[3] * A1 + o[4] * A2;
                Tt[(pt * 3 + 2) * DS + b] = o[2] * A0 + o[4] * A1 + o[5] * A2;
                bs[pt * DS + b] = o[6] * A0 + o[7] * A1 + o[8] * A2;
                hs[pt * DS + b] = o[9] * tp[4 * DS + b];
            }
            __syncthreads();
            // ---- phase 2b: the three contractions of the tile on the FP64 tensor cores.  K runs over (point, derivative)
            // for uu (48 = 12 steps of 4) and over the points for us / ss (16 = 4 steps); per step one A fragment and
            // eight B fragments (8-byte shared loads, conflict free by the row stride) feed eight DMMAs.
            const int arow = 8 * warp + fg;
#pragma unroll 4
            for (int kk = 0; kk < 3 * DPT / 4; ++kk) {
                const int k = 4 * kk + ft, pt = k / 3, j = k - 3 * pt;
                const double a = tile[(pt * DNR + j) * DS + arow];
                const double* brow = Tt + k * DS + fg;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) dmma884(uu[nt][0], uu[nt][1], a, brow[8 * nt]);
            }
#pragma unroll
            for (int kk = 0; kk < DPT / 4; ++kk) {
                const int pt = 4 * kk + ft;
                const double a1 = bs[pt * DS + arow], a2 = tile[(pt * DNR + 4) * DS + arow];
                const double* b1 = tile + (pt * DNR + 4) * DS + fg;
                const double* b2 = hs + pt * DS + fg;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    dmma884(us[nt][0], us[nt][1], a1, b1[8 * nt]);
                    dmma884(ss[nt][0], ss[nt][1], a2, b2[8 * nt]);
                }
            }
        }
        // ---- gradient on threads 0..127 (one unknown each)
        if (WG && tid < 2 * DN) {
            for (int pt = 0; pt < DPT; ++pt) {
                const double* tp = tile + pt * (DNR * DS);
                const double* o = pw + pt * 16;
                if (tid < DN) gacc += tp[tid] * o[11] + tp[DS + tid] * o[12] + tp[2 * DS + tid] * o[13] + tp[3 * DS + tid] * o[10];
                else gacc += tp[4 * DS + tid - DN] * o[14];
            }
        }
    }
    // ---- the chunk's records: lane holds C[g][2t], C[g][2t + 1] of every column tile
    if (WH) {
        double* rec = P.sel + ch * (int64_t)(3 * DN * DN) + (8 * warp + fg) * DN + 2 * ft;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            *reinterpret_cast<double2*>(rec + 8 * nt) = make_double2(uu[nt][0], uu[nt][1]);
            *reinterpret_cast<double2*>(rec + DN * DN + 8 * nt) = make_double2(us[nt][0], us[nt][1]);
            *reinterpret_cast<double2*>(rec + 2 * DN * DN + 8 * nt) = make_double2(ss[nt][0], ss[nt][1]);
        }
    }
    if (WG && tid < 2 * DN) P.rel[ch * (2 * DN) + tid] = gacc;
    block_scalars(sc0, sc1, sc2, P.part);
}

}  // namespace mgb
