// Stage 1 of an assembly on the DENSE path (plan_host.h ElementPlan::dense): coarse levels of large elements (fem3d
// Q3 hexahedra), where every fine quadrature point touches up to 64 unknowns per variable and a product-list replay
// would need ~1 G products per level.  Per chunk of <= 512 points of one group (one CTA of 256 threads):
//   * the group's 2 x 64 unknowns are gathered once;
//   * the points' dense operator rows ([point][dx dy dz u.id s.id][64 dofs], 2.5 KB per point) stream in tiles of 16
//     points through a two-stage shared-memory ring filled by 1-D bulk copies (TMA: cp.async.bulk + mbarrier);
//   * phase 1 (two points per warp): apply_D as 64-long dot products (lane = 2 dofs, shuffle reduction), the barrier
//     at the two points on lanes 0 and 1 side by side, w .* F1 / F2 into shared memory, objective partials;
//   * phase 2: the three contractions  uu += (w F2_qq A)' A,  us += (A' w F2_qs) I_s,  ss += (w F2_ss I_s)' I_s  on the
//     FP64 tensor cores (mma.sync m8n8k4 = SASS DMMA.8x8x4): warp w owns rows [8w, 8w+8) of each 64 x 64 block (8
//     column tiles, 48 accumulators per lane), the barrier's small blocks are applied while the A fragments are formed,
//     the B fragments are the operator rows as they lie in the tile; the one dense contraction of this path, and the
//     one place where tensor cores pay.  The gradient rides along (one unknown per thread, two halves of the points);
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
constexpr int DPW = 20;        // doubles per point of the barrier record: w F2_qq (3 x 3), w F2_qs (3), w F2_ss, w (F1 + t c) (5), pad
constexpr int DST = 4;         // ring stages: the MMA warps on tile t, the two producer groups on t+1 and t+2, TMA filling t+3
constexpr int D_MMA_WARPS = 8, D_PRO_WARPS = 4;
constexpr int D_THREADS = 32 * (D_MMA_WARPS + D_PRO_WARPS);   // 384
constexpr int D_SMEM = DST * D_TILE_BYTES          // operator-row ring
                     + 2 * DN * 8                  // unknowns
                     + DST * DPT * DPW * 8         // barrier records of the tiles' points
                     + 128;                        // mbarriers

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

__device__ __forceinline__ void d_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(d_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void d_mbar_wait(uint64_t* bar, unsigned parity) {
    while (!d_mbar_try_wait(bar, parity)) {}
}

// FLAGS bits as in the element kernels: 1 objective, 2 gradient, 4 Hessian, 8 store Dz.
// Warp-specialised: warps 0..7 run the tensor-core contraction (+ the gradient) of tile t while warps 8..11
// ("producers") evaluate apply_D + barrier ahead of them - two groups of two warps taking alternate tiles, because the
// barrier is a ~400-long dependent FP64 chain that shares the FP64 pipe with the DMMAs (ncu: `stall_math` on every
// producer instruction): one group alone took longer per tile than the contraction, two may each take two tile periods.
// Three mbarriers per ring stage: full (TMA bytes landed), ready (the group's barrier records are written), empty
// (everybody is done with the stage; thread 0 then refills it four tiles ahead).
template <int FLAGS>
__global__ void __launch_bounds__(D_THREADS, 1) dense_element_kernel(const __grid_constant__ DenseParams P) {
    constexpr bool WF = (FLAGS & 1) != 0, WG = (FLAGS & 2) != 0, WH = (FLAGS & 4) != 0, WDZ = (FLAGS & 8) != 0;
    extern __shared__ __align__(128) unsigned char dsm[];
    double* tile0 = reinterpret_cast<double*>(dsm);
    double* zu = reinterpret_cast<double*>(dsm + DST * D_TILE_BYTES);
    double* zs = zu + DN;
    double* pw0 = zs + DN;                 // [DST][DPT][DPW]
    uint64_t* bars = reinterpret_cast<uint64_t*>(pw0 + DST * DPT * DPW);
    uint64_t* full = bars, *ready = bars + DST, *empty = bars + 2 * DST;
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
        for (int st = 0; st < DST; ++st) {
            d_mbar_init(full + st, 1);
            d_mbar_init(ready + st, 64);
            d_mbar_init(empty + st, 32 * D_MMA_WARPS + 64);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    double sc0 = 0.0, sc1 = 0.0, sc2 = 0.0;
    auto issue = [&](int tl) {   // tile tl -> stage tl % DST (thread 0)
        const int64_t q0 = p0 + (int64_t)tl * DPT;
        const unsigned bytes = (unsigned)(min((int64_t)DPT, p1 - q0) * DNR * DS * 8);
        d_mbar_expect_tx(full + tl % DST, bytes);
        d_bulk_g2s(tile0 + (size_t)(tl % DST) * (DPT * DNR * DS), P.rows + q0 * (DNR * DS), bytes, full + tl % DST);
    };
    if (tid == 0)
        for (int tl = 0; tl < DST && tl < ntile; ++tl) issue(tl);

    if (warp >= D_MMA_WARPS) {
        // ================================================================ producers (2 groups x 64 threads)
        const int pwarp = warp - D_MMA_WARPS, grp = pwarp >> 1, wg = pwarp & 1, gtid = tid & 63;
        // eight points per warp; after the transposing reduction lane 4q holds the five dot products of point 8 wg + q.
        // Per-point scalars (c, Dz0, w) of the evaluating lanes are fetched ONE TILE AHEAD: loaded on demand they put a
        // full DRAM latency into every tile's critical path (the barrier chain starts from them)
        const bool ev = (lane & 3) == 0;
        const int ptl = 8 * wg + (lane >> 2);
        double ncc[5], ndz[5], nwi = 0.0;
        auto prefetch = [&](int tl) {
            const int64_t i = p0 + (int64_t)tl * DPT + ptl;
            const bool on = ev && i < p1;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                ncc[k] = on ? __ldg(&P.c[(int64_t)k * n + i]) : 0.0;
                ndz[k] = (on && P.Dz0) ? __ldg(&P.Dz0[(int64_t)k * n + i]) : 0.0;
            }
            nwi = on ? __ldg(&P.w[i]) : 0.0;
        };
        if (grp < ntile) prefetch(grp);
        const double zu0 = zu[lane], zu1 = zu[lane + 32], zs0 = zs[lane], zs1 = zs[lane + 32];
        for (int tl = grp; tl < ntile; tl += 2) {
            const int st = tl % DST;
            double cc[5], dz[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) { cc[k] = ncc[k]; dz[k] = ndz[k]; }
            const double wi = nwi;
            if (tl + 2 < ntile) prefetch(tl + 2);
            double* tile = tile0 + (size_t)st * (DPT * DNR * DS);
            double* pw = pw0 + st * (DPT * DPW);
            const int npt = (int)min((int64_t)DPT, npts - (int64_t)tl * DPT);
            d_mbar_wait(full + st, (unsigned)((tl / DST) & 1));
            if (npt < DPT)   // the tail of a partial tile was not loaded: it must not contribute (0 * stale NaN)
                for (int k = npt * DNR * DS + gtid; k < DPT * DNR * DS; k += 64) tile[k] = 0.0;
            // ---- apply_D as 64-long dot products: lane = 2 dofs, 8 points x 5 rows = 40 partial sums per lane, reduced by
            // a transposing butterfly (xor 16 halves the value set - done while the rows are read, points q and q + 4
            // together -, xor 8 and xor 4 halve it again, then xor 2 / 1 on the five that are left)
            {
                auto dot5 = [&](int pp, double* out) {
                    const int pt = 8 * wg + pp;
                    const bool actp = pt < npt;
                    const double* tp = tile + pt * (DNR * DS);
#pragma unroll
                    for (int r = 0; r < 4; ++r) out[r] = actp ? tp[r * DS + lane] * zu0 + tp[r * DS + lane + 32] * zu1 : 0.0;
                    out[4] = actp ? tp[4 * DS + lane] * zs0 + tp[4 * DS + lane + 32] * zs1 : 0.0;
                };
                const bool up1 = (lane & 16) != 0, up2 = (lane & 8) != 0, up3 = (lane & 4) != 0;
                double e[20], f[10], dd[5];
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                    double lo[5], hi[5];
                    dot5(pp, lo); dot5(pp + 4, hi);
#pragma unroll
                    for (int r = 0; r < 5; ++r) e[pp * 5 + r] = (up1 ? hi[r] : lo[r]) + shfl_xor_d(up1 ? lo[r] : hi[r], 16);
                }
#pragma unroll
                for (int i = 0; i < 10; ++i) f[i] = (up2 ? e[i + 10] : e[i]) + shfl_xor_d(up2 ? e[i] : e[i + 10], 8);
#pragma unroll
                for (int i = 0; i < 5; ++i) dd[i] = (up3 ? f[i + 5] : f[i]) + shfl_xor_d(up3 ? f[i] : f[i + 5], 4);
#pragma unroll
                for (int mk = 2; mk >= 1; mk >>= 1)
#pragma unroll
                    for (int i = 0; i < 5; ++i) dd[i] += shfl_xor_d(dd[i], mk);
                if (ev) {
                    const bool actp = ptl < npt;
                    const int64_t i = p0 + (int64_t)tl * DPT + ptl;
                    double* o = pw + ptl * DPW;
                    if (actp) {
                        dz[0] += dd[3]; dz[1] += dd[0]; dz[2] += dd[1]; dz[3] += dd[2]; dz[4] += dd[4];
                        if (WDZ && P.Dz) {
#pragma unroll
                            for (int k = 0; k < 5; ++k) P.Dz[(int64_t)k * n + i] = dz[k];
                        }
                        const double qv[3] = {dz[1], dz[2], dz[3]};
                        BarrierOut bo;
                        barrier_eval<3, WF, (WG || WH)>(qv, dz[4], P.p, bo);
                        double cd = 0.0;
#pragma unroll
                        for (int k = 0; k < 5; ++k) cd = fma(cc[k], dz[k], cd);
                        sc0 += WF ? wi * bo.F : 0.0; sc1 += wi * cd; sc2 += bo.feasible ? 0.0 : 1.0;
                        if (WG || WH) {
#pragma unroll
                            for (int j = 0; j < 3; ++j) {
#pragma unroll
                                for (int j2 = 0; j2 < 3; ++j2) o[3 * j + j2] = wi * bo.Hqq[j < j2 ? j : j2][j < j2 ? j2 : j];
                                o[9 + j] = wi * bo.Hqs[j];
                            }
                            o[12] = wi * bo.Hss;
                            o[13] = wi * (P.t * cc[0]);
                            o[14] = wi * (bo.gq[0] + P.t * cc[1]); o[15] = wi * (bo.gq[1] + P.t * cc[2]); o[16] = wi * (bo.gq[2] + P.t * cc[3]);
                            o[17] = wi * (bo.gs + P.t * cc[4]);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 18; ++k) o[k] = 0.0;
                    }
                }
            }
            d_mbar_arrive(ready + st);            // this thread's part of the tile's barrier records (and zero fill) is written
            d_mbar_arrive(empty + st);
        }
    } else {
        // ================================================================ tensor-core warps (256 threads)
        // warp w owns output rows [8w, 8w + 8) of the three 64 x 64 blocks: 8 column tiles of 8, two accumulators per lane each
        const int fg = lane >> 2, ft = lane & 3;
        double uu[8][2], us[8][2], ss[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) uu[i][0] = uu[i][1] = us[i][0] = us[i][1] = ss[i][0] = ss[i][1] = 0.0;
        const int arow = 8 * warp + fg;
        double gu = 0.0, gs = 0.0;   // gradient partials of unknowns arow (u) and arow (s) over this lane's points
        for (int tl = 0; tl < ntile; ++tl) {
            const int st = tl % DST;
            const double* tile = tile0 + (size_t)st * (DPT * DNR * DS);
            const double* pw = pw0 + st * (DPT * DPW);
            d_mbar_wait(ready + st, (unsigned)((tl / DST) & 1));   // (also without a Hessian: nobody may run a phase ahead)
            if (WH) {
                // ---- the three contractions of the tile,  uu += (F2_qq A)' A,  us += (A' F2_qs) I_s,  ss += (F2_ss I_s)' I_s.
                // K runs over (point, derivative) for uu and over the points for us / ss; a k-step of 4 takes the SAME
                // derivative of four consecutive points (lane t of a quad: point 4q + t), so a lane keeps its point for
                // the three uu steps and the us / ss step of a quad: the three derivative rows of its output column are
                // loaded once, the barrier's 3 x 3 / 3 x 1 / 1 x 1 blocks are applied to them in registers (no scaled copy
                // of the tile is ever written), and the rows a quad of lanes reads lie 5 * 68 doubles apart - distinct
                // 8-byte bank pairs for A and B fragments alike.  The B fragments are the operator rows themselves;
                // us / ss share theirs.
#ifndef MGB_DENSE_SKIP_MMA   // (timing experiments only: wrong results)
#pragma unroll 2
                for (int q = 0; q < DPT / 4; ++q) {
                    const int pt = 4 * q + ft;
                    const double* tp = tile + pt * (DNR * DS);
                    const double2* o2 = reinterpret_cast<const double2*>(pw + pt * DPW);
                    const double t0 = tp[arow], t1 = tp[DS + arow], t2 = tp[2 * DS + arow], t3 = tp[3 * DS + arow], t4 = tp[4 * DS + arow];
                    double o[18];
#pragma unroll
                    for (int i = 0; i < 9; ++i) { const double2 v = o2[i]; o[2 * i] = v.x; o[2 * i + 1] = v.y; }
                    // the gradient rides along on the operands already in registers (the FP64 pipe is the tensor pipe:
                    // a separate pass over the tile by other warps would compete with the DMMAs for it)
                    gu += t0 * o[14] + t1 * o[15] + t2 * o[16] + t3 * o[13];
                    gs += t4 * o[17];
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const double a = o[3 * j] * t0 + o[3 * j + 1] * t1 + o[3 * j + 2] * t2;
                        const double* brow = tp + j * DS + fg;
#pragma unroll
                        for (int nt = 0; nt < 8; ++nt) dmma884(uu[nt][0], uu[nt][1], a, brow[8 * nt]);
                    }
                    const double a1 = o[9] * t0 + o[10] * t1 + o[11] * t2;
                    const double a2 = o[12] * t4;
                    const double* brow = tp + 4 * DS + fg;
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) {
                        const double b = brow[8 * nt];
                        dmma884(us[nt][0], us[nt][1], a1, b);
                        dmma884(ss[nt][0], ss[nt][1], a2, b);
                    }
                }
#endif
            }
            d_mbar_arrive(empty + st);
            if (tid == 0 && tl + DST < ntile) {   // refill this stage once the other warps have left it too
                d_mbar_wait(empty + st, (unsigned)((tl / DST) & 1));
                issue(tl + DST);
            }
        }
        if (WG) {   // the four lanes of a quad hold the partials of the same unknowns over different points
            gu += shfl_xor_d(gu, 1); gu += shfl_xor_d(gu, 2);
            gs += shfl_xor_d(gs, 1); gs += shfl_xor_d(gs, 2);
            if (ft == 0) { P.rel[ch * (2 * DN) + arow] = gu; P.rel[ch * (2 * DN) + DN + arow] = gs; }
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
    }
    block_scalars(sc0, sc1, sc2, P.part);
}

}  // namespace mgb
