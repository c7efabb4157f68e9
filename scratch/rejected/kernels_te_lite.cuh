// Thread-per-element stage 1 of an assembly on fine levels (one cone; fem1d / fem2d finest level and every level
// whose id-like operators own one column per row).  Opt-in (MGB_TE=1) variant of element_kernel.
//
// element_kernel (kernels.cuh) gives every quadrature point a lane and pays for it in cross-lane traffic: ~1,000 warp
// instructions per 4 elements, a fifth of them arithmetic.  Here ONE THREAD owns an element: it walks the element's
// B points, keeps the B(B+1)/2 unique u-u block entries and the u-gradient in registers, and needs no shuffle, no
// select tree and no reduction at all.  Shared memory per warp is kept near 31 KB so that seven warps - one tile of
// 32 elements each - are resident per SM and an L=8 level runs in a single round:
//   * operator records are stored per warp tile as [point][16-byte chunk][lane], so the records of one point of all
//     32 elements are one contiguous slab: a 1-D bulk copy (TMA, SASS UBLKCP) brings slab g + 1 into a two-stage
//     shared-memory ring behind an mbarrier while slab g is consumed (conflict-free 16-byte loads);
//   * the next tile's unknowns (gathered through the dof ids) and its c / Dz0 rows are staged by cp.async (LDGSTS)
//     as soon as the current tile has read its last point;
//   * results leave as 2 KB slabs too: per point the s-row [su(q',.) | ss(q')] of the 32 elements, per local u-row
//     [uu(q,.)], written to shared memory with a 16-byte-chunk XOR swizzle (conflict free) and sent to HBM by one
//     bulk store each.  The u-s entries of a u-row are read by the gather kernel from the s-row cells (column access
//     into the element's seven adjacent 64-byte cells), so no transpose scratch is needed.
// Record layout in `sel`: [tile][u-row q | s-row of point l][lane][8 doubles] (abs_slot in plan_host.cpp bakes the
// swizzle into the frozen index lists).
#pragma once
#include "kernels.cuh"

namespace mgb {

// ---- 1-D bulk copy (TMA) + mbarrier helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) { while (!mbar_try_wait(bar, parity)) {} }
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// swizzled position of 16-byte chunk c inside a lane's cell of NC chunks (NC = 1, 2, 4, 8): quarter-warp 16-byte
// stores of the same chunk index land on distinct banks.  Shared by host (slot indices) and device.
__host__ __device__ __forceinline__ constexpr int te_swz(int lane, int c, int NC) { return NC >= 8 ? (c ^ (lane & 7)) : NC == 4 ? (c ^ ((lane >> 1) & 3)) : NC == 2 ? (c ^ ((lane >> 2) & 1)) : c; }

template <int B, int D>
struct TeShape {
    static constexpr int NU = 2, ND = D + 2;
    static constexpr int RW = (D * B + 1 + NU + 1 + 1) / 2 * 2, CH = RW / 2;   // per-point record (fine): doubles / 16-byte chunks
    static constexpr int SLAB_IN = CH * 32 * 16;                                // bytes: one point of 32 elements
    static constexpr int RC = (B + 1 + 1) / 2 * 2 <= 2 ? 2 : ((B + 1) <= 4 ? 4 : 8);   // doubles per output cell: B + 1 values
    static_assert(B + 1 <= 8, "cell holds B values and one diagonal entry");
    static constexpr int NC = RC / 2;                                           // 16-byte chunks per cell (1, 2 or 4)
    static constexpr int SLAB_OUT = 32 * RC * 8;                                // bytes: one row of 32 elements
    static constexpr int TS = 2 * B * 32 * RC;                                  // doubles of slot records per tile
    static constexpr int NSTG = 2;
    // shared memory (bytes), one warp per CTA
    static constexpr int OFF_Z = NSTG * SLAB_IN;                 // [NU][B][32] unknowns of the tile
    static constexpr int OFF_ROWS = OFF_Z + NU * B * 32 * 8;      // [c | Dz0][ND][B][32]
    static constexpr int OFF_OUT = OFF_ROWS + 2 * ND * B * 32 * 8;   // [2][32][RC] output slabs
    static constexpr int OFF_BARS = OFF_OUT + 2 * SLAB_OUT;
    static constexpr int SMEM = OFF_BARS + 64;
};

__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

template <int B, int D, int FLAGS>
__global__ void __launch_bounds__(32, 7) element_te_kernel(const __grid_constant__ ElemParams P) {
    using S = TeShape<B, D>;
    constexpr int NU = S::NU, ND = S::ND, RW = S::RW, CH = S::CH, NSTG = S::NSTG, RC = S::RC, NC = S::NC;
    constexpr bool WF = (FLAGS & 1) != 0, WG = (FLAGS & 2) != 0, WH = (FLAGS & 4) != 0, WDZ = (FLAGS & 8) != 0;
    constexpr int NTRI = B * (B + 1) / 2;
    extern __shared__ __align__(128) unsigned char sm[];
    pdl_launch_dependents();
    const int lane = threadIdx.x;
    double* zbuf = reinterpret_cast<double*>(sm + S::OFF_Z);
    double* rows = reinterpret_cast<double*>(sm + S::OFF_ROWS);
    double* oslab = reinterpret_cast<double*>(sm + S::OFF_OUT);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S::OFF_BARS);
    const int64_t tile0 = blockIdx.x, GW = gridDim.x, n = P.nloc;
    const int64_t nt = tile0 < P.ntiles ? (P.ntiles - tile0 + GW - 1) / GW : 0;
    const int64_t G = nt * B;   // slabs this warp consumes

    if (lane == 0) {
#pragma unroll
        for (int sidx = 0; sidx < NSTG; ++sidx) mbar_init(bars + sidx, 1);
        mbar_init_fence();
    }
    if (WH) for (int k = lane; k < 2 * 32 * RC; k += 32) oslab[k] = 0.0;
    __syncwarp();

    auto issue_slab = [&](int64_t g) {   // slab g of this warp's sequence -> stage g % NSTG
        if (lane == 0) {
            const int64_t tile = tile0 + (g / B) * GW;
            const int stage = (int)(g % NSTG);
            mbar_expect_tx(bars + stage, (unsigned)S::SLAB_IN);
            bulk_g2s(sm + stage * S::SLAB_IN, reinterpret_cast<const unsigned char*>(P.prec) + (tile * B + g % B) * (int64_t)S::SLAB_IN,
                     (unsigned)S::SLAB_IN, bars + stage);
        }
    };
    int32_t ids[NU * B];   // dof ids of this lane's element in the tile that is staged next
    auto load_ids = [&](int64_t tile) {
#pragma unroll
        for (int a = 0; a < NU * B; ++a) ids[a] = (tile < P.ntiles) ? __ldg(&P.lcols[(tile * NU * B + a) * 32 + lane]) : -1;
    };
    auto stage_rows = [&](int64_t tile) {   // unknowns (through ids) + c / Dz0 rows of `tile` -> shared memory
#pragma unroll
        for (int a = 0; a < NU * B; ++a) {
            if (ids[a] >= 0) cp_async_8(zbuf + a * 32 + lane, P.s + ids[a]);
            else zbuf[a * 32 + lane] = 0.0;
        }
        const int64_t row0 = tile * (32 * B);
#pragma unroll
        for (int r = 0; r < B; ++r) {   // coalesced: lane copies points lane, lane + 32, ... of the tile's 32 B rows
            const int idx = r * 32 + lane, el = idx / B, l = idx % B;
            const bool ok = row0 + idx < n;
#pragma unroll
            for (int k = 0; k < ND; ++k) {
                double* dc = rows + ((k * B + l) * 32 + el);
                double* dd = rows + (((ND + k) * B + l) * 32 + el);
                if (ok) cp_async_8(dc, P.c + (int64_t)k * n + row0 + idx); else *dc = 0.0;
                if (ok && P.Dz0) cp_async_8(dd, P.Dz0 + (int64_t)k * n + row0 + idx); else *dd = 0.0;
            }
        }
        cp_async_commit();
    };

    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    if (nt > 0) {
        issue_slab(0);
        load_ids(tile0);
        stage_rows(tile0);
        load_ids(tile0 + GW);
    }
    int64_t nstore = 0;   // bulk stores issued (two output slab buffers, alternating)
    for (int64_t it = 0; it < nt; ++it) {
        const int64_t tile = tile0 + it * GW;
        const bool more = it + 1 < nt;
        cp_async_wait<0>();
        __syncwarp();
        const int64_t e = tile * 32 + lane;
        const bool act = e < P.E;
        double z[NU][B];
#pragma unroll
        for (int v = 0; v < NU; ++v)
#pragma unroll
            for (int q = 0; q < B; ++q) z[v][q] = zbuf[(v * B + q) * 32 + lane];
        double uu[NTRI], ru[B];
#pragma unroll
        for (int r = 0; r < NTRI; ++r) uu[r] = 0.0;
#pragma unroll
        for (int q = 0; q < B; ++q) ru[q] = 0.0;

        // NOT unrolled: one copy of the point body keeps the kernel inside the instruction cache
#pragma unroll 1
        for (int l = 0; l < B; ++l) {
            const int64_t g = it * B + l;
            __syncwarp();                        // every lane is done with the stage that slab g + 1 overwrites
            if (g + 1 < G) issue_slab(g + 1);
            const int stage = (int)(g % NSTG);
            mbar_wait(bars + stage, (unsigned)((g / NSTG) & 1));
            double rec[RW];
            {
                const double2* rp = reinterpret_cast<const double2*>(sm + stage * S::SLAB_IN) + lane;
#pragma unroll
                for (int j = 0; j < CH; ++j) { const double2 t2 = rp[j * 32]; rec[2 * j] = t2.x; rec[2 * j + 1] = t2.y; }
            }
            const double wi = act ? rec[D * B] : 0.0;
            const unsigned long long lqbits = (unsigned long long)__double_as_longlong(rec[D * B + 1 + NU]);
            const int olq0 = act ? (int)(lqbits & 0xFFull) : 255, olq1 = act ? (int)((lqbits >> 8) & 0xFFull) : 255;
            const bool oh0 = olq0 != 255, oh1 = olq1 != 255;
            const double oval0 = oh0 ? rec[D * B + 1] : 0.0, oval1 = oh1 ? rec[D * B + 2] : 0.0;
            double cc[ND], dz[ND];
#pragma unroll
            for (int k = 0; k < ND; ++k) { cc[k] = rows[(k * B + l) * 32 + lane]; dz[k] = rows[((ND + k) * B + l) * 32 + lane]; }
            if (l == B - 1 && more) {
                // the tile's last reads of the staging buffers are done: stage the next tile behind the rest of this one
                __syncwarp();
                stage_rows(tile + GW);
                load_ids(tile + 2 * GW);
            }
            // ---- apply_D
            double zo0 = z[0][0], zo1 = z[1][0];
#pragma unroll
            for (int q = 1; q < B; ++q) { zo0 = (olq0 == q) ? z[0][q] : zo0; zo1 = (olq1 == q) ? z[1][q] : zo1; }
            dz[0] = fma(oval0, zo0, dz[0]);
            dz[D + 1] = fma(oval1, zo1, dz[D + 1]);
#pragma unroll
            for (int k = 0; k < D; ++k)
#pragma unroll
                for (int q = 0; q < B; ++q) dz[1 + k] = fma(rec[k * B + q], z[0][q], dz[1 + k]);
            const int64_t i = e * B + l;
            if (WDZ && act && P.Dz) {
#pragma unroll
                for (int k = 0; k < ND; ++k) P.Dz[(int64_t)k * n + i] = dz[k];
            }
            // ---- barrier
            double qv[D];
#pragma unroll
            for (int j = 0; j < D; ++j) qv[j] = act ? dz[1 + j] : 0.0;
            const double sv = act ? dz[D + 1] : 1.0;
            BarrierOut bo;
            barrier_eval<D, WF, (WG || WH)>(qv, sv, P.p, bo);
            {
                double cd = 0.0;
#pragma unroll
                for (int k = 0; k < ND; ++k) cd = fma(cc[k], dz[k], cd);
                const bool act_s = act && e < P.Eprim;
                acc0 += (WF && act_s) ? wi * bo.F : 0.0;
                acc1 += act_s ? wi * cd : 0.0;
                acc2 += (act_s && !bo.feasible) ? 1.0 : 0.0;
            }
            // ---- gradient
            if (WG) {
                const double gy0 = wi * (P.t * cc[0]);
                double gyd[D];
#pragma unroll
                for (int j = 0; j < D; ++j) gyd[j] = wi * (bo.gq[j] + P.t * cc[1 + j]);
                const double gys = wi * (bo.gs + P.t * cc[D + 1]);
#pragma unroll
                for (int q = 0; q < B; ++q) {
#pragma unroll
                    for (int k = 0; k < D; ++k) ru[q] = fma(rec[k * B + q], gyd[k], ru[q]);
                    ru[q] += (q == olq0) ? oval0 * gy0 : 0.0;
                }
                if (oh1) P.rel[((tile * NU + 1) * B + olq1) * 32 + lane] = oval1 * gys;
            }
            // ---- Hessian
            if (WH) {
                double T[D][B];
#pragma unroll
                for (int j = 0; j < D; ++j)
#pragma unroll
                    for (int q = 0; q < B; ++q) {
                        double tacc = 0.0;
#pragma unroll
                        for (int j2 = 0; j2 < D; ++j2) tacc = fma(wi * bo.Hqq[j][j2], rec[j2 * B + q], tacc);
                        T[j][q] = tacc;
                    }
#pragma unroll
                for (int q = 0; q < B; ++q)
#pragma unroll
                    for (int q2 = q; q2 < B; ++q2) {
                        double a2 = uu[q * B - q * (q - 1) / 2 + (q2 - q)];
#pragma unroll
                        for (int j = 0; j < D; ++j) a2 = fma(rec[j * B + q], T[j][q2], a2);
                        uu[q * B - q * (q - 1) / 2 + (q2 - q)] = a2;
                    }
                double val[RC];   // su(olq1, 0..B-1), ss(olq1), padding
#pragma unroll
                for (int q = 0; q < RC; ++q) val[q] = 0.0;
#pragma unroll
                for (int q = 0; q < B; ++q) {
                    double bsq = 0.0;
#pragma unroll
                    for (int j = 0; j < D; ++j) bsq = fma(rec[j * B + q], wi * bo.Hqs[j], bsq);
                    val[q] = bsq * oval1;
                }
                val[B] = wi * bo.Hss * oval1 * oval1;
                // s-row slab of this point: [su | ss] of 32 elements, swizzled 16-byte chunks, one bulk store
                if (lane == 0) bulk_wait_read1();
                __syncwarp();
                double2* sc = reinterpret_cast<double2*>(oslab + (nstore & 1) * (32 * RC) + lane * RC);
#pragma unroll
                for (int c2 = 0; c2 < NC; ++c2) sc[te_swz(lane, c2, NC)] = make_double2(val[2 * c2], val[2 * c2 + 1]);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) bulk_s2g(P.sel + tile * (int64_t)S::TS + (B + l) * (32 * RC), oslab + (nstore & 1) * (32 * RC), S::SLAB_OUT);
                ++nstore;
            }
        }
        // ---- gradient record of the u dofs, u-row slabs [uu(q,.)]
        if (WG) {
#pragma unroll
            for (int q = 0; q < B; ++q) P.rel[((tile * NU + 0) * B + q) * 32 + lane] = ru[q];
        }
        if (WH) {
#pragma unroll
            for (int q = 0; q < B; ++q) {
                double row[RC];
#pragma unroll
                for (int c = 0; c < RC; ++c) row[c] = 0.0;
#pragma unroll
                for (int q2 = 0; q2 < B; ++q2) {
                    const int lo = q < q2 ? q : q2, hi = q < q2 ? q2 : q;
                    row[q2] = uu[lo * B - lo * (lo - 1) / 2 + (hi - lo)];
                }
                if (lane == 0) bulk_wait_read1();
                __syncwarp();
                double2* uc = reinterpret_cast<double2*>(oslab + (nstore & 1) * (32 * RC) + lane * RC);
#pragma unroll
                for (int c2 = 0; c2 < NC; ++c2) uc[te_swz(lane, c2, NC)] = make_double2(row[2 * c2], row[2 * c2 + 1]);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) bulk_s2g(P.sel + tile * (int64_t)S::TS + q * (32 * RC), oslab + (nstore & 1) * (32 * RC), S::SLAB_OUT);
                ++nstore;
            }
        }
    }
    if (WH && lane == 0) bulk_wait_all();
    block_scalars(acc0, acc1, acc2, P.part);
}

}  // namespace mgb
