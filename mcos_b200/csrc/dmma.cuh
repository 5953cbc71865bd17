// FP64 tensor-core (DMMA) tile machinery shared by the moments SYRK and the sampling TRMM.
//
// tcgen05 has no f64 kind, so the FP64 tensor path on sm_100a is the warp-level mma.sync m8n8k4 (SASS DMMA.8x8x4;
// measured 37.0 TFLOP/s on B200, profiles/fp64_peaks_r01.json).  A warp owns 64 x 32 of the FP64 accumulator tile as
// 8 x 4 m8n8 blocks = 64 doubles per thread.  The kernels built on this (trmm2, syrk, hrp_gram) run CTAs of 4 warps
// (2 x 2: a 128 x 64 tile, or one of the two warp sets of a diagonal 128 x 128 tile), two CTAs per SM, so that the
// prologue / epilogue of one CTA overlaps the main loop of the other; the original 8-warp / 128 x 128 TRMM is kept for
// A/B runs.  Operand tiles move global -> shared with cp.async (LDGSTS, 16-byte when the alignment allows, zero-fill
// for the ragged edges) through a DM_STAGES-deep ring, one __syncthreads per k-slab; the copies of stage kt + 2 are
// issued after the first k-step of slab kt so that they interleave with DMMAs already queued.
#pragma once
#include <type_traits>

#define DM_BM 128
#define DM_BN 128
#ifndef DM_BK
#define DM_BK 16
#endif
#ifndef DM_STAGES
#define DM_STAGES 3
#endif
#define DM_THREADS 256
#define DM_WM 64   // warp tile rows
#define DM_WN 32   // warp tile cols
#define DM_MB (DM_WM / 8)
#define DM_NB (DM_WN / 8)
// shared-memory leading dimensions (doubles); both keep a half-warp's 16 fragment loads on 16 distinct bank pairs
#define DM_LD_KMAJOR (DM_BM + 4)  // tile stored [k][m]
#define DM_LD_MMAJOR (DM_BK + 4)  // tile stored [m][k]
#define DM_TILE_KMAJOR (DM_BK * DM_LD_KMAJOR)
#define DM_TILE_MMAJOR (DM_BM * DM_LD_MMAJOR)

__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// out of place: (d0, d1) = A B + (c0, c1); the C registers are only read
__device__ __forceinline__ void dmma_884_oop(double& d0, double& d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1)
                 : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// 8-byte asynchronous copy global -> shared; copies nothing and zero-fills when !pred (src must still be a valid address)
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, bool pred) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(gsrc), "r"(sz));
}
// 16-byte variant (LDGSTS.128): both addresses must be 16-byte aligned
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc, bool pred) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// Lower-triangle tile p -> (I, J), I >= J
__device__ __forceinline__ void dm_tile_ij(int p, int& I, int& J) {
    int i = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= p) ++i;
    while (i * (i + 1) / 2 > p) --i;
    I = i;
    J = p - i * (i + 1) / 2;
}

// One BK-deep slab of MMAs for this warp.  KMAJOR: tile element (mn, k) lives at tile[k * ld + mn]; otherwise at
// tile[mn * ld + k].  acc[mb][nb][2] follows the m8n8k4 C layout: row = lane/4, cols = 2*(lane%4) + {0,1}.
// CENTER: subtract per-column offsets (amean[mb], bmean[nb]) from the fragments as they are loaded -- the SYRK's
// centring x - mean, done in registers so the tiles can be copied raw; only the first kvalid k-rows of the slab count
// (rows beyond hold zero fill and must contribute nothing).
template <bool A_KMAJOR, bool B_KMAJOR, bool CENTER, bool RAGGED>
__device__ __forceinline__ void dmma_kstep(const double* __restrict__ As, const double* __restrict__ Bs, int wm0,
                                           int wn0, int lane, int k0, double (&acc)[DM_MB][DM_NB][2],
                                           const double* amean, const double* bmean, int kvalid) {
    const int g = lane >> 2, t = lane & 3;
    double a[DM_MB], b[DM_NB];
    const bool live = !RAGGED || (k0 + t < kvalid);
#pragma unroll
    for (int mb = 0; mb < DM_MB; ++mb) {
        int m = wm0 + mb * 8 + g;
#if defined(DM_ABLATE) && (DM_ABLATE & 1)   // timing experiment only: no shared-memory operand loads
        a[mb] = __hiloint2double(0x3ff00000 + m, k0);
#else
        a[mb] = A_KMAJOR ? As[(k0 + t) * DM_LD_KMAJOR + m] : As[m * DM_LD_MMAJOR + k0 + t];
#endif
#if !(defined(DM_ABLATE) && (DM_ABLATE & 2))   // timing experiment only: no centring
        if (CENTER) a[mb] = a[mb] - amean[mb];
#endif
        if (RAGGED) a[mb] = live ? a[mb] : 0.0;
    }
#pragma unroll
    for (int nb = 0; nb < DM_NB; ++nb) {
        int n = wn0 + nb * 8 + g;
#if defined(DM_ABLATE) && (DM_ABLATE & 1)
        b[nb] = __hiloint2double(0x3ff00000 + n, k0);
#else
        b[nb] = B_KMAJOR ? Bs[(k0 + t) * DM_LD_KMAJOR + n] : Bs[n * DM_LD_MMAJOR + k0 + t];
#endif
#if !(defined(DM_ABLATE) && (DM_ABLATE & 2))
        if (CENTER) b[nb] = b[nb] - bmean[nb];
#endif
        if (RAGGED) b[nb] = live ? b[nb] : 0.0;
    }
#pragma unroll
    for (int mb = 0; mb < DM_MB; ++mb)
#pragma unroll
        for (int nb = 0; nb < DM_NB; ++nb) dmma_884(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
}

// `mid()` runs after the first k-step: the cp.async issue of a later stage goes there, so that its address arithmetic
// and LDGSTS instructions interleave with DMMAs already queued on the tensor pipe instead of idling it after the
// barrier.  Slabs with kvalid >= DM_BK take the path without the per-fragment selects.
template <bool A_KMAJOR, bool B_KMAJOR, bool CENTER, class Mid>
__device__ __forceinline__ void dmma_slab(const double* __restrict__ As, const double* __restrict__ Bs, int wm0,
                                          int wn0, int lane, double (&acc)[DM_MB][DM_NB][2], Mid mid,
                                          const double* amean = nullptr, const double* bmean = nullptr,
                                          int kvalid = DM_BK) {
    if (!CENTER || kvalid >= DM_BK) {
        dmma_kstep<A_KMAJOR, B_KMAJOR, CENTER, false>(As, Bs, wm0, wn0, lane, 0, acc, amean, bmean, kvalid);
        mid();
#pragma unroll
        for (int k0 = 4; k0 < DM_BK; k0 += 4)
            dmma_kstep<A_KMAJOR, B_KMAJOR, CENTER, false>(As, Bs, wm0, wn0, lane, k0, acc, amean, bmean, kvalid);
    } else {
        mid();
#pragma unroll
        for (int k0 = 0; k0 < DM_BK; k0 += 4)
            dmma_kstep<A_KMAJOR, B_KMAJOR, CENTER, true>(As, Bs, wm0, wn0, lane, k0, acc, amean, bmean, kvalid);
    }
}

// ---- diagonal tiles of a SYRK -------------------------------------------------------------------------------------------
// Only the lower triangle of a diagonal 128 x 128 tile is needed: 10 of its 16 sub-blocks of 32 x 32, 4 of them (on the
// diagonal) only as 10 of their 16 m8n8 blocks.  Warps 0..5 take one full sub-block each, warps 6 and 7 two diagonal
// sub-blocks each: at most 20 DMMAs per warp and k-step instead of 32.  A and B fragments come from the same tile
// (k-major).  Sub-tile U of the warp lives in acc[4 U .. 4 U + 3][*]; its means in amean / bmean[4 U .. 4 U + 3].
template <bool TRI, bool RAGGED, int U, bool KMAJOR = true, bool CENTER = true>
__device__ __forceinline__ void dmma_kstep_sub(const double* __restrict__ As, int r0, int c0, int lane, int k0,
                                               double (&acc)[DM_MB][DM_NB][2], const double* amean,
                                               const double* bmean, int kvalid) {
    const int g = lane >> 2, t = lane & 3;
    double a[4], b[4];
    const bool live = !RAGGED || (k0 + t < kvalid);
    const double* base = KMAJOR ? As + (k0 + t) * DM_LD_KMAJOR + g : As + g * DM_LD_MMAJOR + k0 + t;
    constexpr int RS = KMAJOR ? 1 : DM_LD_MMAJOR;     // stride between consecutive rows / columns of the tile
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
        a[mb] = base[(r0 + mb * 8) * RS];
        if (CENTER) a[mb] -= amean[4 * U + mb];
        if (RAGGED) a[mb] = live ? a[mb] : 0.0;
    }
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
        b[nb] = base[(c0 + nb * 8) * RS];
        if (CENTER) b[nb] -= bmean[4 * U + nb];
        if (RAGGED) b[nb] = live ? b[nb] : 0.0;
    }
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int nb = 0; nb < 4; ++nb)
            if (!TRI || nb <= mb) dmma_884(acc[4 * U + mb][nb][0], acc[4 * U + mb][nb][1], a[mb], b[nb]);
}

// sub-blocks (row, column in units of 32) of warp `wid` in a diagonal tile; two = warps 6, 7
__device__ __forceinline__ void dmma_diag_subtiles(int wid, int& r0a, int& c0a, int& r0b, int& c0b) {
    if (wid < 6) {
        const int rb = wid < 3 ? wid + 1 : (wid < 5 ? wid - 1 : 3), cb = wid < 3 ? 0 : (wid < 5 ? 1 : 2);
        r0a = 32 * rb; c0a = 32 * cb; r0b = r0a; c0b = c0a;
    } else {
        r0a = c0a = 64 * (wid - 6);
        r0b = c0b = r0a + 32;
    }
}

template <bool KMAJOR = true, bool CENTER = true, class Mid>
__device__ __forceinline__ void dmma_slab_diag(const double* __restrict__ As, int wid, int lane,
                                               double (&acc)[DM_MB][DM_NB][2], Mid mid, const double* amean = nullptr,
                                               const double* bmean = nullptr, int kvalid = DM_BK) {
    int r0a, c0a, r0b, c0b;
    dmma_diag_subtiles(wid, r0a, c0a, r0b, c0b);
    auto step = [&](int k0, auto ragged) {
        constexpr bool R = decltype(ragged)::value;
        if (wid < 6) {
            dmma_kstep_sub<false, R, 0, KMAJOR, CENTER>(As, r0a, c0a, lane, k0, acc, amean, bmean, kvalid);
        } else {
            dmma_kstep_sub<true, R, 0, KMAJOR, CENTER>(As, r0a, c0a, lane, k0, acc, amean, bmean, kvalid);
            dmma_kstep_sub<true, R, 1, KMAJOR, CENTER>(As, r0b, c0b, lane, k0, acc, amean, bmean, kvalid);
        }
    };
    if (!CENTER || kvalid >= DM_BK) {      // without centring the zero fill of a ragged slab contributes nothing
        step(0, std::false_type{});
        mid();
#pragma unroll
        for (int k0 = 4; k0 < DM_BK; k0 += 4) step(k0, std::false_type{});
    } else {
        mid();
#pragma unroll
        for (int k0 = 0; k0 < DM_BK; k0 += 4) step(k0, std::true_type{});
    }
}
