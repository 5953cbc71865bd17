// Stage 2 of the two-stage tridiagonalisation: symmetric band (half bandwidth 8, left in the lower band of A by
// sy2sb_kernel) -> tridiagonal, by bulge chasing, entirely in shared memory; and the matching back-transformation.
//
// Sweep s annihilates column s below the sub-diagonal with a length-8 Householder reflector on rows s+1 .. s+8; applied
// two-sidedly it fills the 8 x 8 block below the band (the bulge), whose first column is annihilated by the next
// reflector on rows s+9 .. s+16, and so on down the band (step k works on rows lo .. lo+7, lo = s + 1 + 8 k).  Only the
// first column of each bulge is removed; the rest stays until the next sweep reaches it, so the band is held with 16
// diagonals: Bs[j][d] = B[j + d][j], 128 bytes per column, 64 KB per 500 x 500 matrix -> three matrices per SM.
//
// One warp executes a step.  Three groups of 8 lanes hold 8 values each, addressed lo * 16 + off[m] with per-lane
// offsets that never change:
//   lanes  0- 7  column c of C_k     = B[lo.., lo-8+c]   left application   z -= tau (v . z) v
//   lanes  8-15  row r of D_k        = B[lo+r, lo..]     two-sided          z -= v_r w + w_r v,  w = p - (tau/2)(v . p) v
//   lanes 16-23  row r of C_{k+1}    = B[lo+8+r, lo..]   right application  z -= tau (z . v) v
// so every reduction is a private dot product; the only cross-lane traffic is the broadcast of x (lane 0's values) and
// of w.  Rows and columns beyond N are zero padding and stay zero, which handles the short windows at the end of the band.
//
// W warps work on one matrix as a software pipeline: warp u owns sweeps s = u (mod W) and may run step k of sweep s once
// sweep s - 1 has completed step k + 2 (the two then touch disjoint elements); progress is published through shared
// memory.  The reflectors go to Q2[s][lo - s - 1 + m] (tau in the slot of the implicit 1) for q2back_kernel.
#include <cstdlib>

#include "common.cuh"
#include "sbr.cuh"

// 1/x and 1/sqrt(x) to ~1 ulp: hardware approximation (2^-23) + two Newton steps
__device__ __forceinline__ double sbr_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}
__device__ __forceinline__ double sbr_rsqrt(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double h = 0.5 * x;
    r = fma(r, fma(-h * r, r, 0.5), r);
    r = fma(r, fma(-h * r, r, 0.5), r);
    return r;
}

// predicated shared-memory store without a branch
__device__ __forceinline__ void sbr_st_if(double* p, double v, unsigned pred) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.shared.f64 [%0], %1; }" ::"r"(a), "d"(v), "r"(pred) : "memory");
}

// acquire/release fence at CTA scope: orders the band updates against the progress flag (the sequentially consistent
// __threadfence_block is not needed)
#ifdef SBR_NO_FENCE
__device__ __forceinline__ void sbr_fence_cta() { asm volatile("" ::: "memory"); }
#else
__device__ __forceinline__ void sbr_fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }
#endif

template <int W>
__global__ void __launch_bounds__(W * 32) sb2st_kernel(const double* __restrict__ Aall, int N, double* __restrict__ dall,
                                                       double* __restrict__ eall, double* __restrict__ Q2all) {
    extern __shared__ double Bs[];                 // [(N + 16) * 16]
    __shared__ volatile int prog[W];
    const int sim = blockIdx.x;
    const double* A = Aall + (size_t)sim * N * N;
    double* Q2 = Q2all + (size_t)sim * N * N;
    const int tid = threadIdx.x, lane = tid & 31, u = tid >> 5;
    constexpr int NT = W * 32;

    for (int i = tid; i < (N + 16) * 16; i += NT) Bs[i] = 0.0;
    if (tid < W) prog[tid] = -1;
    __syncthreads();
    for (int i = tid; i < N; i += NT) {
        const double* row = A + (size_t)i * N;
        for (int d = 0; d <= SBR_B && d <= i; ++d) Bs[(i - d) * 16 + d] = row[i - d];
    }
    __syncthreads();

    const int grp = lane >> 3, l = lane & 7;
    int off[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        if (grp == 0) off[m] = (l - 8) * 16 + 8 - l + m;
        else if (grp == 1) off[m] = (l < m ? l : m) * 16 + (l < m ? m - l : l - m);
        else off[m] = m * 16 + 8 + l - m;
    }
    const bool g0 = grp == 0, g1 = grp == 1;
    // which of its 8 values a lane stores: group 0 and 2 all, group 1 the lower part of its row, group 3 nothing
    const unsigned smask = (grp == 3) ? 0u : (g1 ? ((2u << l) - 1u) : 0xffu);
    const int pu = (u + W - 1) % W;
    int seen = -1;   // last progress value read from the predecessor

    for (int s = u; s + 2 < N; s += W) {
        double xn[8];                             // first column of the next bulge, handed on in registers
        double unext = 0.0;                       // x[l] of the next step for this lane (u_r of the rank-2 update)
        for (int k = 0;; ++k) {
            const int lo = s + 1 + 8 * k;
            if (N - lo < 2) break;
            if (s > 0) {
                int need_done = k + 3;
                if (need_done > 255) need_done = 255;
                const int need = (s - 1) * 256 + need_done;   // the predecessor publishes ..255 when its sweep is complete
                if (seen < need) {
                    if (lane == 0) {
                        int cur = prog[pu];
                        while (cur < need) cur = prog[pu];
                        seen = cur;
                    }
                    seen = __shfl_sync(0xffffffffu, seen, 0);
                    sbr_fence_cta();
                }
            }
            const int base = lo * 16;
            const bool first = (k == 0);
            // step 0 has no block left of the window: group 0 reads (and ignores) valid data one block further down
            const int zb = (first && g0) ? base + 128 : base;
            double z[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) z[m] = Bs[zb + off[m]];
            double x[8], ur;
            if (first) {
#pragma unroll
                for (int m = 0; m < 8; ++m) x[m] = Bs[base - 15 + m];       // column s below the diagonal
                ur = x[0];
#pragma unroll
                for (int m = 1; m < 8; ++m) ur = (l == m) ? x[m] : ur;
            } else {
#pragma unroll
                for (int m = 0; m < 8; ++m) x[m] = xn[m];
                ur = unext;
            }
            // Householder in unnormalised form H = I - gam u u', u = (alpha - beta, x_1 .. x_7): only one rsqrt and one
            // reciprocal sit on the critical path (IEEE sqrt and two divisions cost ~350 cycles)
            const double ss = (x[1] * x[1] + x[2] * x[2]) + (x[3] * x[3] + x[4] * x[4]) +
                              ((x[5] * x[5] + x[6] * x[6]) + x[7] * x[7]);
            const double alpha = x[0];
            double gam = 0.0, tau = 0.0, scale = 0.0, beta = alpha, u0 = 0.0;
            if (ss > 1e-280) {
                const double n2 = fma(alpha, alpha, ss);
                const double rs = sbr_rsqrt(n2);
                const double tt = fma(fabs(alpha), rs, 1.0);          // tau = (beta - alpha) / beta
                const double rt = sbr_rcp(tt);
                beta = -copysign(n2 * rs, alpha);
                u0 = alpha - beta;
                gam = rs * rs * rt;                                   // 1 / (nrm (nrm + |alpha|))
                scale = copysign(rs * rt, alpha);                     // 1 / u0
                tau = tt;
            }
            if (l == 0) ur = u0;
            double dot = ((x[1] * z[1] + x[2] * z[2]) + (x[3] * z[3] + x[4] * z[4])) +
                         ((x[5] * z[5] + x[6] * z[6]) + x[7] * z[7]);
            dot = fma(u0, z[0], dot);
            const double p = gam * dot;
            // the right application's first column is the next reflector's x: hand it on before anything else
            {
                const double z0new = fma(-p, u0, z[0]);
#pragma unroll
                for (int m = 0; m < 8; ++m) xn[m] = __shfl_sync(0xffffffffu, z0new, 16 + m);
                unext = __shfl_sync(0xffffffffu, z0new, 16 + l);
            }
            double kp = ur * p;
            kp += __shfl_xor_sync(0xffffffffu, kp, 1);
            kp += __shfl_xor_sync(0xffffffffu, kp, 2);
            kp += __shfl_xor_sync(0xffffffffu, kp, 4);
            const double wr = fma(-0.5 * gam * kp, ur, p);
            // z -= ca u + cb w:  groups 0 / 2: ca = gam (u . z), cb = 0;  group 1 (row r of D): ca = w_r, cb = u_r
            const double ca = g1 ? wr : p;
            const double cb = g1 ? ur : 0.0;
            z[0] = fma(-cb, __shfl_sync(0xffffffffu, wr, 8), fma(-ca, u0, z[0]));
#pragma unroll
            for (int m = 1; m < 8; ++m) z[m] = fma(-cb, __shfl_sync(0xffffffffu, wr, 8 + m), fma(-ca, x[m], z[m]));
            if (first) {
                if (lane == 0) {                                   // column s: (beta, 0, ..., 0)
                    Bs[base - 15] = beta;
#pragma unroll
                    for (int m = 1; m < 8; ++m) Bs[base - 15 + m] = 0.0;
                }
                if (!g0) {
#pragma unroll
                    for (int m = 0; m < 8; ++m) sbr_st_if(&Bs[base + off[m]], z[m], (smask >> m) & 1u);
                }
            } else {
                if (lane == 0) {
                    z[0] = beta;
#pragma unroll
                    for (int m = 1; m < 8; ++m) z[m] = 0.0;
                }
#pragma unroll
                for (int m = 0; m < 8; ++m) sbr_st_if(&Bs[base + off[m]], z[m], (smask >> m) & 1u);
            }
            if (lane < 8 && lo + lane < N) Q2[(size_t)s * N + (lo - s - 1) + lane] = (lane == 0) ? tau : ur * scale;
            __syncwarp();
            sbr_fence_cta();
            if (lane == 0) prog[u] = s * 256 + k + 1;
        }
        __syncwarp();
        if (lane == 0) prog[u] = s * 256 + 255;
    }
    __syncthreads();
    double* d = dall + (size_t)sim * N;
    double* e = eall + (size_t)sim * N;
    for (int i = tid; i < N; i += NT) {
        d[i] = Bs[i * 16];
        e[i] = (i + 1 < N) ? Bs[i * 16 + 1] : 0.0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The same chase with the reflector applied as tiny GEMMs on the FP64 tensor core: H = I - gam u u' is formed explicitly in
// fragment layout (each lane needs five entries of u) and a step is 8 DMMA.8x8x4:
//   C_k' = H C_k (2),  T = H D_k, D_k' = T H (2 + 2; the C fragment of T is reused as A operand with the k index permuted),
//   C_{k+1}' = C_{k+1} H (2; H is symmetric, so its A fragments double as B fragments).
// No dot products, no reductions: the step shrinks from ~480 to ~190 instructions and its dependent chain from ~50 FP64
// operations to the Householder scalars plus two DMMA latencies (step 1030 cycles instead of 1350 for a lone warp).  But
// the explicit 8 x 8 x 8 products cost 128 FP64-pipe cycles per step against ~130 for ALL the FP64 work of the rank-1
// form, so with 24 warps per SM this kernel is FP64-pipe-bound where the other is issue-bound: 7.2 us per 500 x 500 matrix
// either way.  Kept selectable (MCOSB_SB2ST_MMA=1) and tested; same storage, pipeline protocol and Q2 layout as above;
// the follower sleeps on an mbarrier instead of polling.
// ---------------------------------------------------------------------------------------------------------------
#include "dmma.cuh"
#define SBR_MBARS 100   // >= steps per sweep + 2 (N <= 768: 96 steps)

// mbarrier helpers: the follower of a sweep sleeps in hardware (try_wait) instead of polling shared memory -- polling
// warps flood the shared-memory / shuffle pipe the working warp of the same SM sub-partition needs
__device__ __forceinline__ void sbr_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void sbr_mbar_arrive(unsigned long long* bar) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared.b64 st, [%0]; }" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void sbr_mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{ .reg .pred p;\n"
        "W_%=: mbarrier.try_wait.parity.shared.b64 p, [%0], %1;\n"
        "@!p bra W_%=;\n }" ::"r"(a), "r"(parity) : "memory");
}

template <int W>
__global__ void __launch_bounds__(W * 32) sb2st_mma_kernel(const double* __restrict__ Aall, int N, double* __restrict__ dall,
                                                           double* __restrict__ eall, double* __restrict__ Q2all) {
    extern __shared__ double Bs[];                 // [(N + 16) * 16]
    __shared__ volatile int prog[W];
    __shared__ unsigned long long mbar[W][SBR_MBARS];   // mbar[u][j]: warp u has finished step j of its current sweep
    const int sim = blockIdx.x;
    const double* A = Aall + (size_t)sim * N * N;
    double* Q2 = Q2all + (size_t)sim * N * N;
    const int tid = threadIdx.x, lane = tid & 31, u = tid >> 5;
    constexpr int NT = W * 32;

    for (int i = tid; i < (N + 16) * 16; i += NT) Bs[i] = 0.0;
    if (tid < W) prog[tid] = -1;
    for (int i = tid; i < W * SBR_MBARS; i += NT) sbr_mbar_init(&mbar[0][0] + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    for (int i = tid; i < N; i += NT) {
        const double* row = A + (size_t)i * N;
        for (int d = 0; d <= SBR_B && d <= i; ++d) Bs[(i - d) * 16 + d] = row[i - d];
    }
    __syncthreads();

    const int g = lane >> 2, t = lane & 3;
    // element (row lo + r, column lo + c) of the band sits at lo * 16 + c * 16 + (r - c); fixed per-lane offsets:
    const int oCkL = (g - 8) * 16 + 8 + t - g;                 // C_k[t][g]            (+4: row 4 + t)
    const int oCkS = (2 * t - 8) * 16 + 8 + g - 2 * t;         // C_k'[g][2t]          (+15: column 2t + 1)
    auto offD = [](int r, int c) { return r >= c ? c * 16 + r - c : r * 16 + c - r; };
    const int oDL0 = offD(t, g), oDL1 = offD(4 + t, g);        // D[t][g], D[4+t][g] (symmetric, lower stored)
    const int oDS = 2 * t * 16 + g - 2 * t;                    // D'[g][2t]            (+15: column 2t + 1)
    const int oEL = t * 16 + 8 + g - t;                        // E[g][t]              (+60: column 4 + t)
    const int oES = 2 * t * 16 + 8 + g - 2 * t;                // E'[g][2t]            (+15: column 2t + 1)
    const unsigned stD0 = 2 * t <= g, stD1 = 2 * t + 1 <= g;
    const double dA0 = (g == t) ? 1.0 : 0.0, dA1 = (g == 4 + t) ? 1.0 : 0.0;
    const double dB0 = (2 * t == g) ? 1.0 : 0.0, dB1 = (2 * t + 1 == g) ? 1.0 : 0.0;
    const int pu = (u + W - 1) % W;
    int seen = -1;
#ifdef SB_TIMING
    long long tw = 0, tc = 0, tmark = clock64();
    int nsteps = 0;
#endif

    for (int s = u; s + 2 < N; s += W) {
        double xsrc = 0.0;                        // E'[g][0] of the previous step (lanes with t == 0): the next x
        for (int k = 0;; ++k) {
            const int lo = s + 1 + 8 * k;
            if (N - lo < 2) break;
            if (s > 0) {
                int need_done = k + 3;
                if (need_done > 255) need_done = 255;
                const int need = (s - 1) * 256 + need_done;
                if (seen < need) {
                    if (lane == 0) {
                        int cur = prog[pu];
                        if (cur < need) {          // not there yet: sleep on the barrier of the step we depend on
                            sbr_mbar_wait(&mbar[pu][min(k + 2, SBR_MBARS - 1)], ((s - 1) / W) & 1);
                            cur = prog[pu];
                        }
                        seen = cur;
                    }
                    seen = __shfl_sync(0xffffffffu, seen, 0);
                    sbr_fence_cta();
                }
            }
#ifdef SB_TIMING
            { long long n0 = clock64(); tw += n0 - tmark; tmark = n0; }
#endif
            const int base = lo * 16;
            const bool first = (k == 0);
            // operands of the three blocks (C_k does not exist in step 0: valid memory is read and ignored)
            const int cb = first ? base + 128 : base;
            const double ck0 = Bs[cb + oCkL], ck1 = Bs[cb + oCkL + 4];
            const double d0 = Bs[base + oDL0], d1 = Bs[base + oDL1];
            const double e0 = Bs[base + oEL], e1 = Bs[base + oEL + 60];
            double x[8], ug, ut, ut4, u2t, u2t1, xq;
            if (first) {
                const double* col = Bs + base - 15;           // column s below the diagonal
#pragma unroll
                for (int m = 0; m < 8; ++m) x[m] = col[m];
                ug = col[g]; ut = col[t]; ut4 = col[4 + t]; u2t = col[2 * t]; u2t1 = col[2 * t + 1];
                xq = col[lane & 7];
            } else {
#pragma unroll
                for (int m = 0; m < 8; ++m) x[m] = __shfl_sync(0xffffffffu, xsrc, 4 * m);
                ug = __shfl_sync(0xffffffffu, xsrc, 4 * g);
                ut = __shfl_sync(0xffffffffu, xsrc, 4 * t);
                ut4 = __shfl_sync(0xffffffffu, xsrc, 16 + 4 * t);
                u2t = __shfl_sync(0xffffffffu, xsrc, 8 * t);
                u2t1 = __shfl_sync(0xffffffffu, xsrc, 8 * t + 4);
                xq = __shfl_sync(0xffffffffu, xsrc, 4 * (lane & 7));
            }
            const double ss = (x[1] * x[1] + x[2] * x[2]) + (x[3] * x[3] + x[4] * x[4]) +
                              ((x[5] * x[5] + x[6] * x[6]) + x[7] * x[7]);
            const double alpha = x[0];
            double gam = 0.0, tau = 0.0, scale = 0.0, beta = alpha, u0 = 0.0;
            if (ss > 1e-280) {
                const double n2 = fma(alpha, alpha, ss);
                const double rs = sbr_rsqrt(n2);
                const double tt = fma(fabs(alpha), rs, 1.0);
                const double rt = sbr_rcp(tt);
                beta = -copysign(n2 * rs, alpha);
                u0 = alpha - beta;
                gam = rs * rs * rt;
                scale = copysign(rs * rt, alpha);
                tau = tt;
            }
            if (g == 0) ug = u0;
            if (t == 0) { ut = u0; u2t = u0; }
            // H = I - gam u u' in fragment layout
            const double gu = gam * ug;
            const double hA0 = fma(-gu, ut, dA0), hA1 = fma(-gu, ut4, dA1);
            const double hB0 = fma(-gu, u2t, dB0), hB1 = fma(-gu, u2t1, dB1);
            double c0 = 0.0, c1 = 0.0, t0 = 0.0, t1 = 0.0, ee0 = 0.0, ee1 = 0.0, dd0 = 0.0, dd1 = 0.0;
            dmma_884(ee0, ee1, e0, hA0);               // E' = E H   (critical: its first column is the next x)
            dmma_884(t0, t1, hA0, d0);                 // T  = H D
            dmma_884(c0, c1, hA0, ck0);                // C' = H C
            dmma_884(ee0, ee1, e1, hA1);
            dmma_884(t0, t1, hA1, d1);
            dmma_884(c0, c1, hA1, ck1);
            dmma_884(dd0, dd1, t0, hB0);               // D' = T H   (k index permuted: even / odd columns of T)
            dmma_884(dd0, dd1, t1, hB1);
            xsrc = ee0;
            if (first) {
                if (lane < 8) Bs[base - 15 + lane] = (lane == 0) ? beta : 0.0;     // column s: (beta, 0, ..., 0)
            } else {
                if (t == 0) c0 = (g == 0) ? beta : 0.0;                             // first column of C_k: annihilated
                Bs[base + oCkS] = c0;
                Bs[base + oCkS + 15] = c1;
            }
            sbr_st_if(&Bs[base + oDS], dd0, stD0);
            sbr_st_if(&Bs[base + oDS + 15], dd1, stD1);
            Bs[base + oES] = ee0;
            Bs[base + oES + 15] = ee1;
            if (lane < 8 && lo + lane < N) Q2[(size_t)s * N + (lo - s - 1) + lane] = (lane == 0) ? tau : xq * scale;
            __syncwarp();
            sbr_fence_cta();
            if (lane == 0) {
                prog[u] = s * 256 + k + 1;
                sbr_mbar_arrive(&mbar[u][k]);
            }
#ifdef SB_TIMING
            { long long n0 = clock64(); tc += n0 - tmark; tmark = n0; ++nsteps; }
#endif
        }
        __syncwarp();
        if (lane == 0) {
            // the follower may wait for up to two steps this sweep does not have
            const int nst = (N - 2 - s + 7) / 8;   // steps of sweep s: lo = s + 1 + 8 k <= N - 2
            prog[u] = s * 256 + 255;
            sbr_mbar_arrive(&mbar[u][min(nst, SBR_MBARS - 1)]);
            sbr_mbar_arrive(&mbar[u][min(nst + 1, SBR_MBARS - 1)]);
        }
    }
#ifdef SB_TIMING
    if (blockIdx.x == 0 && lane == 0) printf("sb2st warp %d: steps %d wait %lld compute %lld (per step %lld / %lld)\n", u, nsteps, tw, tc, tw / max(nsteps, 1), tc / max(nsteps, 1));
#endif
    __syncthreads();
    double* d = dall + (size_t)sim * N;
    double* e = eall + (size_t)sim * N;
    for (int i = tid; i < N; i += NT) {
        d[i] = Bs[i * 16];
        e[i] = (i + 1 < N) ? Bs[i * 16 + 1] : 0.0;
    }
}

int mcosb_launch_sb2st(mcosb_ctx* ctx, int S, const double* A, double* d, double* e, double* Q2) {
    const int N = ctx->N;
    const size_t smem = (size_t)(N + 16) * 16 * sizeof(double);
    if (smem + 1024 > (size_t)ctx->smem_optin || N > 1024) return 0;
    static int Wenv = -1;
    if (Wenv < 0) {
        const char* e = getenv("MCOSB_SB2ST_W");
        Wenv = e ? atoi(e) : 0;
    }
    // warps pipelining the sweeps of one matrix.  Three matrices share an SM up to N = 512 (64 KB of band each): measured at
    // N = 500 (us per matrix): W = 3 8.50, 4 6.89, 5 6.60, 6 7.22, 8 7.26.  Above, the band takes the SM to two matrices and
    // then (N > 880) to one, and a sweep has up to 125 steps to pipeline: more warps per matrix (profiles/sbr_r02.md).
    const int Wsel = Wenv > 0 ? Wenv : (N <= 512 ? 5 : (N <= 880 ? 8 : 16));
#define SB2ST_LAUNCH(W)                                                                                               \
    do {                                                                                                              \
        MCOSB_CUDA(ctx, cudaFuncSetAttribute(sb2st_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        sb2st_kernel<W><<<S, W * 32, smem, ctx->stream>>>(A, N, d, e, Q2);                                            \
    } while (0)
    static int mma = -1;
    if (mma < 0) {
        const char* e2 = getenv("MCOSB_SB2ST_MMA");
        mma = e2 ? atoi(e2) : 0;   // measured: 7.2 us per 500 x 500 matrix either way (profiles/sbr_kernels_r01c.md)
    }
#define SB2ST_LAUNCH_MMA(W)                                                                                           \
    do {                                                                                                              \
        MCOSB_CUDA(ctx, cudaFuncSetAttribute(sb2st_mma_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        sb2st_mma_kernel<W><<<S, W * 32, smem, ctx->stream>>>(A, N, d, e, Q2);                                        \
    } while (0)
    if (mma) {
        if (Wsel == 1) SB2ST_LAUNCH_MMA(1);
        else if (Wsel == 2) SB2ST_LAUNCH_MMA(2);
        else if (Wsel == 8) SB2ST_LAUNCH_MMA(8);
        else SB2ST_LAUNCH_MMA(4);
    } else if (Wsel == 1) SB2ST_LAUNCH(1);
    else if (Wsel == 2) SB2ST_LAUNCH(2);
    else if (Wsel == 3) SB2ST_LAUNCH(3);
    else if (Wsel == 4) SB2ST_LAUNCH(4);
    else if (Wsel == 6) SB2ST_LAUNCH(6);
    else if (Wsel == 8) SB2ST_LAUNCH(8);
    else if (Wsel == 12) SB2ST_LAUNCH(12);
    else if (Wsel == 16) SB2ST_LAUNCH(16);
    else SB2ST_LAUNCH(5);
#undef SB2ST_LAUNCH
#undef SB2ST_LAUNCH_MMA
    MCOSB_LAUNCHED(ctx, MK_SB2ST);
    return 1;
}

// ---------------------------------------------------------------------------------------------------------------
// y <- Q2 y for the inverse-iteration vectors of one simulation (layout of inviter_kernel: y[i * G + g]).  The
// reflectors of one sweep act on disjoint row windows, so a sweep is one parallel step over (reflector, vector) pairs;
// sweeps run last to first.  The vectors sit in shared memory as Y[row][vector].
// ---------------------------------------------------------------------------------------------------------------
#ifndef Q2B_NT
#define Q2B_NT 256
#endif
#define Q2B_V 16   // = EV3_VMAX

__global__ void __launch_bounds__(Q2B_NT) q2back_kernel(const double* __restrict__ Q2all, const int* __restrict__ nfall,
                                                        int N, int S, double* __restrict__ Wy) {
    extern __shared__ double Y[];                  // [(N + 8)][16]
    const int sim = blockIdx.x, tid = threadIdx.x;
    const int nf = nfall[sim];
    if (nf > Q2B_V || nf <= 0) return;
    const size_t G = (size_t)S * Q2B_V;
    double* y = Wy + sim * (size_t)Q2B_V;          // element (i, vec) at y[i * G + vec]
    const double* Q2 = Q2all + (size_t)sim * N * N;
    for (int idx = tid; idx < (N + 8) * Q2B_V; idx += Q2B_NT) {
        const int i = idx >> 4, vec = idx & 15;
        Y[idx] = (i < N && vec < nf) ? y[(size_t)i * G + vec] : 0.0;
    }
    __syncthreads();
    // the reflectors of sweep s (N - 1 - s doubles) are staged in shared memory one sweep ahead: the global load of the
    // next row overlaps the work on the current one
    double* Qs = Y + (size_t)(N + 8) * Q2B_V;      // [2][N + 8]
    constexpr int PF = (1024 + Q2B_NT - 1) / Q2B_NT;   // N <= 1024
    for (int idx = tid; idx < 2 * (N + 8); idx += Q2B_NT) Qs[idx] = 0.0;
    __syncthreads();
    if (N >= 3)
        for (int idx = tid; idx < 2; idx += Q2B_NT) Qs[idx] = Q2[(size_t)(N - 3) * N + idx];   // sweep N-3: rows N-2, N-1
    __syncthreads();
    for (int s = N - 3; s >= 0; --s) {
        const int nk = (N - 3 - s) / 8 + 1;
        const int cur = (N - 3 - s) & 1;
        const double* q = Qs + (size_t)cur * (N + 8);
        double pf[PF];
#pragma unroll
        for (int r = 0; r < PF; ++r) {
            const int idx = tid + Q2B_NT * r;
            pf[r] = (s > 0 && idx < N - s) ? Q2[(size_t)(s - 1) * N + idx] : 0.0;
        }
        for (int item = tid; item < nk * Q2B_V; item += Q2B_NT) {
            const int k = item >> 4, vec = item & 15;
            if (vec < nf) {
                const int lo = s + 1 + 8 * k;
                const double tau = q[8 * k];
                double v[8], yy[8];
                v[0] = 1.0;
#pragma unroll
                for (int m = 1; m < 8; ++m) v[m] = (lo + m < N) ? q[8 * k + m] : 0.0;
                double dot = 0.0;
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    yy[m] = Y[(lo + m) * Q2B_V + vec];
                    dot = fma(v[m], yy[m], dot);
                }
                dot *= tau;
#pragma unroll
                for (int m = 0; m < 8; ++m) Y[(lo + m) * Q2B_V + vec] = fma(-dot, v[m], yy[m]);
            }
        }
        double* qn = Qs + (size_t)(cur ^ 1) * (N + 8);
#pragma unroll
        for (int r = 0; r < PF; ++r) {
            const int idx = tid + Q2B_NT * r;
            if (s > 0 && idx < N - s) qn[idx] = pf[r];
        }
        __syncthreads();
    }
    for (int idx = tid; idx < N * Q2B_V; idx += Q2B_NT) {
        const int i = idx >> 4, vec = idx & 15;
        if (vec < nf) y[(size_t)i * G + vec] = Y[idx];
    }
}

int mcosb_launch_q2back(mcosb_ctx* ctx, int S, const double* Q2, const int* nf, double* Wy) {
    const int N = ctx->N;
    const size_t smem = ((size_t)(N + 8) * Q2B_V + 2 * (size_t)(N + 8)) * sizeof(double);
    MCOSB_CUDA(ctx, cudaFuncSetAttribute(q2back_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    q2back_kernel<<<S, Q2B_NT, smem, ctx->stream>>>(Q2, nf, N, S, Wy);
    MCOSB_LAUNCHED(ctx, MK_Q2BACK);
    return 1;
}

// Which dense -> band kernel: the round-1 kernel (sy2sb.cu: one 16-warp CTA per SM, register ring of 2 tiles, one barrier per
// 32-row block row; N <= 696) or sy2sb2.cu (barrier-free pass with the Z partials in L2, cp.async tile ring, two matrices
// per SM for small N, N <= 1024).  Measured per matrix on B200 (profiles/sy2sb2_r02.md): N = 400: 10.9 vs 10.8 us,
// N = 500: 18.4 vs 21.0, N = 600: 35.9 vs 40.0, N = 696: 52.1 vs 75.4 -- removing the barrier and the register ring did not
// pay (the pass is bound by shared-memory / shuffle issue and DMMA dependency latency, not by the barrier), so the old
// kernel stays the default wherever it runs and the new one takes N = 697 .. 1024, where the alternative is the one-stage
// kernel at 527 us per matrix (N = 1000: 183 us).  MCOSB_SY2SB_OLD=0 forces the new kernel everywhere (A/B runs, tests).
// MCOSB_STAGE1 = 1 | 2 | 3 forces sy2sb.cu (N <= 696 only) | sy2sb2.cu | sy2sb3.cu (batched multi-kernel form) for A/B runs
// and tests; MCOSB_SY2SB_OLD (round 2, first session) is still honoured: 1 -> kind 1, 0 -> kind 2.
int sbr_stage1_kind(int N) {
    static int v = -2;
    if (v == -2) {
        const char* e = getenv("MCOSB_STAGE1");
        const char* o = getenv("MCOSB_SY2SB_OLD");
        v = e ? atoi(e) : (o ? (atoi(o) ? 1 : 2) : -1);
    }
    if (v == 1) return N <= 696 ? 1 : 2;
    if (v == 2 || v == 3) return v;
    return SBR_STAGE1_DEFAULT(N);
}

bool mcosb_sbr_supported(mcosb_ctx* ctx) {
    const int N = ctx->N;
    extern size_t sbr_stage1_smem(int N);
    const int kind = sbr_stage1_kind(N);
    const bool stage1 = kind == 1 ? (sbr_stage1_smem(N) <= (size_t)ctx->smem_optin)
                                  : (kind == 2 ? sbr2_supported(ctx) : sbr3_supported(ctx));
    const size_t q2b = ((size_t)(N + 8) * Q2B_V + 2 * (size_t)(N + 8)) * sizeof(double);
    return N >= SBR_B + 2 && N <= 1024 && stage1 && q2b <= (size_t)ctx->smem_optin &&
           (size_t)(N + 16) * 16 * sizeof(double) + 1024 <= (size_t)ctx->smem_optin;
}

int mcosb_launch_stage1(mcosb_ctx* ctx, int S, double* A, double* tau) {
    const int kind = sbr_stage1_kind(ctx->N);
    if (kind == 1) return mcosb_launch_sy2sb(ctx, S, A, tau);
    if (kind == 2) return mcosb_launch_sy2sb2(ctx, S, A, tau);
    return mcosb_launch_sy2sb3(ctx, S, A, tau);
}
