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
__device__ __forceinline__ void sbr_fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }

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

int mcosb_launch_sb2st(mcosb_ctx* ctx, int S, const double* A, double* d, double* e, double* Q2) {
    const int N = ctx->N;
    const size_t smem = (size_t)(N + 16) * 16 * sizeof(double);
    if (smem + 1024 > (size_t)ctx->smem_optin || N > 768) return 0;
    static int Wsel = -1;
    if (Wsel < 0) {
        const char* e = getenv("MCOSB_SB2ST_W");
        Wsel = e ? atoi(e) : 4;
    }
#define SB2ST_LAUNCH(W)                                                                                               \
    do {                                                                                                              \
        MCOSB_CUDA(ctx, cudaFuncSetAttribute(sb2st_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        sb2st_kernel<W><<<S, W * 32, smem, ctx->stream>>>(A, N, d, e, Q2);                                            \
    } while (0)
    if (Wsel == 1) SB2ST_LAUNCH(1);
    else if (Wsel == 2) SB2ST_LAUNCH(2);
    else if (Wsel == 8) SB2ST_LAUNCH(8);
    else SB2ST_LAUNCH(4);
#undef SB2ST_LAUNCH
    MCOSB_LAUNCHED(ctx, MK_SB2ST);
    return 1;
}

// ---------------------------------------------------------------------------------------------------------------
// y <- Q2 y for the inverse-iteration vectors of one simulation (layout of inviter_kernel: y[i * G + g]).  The
// reflectors of one sweep act on disjoint row windows, so a sweep is one parallel step over (reflector, vector) pairs;
// sweeps run last to first.  The vectors sit in shared memory as Y[row][vector].
// ---------------------------------------------------------------------------------------------------------------
#define Q2B_NT 256
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
    constexpr int PF = 3;                          // N <= 768
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

bool mcosb_sbr_supported(mcosb_ctx* ctx) {
    const int N = ctx->N;
    extern size_t sbr_stage1_smem(int N);
    return N >= SBR_B + 2 && N <= 768 && sbr_stage1_smem(N) <= (size_t)ctx->smem_optin &&
           (size_t)(N + 16) * 16 * sizeof(double) + 1024 <= (size_t)ctx->smem_optin;
}
