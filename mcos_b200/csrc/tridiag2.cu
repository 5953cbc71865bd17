// D2 (second generation): blocked Householder tridiagonalisation on the LOWER triangle only.
//
// tridiag_kernel (denoise.cu) streams the whole trailing block through HBM once per Householder step, read + write:
// 16 N^3 / 3 bytes per matrix (628 MB measured at N = 500, profiles/tridiag_v1_r01.md).  This kernel follows LAPACK's
// dsytrd/dlatrd structure instead: inside a panel of NB columns the matrix is NOT updated -- the symmetric
// matrix-vector product of every step reads the panel-start matrix (lower triangle only, 4 m^2 bytes) and is corrected
// with the NB accumulated (V, W) pairs held in shared memory -- and the rank-2NB update A -= V W' + W V' is applied once
// per panel (8 m^2 bytes per NB steps).  Traffic: (4 + 8/NB) N^3 / 3 bytes = 1.5 N^3 at NB = 16: 3.6x less than v1.
//
// symv on the lower triangle, one CTA (16 warps) per matrix: a warp owns 16-row block-rows (paired long+short for
// balance) and sweeps 32-column tiles; a lane owns one column of the tile, so  y_c += A[r][c] v[r]  accumulates
// privately per lane, and  y_r += A[r][c] v[c]  accumulates in 16 registers that are reduced across the warp once per
// block-row.  Per-warp partial y vectors live in shared memory and are summed in a fixed order (deterministic).
#include "common.cuh"

#define T2_THREADS 512
#define T2_NW 16
#define T2_RB 16

template <int NB>
__global__ void __launch_bounds__(T2_THREADS, 1)
tridiag2_kernel(double* __restrict__ Aall, int N, double* __restrict__ dall, double* __restrict__ eall,
                double* __restrict__ tauall) {
    extern __shared__ double sm[];
    double* Vp = sm;                         // [NB][N]
    double* Wp = Vp + (size_t)NB * N;        // [NB][N]
    double* v = Wp + (size_t)NB * N;         // [N]
    double* y = v + N;                       // [N]
    double* ypart = y + N;                   // [T2_NW][N]
    double* tt = ypart + (size_t)T2_NW * N;  // [2 NB]
    double* red = tt + 2 * NB;               // [32]
    __shared__ double s_tau, s_beta;

    const int s = blockIdx.x;
    double* A = Aall + (size_t)s * N * N;
    double* d = dall + (size_t)s * N;
    double* e = eall + (size_t)s * N;
    double* tau = tauall + (size_t)s * N;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    for (int i = tid; i < 2 * NB * N; i += T2_THREADS) sm[i] = 0.0;
    __syncthreads();

    for (int k0 = 0; k0 + 2 < N; k0 += NB) {
        const int nbv = min(NB, N - 2 - k0);
        for (int j = 0; j < nbv; ++j) {
            const int k = k0 + j, base = k + 1, m = N - base;
            // ---- 1. column k of the current matrix = panel-start column minus the panel's rank-2j correction ----------
            for (int i = k + tid; i < N; i += T2_THREADS) {
                double a = A[(size_t)i * N + k];
                for (int p = 0; p < j; ++p)
                    a -= Vp[p * N + i] * Wp[p * N + k] + Wp[p * N + i] * Vp[p * N + k];
                y[i] = a;
            }
            __syncthreads();
            // ---- 2. Householder reflector ---------------------------------------------------------------------------
            double xsq = 0.0;
            for (int i = base + 1 + tid; i < N; i += T2_THREADS) xsq = fma(y[i], y[i], xsq);
            xsq = block_sum(xsq, red);
            if (tid == 0) {
                const double alpha = y[base];
                double t = 0.0, beta = alpha;
                if (xsq != 0.0) {
                    beta = -copysign(sqrt(alpha * alpha + xsq), alpha);
                    t = (beta - alpha) / beta;
                }
                s_tau = t;
                s_beta = beta;
                e[k] = beta;
                tau[k] = t;
                d[k] = y[k];
            }
            __syncthreads();
            const double tk = s_tau;
            {
                const double scal = (tk != 0.0) ? 1.0 / (y[base] - s_beta) : 0.0;
                for (int i = base + tid; i < N; i += T2_THREADS) {
                    const double vi = (i == base) ? 1.0 : y[i] * scal;
                    v[i] = vi;
                    Vp[j * N + i] = vi;
                    if (i > base) A[(size_t)k * N + i] = vi;  // reflector kept in row k (upper part, never read here)
                }
            }
            __syncthreads();
            // ---- 3. y = A_sym(base.., base..) v from the lower triangle of the panel-start matrix --------------------
            {
                double* yp = ypart + (size_t)wid * N;
                for (int i = base + lane; i < N; i += 32) yp[i] = 0.0;
                __syncwarp();
                const int nbr = (m + T2_RB - 1) / T2_RB;
                const int units = (nbr + 1) / 2;
                for (int u = wid; u < units; u += T2_NW) {
#pragma unroll 1
                    for (int half = 0; half < 2; ++half) {
                        const int b = half ? (nbr - 1 - u) : u;
                        if (half && b == u) break;
                        const int r0 = base + T2_RB * b;
                        double vr[T2_RB], rowp[T2_RB];
#pragma unroll
                        for (int t = 0; t < T2_RB; ++t) {
                            vr[t] = (r0 + t < N) ? v[r0 + t] : 0.0;
                            rowp[t] = 0.0;
                        }
                        const int cend = min(r0 + T2_RB - 1, N - 1);
                        for (int c0 = base; c0 <= cend; c0 += 32) {
                            const int c = c0 + lane;
                            const bool cact = c <= cend;
                            const double vc = cact ? v[c] : 0.0;
                            double av[T2_RB];
#pragma unroll
                            for (int t = 0; t < T2_RB; ++t) {
                                const int r = r0 + t;
                                av[t] = (cact && r < N && c <= r) ? A[(size_t)r * N + c] : 0.0;
                            }
                            double colacc = 0.0;
#pragma unroll
                            for (int t = 0; t < T2_RB; ++t) {
                                colacc = fma(av[t], vr[t], colacc);
                                rowp[t] = fma(av[t], (c != r0 + t) ? vc : 0.0, rowp[t]);  // diagonal counted once
                            }
                            if (cact) yp[c] += colacc;
                        }
                        __syncwarp();
                        double mine = 0.0;
#pragma unroll
                        for (int t = 0; t < T2_RB; ++t) {
                            const double r = warp_sum(rowp[t]);
                            if (lane == t) mine = r;
                        }
                        if (lane < T2_RB && r0 + lane < N) yp[r0 + lane] += mine;
                        __syncwarp();
                    }
                }
            }
            __syncthreads();
            for (int i = base + tid; i < N; i += T2_THREADS) {
                double t = 0.0;
#pragma unroll
                for (int w = 0; w < T2_NW; ++w) t += ypart[(size_t)w * N + i];
                y[i] = t;
            }
            // ---- 4. panel corrections: y -= V (W'v) + W (V'v) -------------------------------------------------------
            if (wid < j) {
                double t1 = 0.0, t2 = 0.0;
                for (int i = base + lane; i < N; i += 32) {
                    const double vi = v[i];
                    t1 = fma(Wp[wid * N + i], vi, t1);
                    t2 = fma(Vp[wid * N + i], vi, t2);
                }
                t1 = warp_sum(t1);
                t2 = warp_sum(t2);
                if (lane == 0) { tt[wid] = t1; tt[NB + wid] = t2; }
            }
            __syncthreads();
            double dot = 0.0;
            for (int i = base + tid; i < N; i += T2_THREADS) {
                double yi = y[i];
                for (int p = 0; p < j; ++p) yi -= Vp[p * N + i] * tt[p] + Wp[p * N + i] * tt[NB + p];
                y[i] = yi;
                dot = fma(tk * yi, v[i], dot);
            }
            // ---- 5. w = tau y - (tau/2)(tau y . v) v ----------------------------------------------------------------
            dot = block_sum(dot, red);
            const double a2 = -0.5 * tk * dot;
            for (int i = base + tid; i < N; i += T2_THREADS) Wp[j * N + i] = tk * y[i] + a2 * v[i];
            __syncthreads();
        }
        // ---- panel flush: A[i][c] -= sum_p V[p][i] W[p][c] + W[p][i] V[p][c]  for i >= c >= k1 -----------------------
        const int k1 = k0 + nbv;
        const int mm = N - k1;
        const int nch = (mm + 31) >> 5;
        const int units = (nch + 1) / 2;
        const int nrg = (units >= T2_NW) ? 1 : T2_NW / units;
        if (units >= T2_NW || wid < units * nrg) {
            const int rg = (units >= T2_NW) ? 0 : wid / units;
            for (int u = (units >= T2_NW) ? wid : wid % units; u < units; u += T2_NW) {
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    const int ch = half ? (nch - 1 - u) : u;
                    if (half && ch == u) break;
                    const int c = k1 + 32 * ch + lane;
                    const bool act = c < N;
                    double vc[NB], wc[NB];
#pragma unroll
                    for (int p = 0; p < NB; ++p) {
                        vc[p] = (act && p < nbv) ? Vp[p * N + c] : 0.0;
                        wc[p] = (act && p < nbv) ? Wp[p * N + c] : 0.0;
                    }
                    for (int i = k1 + 32 * ch + rg; i < N; i += nrg) {
                        if (act && c <= i) {
                            double a = A[(size_t)i * N + c];
#pragma unroll
                            for (int p = 0; p < NB; ++p) a -= Vp[p * N + i] * wc[p] + Wp[p * N + i] * vc[p];
                            A[(size_t)i * N + c] = a;
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        if (N == 1) {
            d[0] = A[0];
        } else {
            const int k = N - 2;
            d[k] = A[(size_t)k * N + k];
            e[k] = A[(size_t)(k + 1) * N + k];
            d[k + 1] = A[(size_t)(k + 1) * N + k + 1];
            tau[k] = 0.0;
        }
        e[N - 1] = 0.0;
        tau[N - 1] = 0.0;
    }
}

static size_t t2_smem(int N, int nb) { return ((size_t)(2 * nb + 2 + T2_NW) * N + 2 * nb + 32) * sizeof(double); }

// Returns 1 if launched, 0 if this size is not supported (caller falls back to tridiag_kernel), <0 on error.
int mcosb_launch_tridiag2(mcosb_ctx* ctx, int S, double* A, double* d, double* e, double* tau) {
    const int N = ctx->N;
    if (N < 96) return 0;  // small matrices: the fused v1 kernel is already cheap and has less fixed overhead
    if (t2_smem(N, 16) <= (size_t)ctx->smem_optin) {
        const size_t smem = t2_smem(N, 16);
        MCOSB_CUDA(ctx, cudaFuncSetAttribute(tridiag2_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tridiag2_kernel<16><<<S, T2_THREADS, smem, ctx->stream>>>(A, N, d, e, tau);
        return 1;
    }
    if (t2_smem(N, 8) <= (size_t)ctx->smem_optin) {
        const size_t smem = t2_smem(N, 8);
        MCOSB_CUDA(ctx, cudaFuncSetAttribute(tridiag2_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tridiag2_kernel<8><<<S, T2_THREADS, smem, ctx->stream>>>(A, N, d, e, tau);
        return 1;
    }
    return 0;
}
