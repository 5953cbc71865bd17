// D2 (second generation): blocked Householder tridiagonalisation on the LOWER triangle only.
//
// tridiag_kernel (denoise.cu) streams the whole trailing block through HBM once per Householder step, read + write:
// 16 N^3 / 3 bytes per matrix (628 MB measured at N = 500, profiles/tridiag_v1_r01.md).  This kernel follows LAPACK's
// dsytrd/dlatrd structure instead: inside a panel of NB columns the matrix is NOT updated -- the symmetric
// matrix-vector product of every step reads the panel-start matrix (lower triangle only, 4 m^2 bytes) and is corrected
// with the NB accumulated (V, W) pairs held in shared memory -- and the rank-2NB update A -= V W' + W V' is applied once
// per panel (8 m^2 bytes per NB steps).  Traffic: (4 + 8/NB) N^3 / 3 bytes.
//
// symv on the lower triangle, one CTA (NW warps) per matrix.  The lower triangle is cut into 16-row x 32-column tiles, enumerated block-row by
// block-row and dealt to the warps in equal contiguous runs.  Inside a tile a lane owns one column:
//   y_c += A[r][c] v[r]   accumulates privately in the lane;
//   y_r += A[r][c] v[c]   accumulates in 16 registers, reduced across the warp (halving butterfly, 16 shuffles) when
//                         the warp leaves the block-row.
// The 16 loads of the next tile are issued before the current tile is consumed (register double buffering), and the
// FP64 accumulations are split into independent chains: measured with clock64 (T2_TIMING), dependent DFMA issue
// latency and one-tile-in-flight memory latency, not bandwidth, bound this kernel.
// Per-warp partial y vectors live in shared memory and are summed in a fixed order (deterministic).  The first column
// of the next step falls out of the sweep (lane 0 of the first tile of every block-row) and is stashed in shared
// memory, so no step waits on a strided column read from global memory.
#include "common.cuh"

#ifdef T2_TIMING
#define T2_TICK(idx) do { if (threadIdx.x == 0) { long long now__ = clock64(); t2_acc[idx] += now__ - t2_last; t2_last = now__; } } while (0)
#else
#define T2_TICK(idx) do { } while (0)
#endif

// Block-row i (RB rows) holds RB (i + 1) columns = i / G + 1 tiles of 32 columns, G = 32 / RB.
// Tiles before block-row b = G h + r:  G h (h + 1) / 2 + r (h + 1).
template <int RB>
__device__ __forceinline__ int t2_ntile(int b) { return b / (32 / RB) + 1; }
template <int RB>
__device__ __forceinline__ int t2_prefix(int b) {
    constexpr int G = 32 / RB;
    const int h = b / G, r = b % G;
    return G * h * (h + 1) / 2 + r * (h + 1);
}
template <int RB>
__device__ __forceinline__ int t2_block_of_tile(int t) {
    constexpr int G = 32 / RB;
    int h = (int)((sqrt(1.0 + 8.0 * t / G) - 1.0) * 0.5);
    while (G * (h + 1) * (h + 2) / 2 <= t) ++h;
    while (G * h * (h + 1) / 2 > t) --h;
    return G * h + (t - G * h * (h + 1) / 2) / (h + 1);
}

// Sum RB per-lane row partials over the 32 lanes with a halving butterfly; the lane with (lane % (32 / RB)) == 0 and
// index lane / (32 / RB) ends up owning the total of row lane / (32 / RB).
template <int RB>
__device__ __forceinline__ double t2_row_reduce(double (&rowp)[RB], int lane) {
    double cur[RB];
#pragma unroll
    for (int q = 0; q < RB; ++q) cur[q] = rowp[q];
    int width = RB;
#pragma unroll
    for (int bit = 16; bit >= 32 / RB; bit >>= 1) {   // RB -> RB/2 -> ... -> 1 values per lane
        const bool up = lane & bit;
        width >>= 1;
#pragma unroll
        for (int q = 0; q < RB / 2; ++q) {
            if (q < width) {
                const double send = up ? cur[q] : cur[q + width];
                const double keep = up ? cur[q + width] : cur[q];
                cur[q] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
            }
        }
    }
    double r = cur[0];
#pragma unroll
    for (int bit = 32 / RB / 2; bit >= 1; bit >>= 1) r += __shfl_xor_sync(0xffffffffu, r, bit);
    return r;
}

template <int NB, int NW, int RB>
__global__ void __launch_bounds__(NW * 32, 1)
tridiag2_kernel(double* __restrict__ Aall, int N, double* __restrict__ dall, double* __restrict__ eall,
                double* __restrict__ tauall) {
    constexpr int NT = NW * 32;
    extern __shared__ double sm[];
    double* Vp = sm;                         // [NB][N]
    double* Wp = Vp + (size_t)NB * N;        // [NB][N]
    double* v = Wp + (size_t)NB * N;         // [N]
    double* y = v + N;                       // [N]
    double* colnext = y + N;                 // [N]  column base of the panel-start matrix, stashed by the sweep
    double* ypart = colnext + N;             // [NW][N]
    double* tt = ypart + (size_t)NW * N;     // [2 NB]
    double* red = tt + 2 * NB;               // [32]
    __shared__ double s_tau, s_beta;

    const int s = blockIdx.x;
    double* A = Aall + (size_t)s * N * N;
    double* d = dall + (size_t)s * N;
    double* e = eall + (size_t)s * N;
    double* tau = tauall + (size_t)s * N;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    for (int i = tid; i < 2 * NB * N; i += NT) sm[i] = 0.0;
    __syncthreads();
#ifdef T2_TIMING
    long long t2_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t2_last = clock64();
    long long t2_dbg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif

    for (int k0 = 0; k0 + 2 < N; k0 += NB) {
        const int nbv = min(NB, N - 2 - k0);
        for (int j = 0; j < nbv; ++j) {
            const int k = k0 + j, base = k + 1, m = N - base;
            // ---- 1. column k of the current matrix = panel-start column minus the panel's rank-2j correction ----------
            for (int i = k + tid; i < N; i += NT) {
                double a = (j == 0) ? A[(size_t)i * N + k] : colnext[i];
                for (int p = 0; p < j; ++p) {
                    a = fma(-Vp[p * N + i], Wp[p * N + k], a);
                    a = fma(-Wp[p * N + i], Vp[p * N + k], a);
                }
                y[i] = a;
            }
            __syncthreads();
            T2_TICK(0);
            // ---- 2. Householder reflector ---------------------------------------------------------------------------
            double xsq = 0.0;
            for (int i = base + 1 + tid; i < N; i += NT) xsq = fma(y[i], y[i], xsq);
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
                for (int i = base + tid; i < N; i += NT) {
                    const double vi = (i == base) ? 1.0 : y[i] * scal;
                    v[i] = vi;
                    Vp[j * N + i] = vi;
                    if (i > base) A[(size_t)k * N + i] = vi;  // reflector kept in row k (upper part, never read here)
                }
            }
            __syncthreads();
            T2_TICK(1);
            // ---- 3. y = A_sym(base.., base..) v from the lower triangle of the panel-start matrix --------------------
            {
                double* yp = ypart + (size_t)wid * N;
                for (int i = base + lane; i < N; i += 32) yp[i] = 0.0;
                __syncwarp();
                const int nbr = (m + RB - 1) / RB;
                const int total = t2_prefix<RB>(nbr);
                int t = (int)(((long long)total * wid) / NW);
                const int t1 = (int)(((long long)total * (wid + 1)) / NW);
                if (t < t1) {
                    int b = t2_block_of_tile<RB>(t);
                    int jt = t - t2_prefix<RB>(b);
                    // register pipeline: the loads of the next two tiles are in flight while one tile is consumed
                    auto load_tile = [&](int bb, int jj, double (&dst)[RB]) {
                        const int r0i = base + RB * bb, c0i = base + 32 * jj, ci = c0i + lane;
                        if (r0i + RB <= N && c0i + 31 <= r0i) {
                            const double* src = A + (size_t)r0i * N + ci;
#pragma unroll
                            for (int q = 0; q < RB; ++q) dst[q] = __ldcg(src + (size_t)q * N);   // independent addresses
                        } else {
#pragma unroll
                            for (int q = 0; q < RB; ++q) {
                                const int r = r0i + q;
                                dst[q] = ((ci < N) && (r < N) && (ci <= r)) ? __ldcg(A + (size_t)r * N + ci) : 0.0;
                            }
                        }
                    };
                    auto advance = [&](int& bb, int& jj) {
                        if (++jj >= t2_ntile<RB>(bb)) { jj = 0; ++bb; }
                    };
                    double n1[RB], n2[RB];
                    int b1 = b, j1 = jt;          // tile t (held in n1), then t + 1 (n2)
                    load_tile(b1, j1, n1);
                    int b2 = b1, j2 = j1;
                    advance(b2, j2);
                    if (t + 1 < t1) load_tile(b2, j2, n2);
                    while (t < t1) {
                        const int r0 = base + RB * b;
                        const int ntile = t2_ntile<RB>(b);
                        double vr[RB], rowp[RB];
#pragma unroll
                        for (int q = 0; q < RB; ++q) {
                            vr[q] = (r0 + q < N) ? v[r0 + q] : 0.0;
                            rowp[q] = 0.0;
                        }
                        for (; jt < ntile && t < t1; ++jt, ++t) {
                            double av[RB];
#pragma unroll
                            for (int q = 0; q < RB; ++q) { av[q] = n1[q]; n1[q] = n2[q]; }
                            advance(b2, j2);
                            if (t + 2 < t1) load_tile(b2, j2, n2);
                            const int c0 = base + 32 * jt, c = c0 + lane;
                            if (c0 + 31 <= r0) {  // no diagonal element in this tile
                                const double vc = v[c];
                                double ca[4] = {0.0, 0.0, 0.0, 0.0};   // four chains: FP64 dependent-issue latency is long
#pragma unroll
                                for (int q = 0; q < RB; ++q) {
                                    ca[q & 3] = fma(av[q], vr[q], ca[q & 3]);
                                    rowp[q] = fma(av[q], vc, rowp[q]);
                                }
                                yp[c] += (ca[0] + ca[1]) + (ca[2] + ca[3]);
                            } else {
                                const bool cact = c < N;
                                const double vc = cact ? v[c] : 0.0;
                                double ca[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                                for (int q = 0; q < RB; ++q) {
                                    ca[q & 3] = fma(av[q], vr[q], ca[q & 3]);
                                    rowp[q] = fma(av[q], (c != r0 + q) ? vc : 0.0, rowp[q]);  // diagonal counted once
                                }
                                if (cact) yp[c] += (ca[0] + ca[1]) + (ca[2] + ca[3]);
                            }
                            if (jt == 0 && lane == 0) {  // column `base`: the next step's column
#pragma unroll
                                for (int q = 0; q < RB; ++q)
                                    if (r0 + q < N) colnext[r0 + q] = av[q];
                            }
                        }
                        __syncwarp();
                        const double rsum = t2_row_reduce<RB>(rowp, lane);
                        constexpr int G = 32 / RB;
                        const int rr = r0 + lane / G;
                        if ((lane % G) == 0 && rr < N) yp[rr] += rsum;
                        __syncwarp();
                        ++b;
                        jt = 0;
                    }
                }
            }
            __syncthreads();
            T2_TICK(2);
            for (int i = base + tid; i < N; i += NT) {
                double acc = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) acc += ypart[(size_t)w * N + i];
                y[i] = acc;
            }
            // ---- 4. panel corrections: y -= V (W'v) + W (V'v) -------------------------------------------------------
            for (int p = wid; p < j; p += NW) {
                double t1s = 0.0, t2s = 0.0;
                for (int i = base + lane; i < N; i += 32) {
                    const double vi = v[i];
                    t1s = fma(Wp[p * N + i], vi, t1s);
                    t2s = fma(Vp[p * N + i], vi, t2s);
                }
                t1s = warp_sum(t1s);
                t2s = warp_sum(t2s);
                if (lane == 0) { tt[p] = t1s; tt[NB + p] = t2s; }
            }
            __syncthreads();
            T2_TICK(3);
            double dot = 0.0;
            for (int i = base + tid; i < N; i += NT) {
                double yi = y[i];
                for (int p = 0; p < j; ++p) {
                    yi = fma(-Vp[p * N + i], tt[p], yi);
                    yi = fma(-Wp[p * N + i], tt[NB + p], yi);
                }
                y[i] = yi;
                dot = fma(tk * yi, v[i], dot);
            }
            // ---- 5. w = tau y - (tau/2)(tau y . v) v ----------------------------------------------------------------
            dot = block_sum(dot, red);
            const double a2 = -0.5 * tk * dot;
            for (int i = base + tid; i < N; i += NT) Wp[j * N + i] = fma(tk, y[i], a2 * v[i]);
            __syncthreads();
            T2_TICK(4);
        }
        // ---- panel flush: A[i][c] -= sum_p V[p][i] W[p][c] + W[p][i] V[p][c]  for i >= c >= k1 -----------------------
        const int k1 = k0 + nbv;
        const int mm = N - k1;
        const int nch = (mm + 31) >> 5;
        const int units = (nch + 1) / 2;
        const int nrg = (units >= NW) ? 1 : NW / units;
        if (units >= NW || wid < units * nrg) {
            const int rg = (units >= NW) ? 0 : wid / units;
            for (int u = (units >= NW) ? wid : wid % units; u < units; u += NW) {
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    const int ch = half ? (nch - 1 - u) : u;
                    if (half && ch == u) break;
                    const int c = k1 + 32 * ch + lane;
                    const bool act = c < N;
                    double vc[NB], wc[NB];
#pragma unroll
                    for (int p = 0; p < NB; ++p) {
                        vc[p] = (act && p < nbv) ? -Vp[p * N + c] : 0.0;
                        wc[p] = (act && p < nbv) ? -Wp[p * N + c] : 0.0;
                    }
                    double* colp = A + c;
                    int i = k1 + 32 * ch + rg;
                    for (; i + 3 * nrg < N; i += 4 * nrg) {   // four rows in flight per lane
                        double a4[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) a4[q] = (act && c <= i + q * nrg) ? colp[(size_t)(i + q * nrg) * N] : 0.0;
#pragma unroll
                        for (int p = 0; p < NB; ++p) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                a4[q] = fma(Vp[p * N + i + q * nrg], wc[p], a4[q]);
                                a4[q] = fma(Wp[p * N + i + q * nrg], vc[p], a4[q]);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (act && c <= i + q * nrg) colp[(size_t)(i + q * nrg) * N] = a4[q];
                    }
                    for (; i < N; i += nrg) {
                        if (act && c <= i) {
                            double a = colp[(size_t)i * N];
#pragma unroll
                            for (int p = 0; p < NB; ++p) {
                                a = fma(Vp[p * N + i], wc[p], a);
                                a = fma(Wp[p * N + i], vc[p], a);
                            }
                            colp[(size_t)i * N] = a;
                        }
                    }
                }
            }
        }
        __syncthreads();
        T2_TICK(5);
    }
#ifdef T2_TIMING
    if (tid == 0 && blockIdx.x == 0)
        printf("warp0: wait %lld issue %lld compute %lld tiles %lld reduce %lld blockrows %lld\n", t2_dbg[0], t2_dbg[1], t2_dbg[2], t2_dbg[3], t2_dbg[4], t2_dbg[5]);
    if (tid == 0 && blockIdx.x == 0)
        printf("t2 cycles: col %lld refl %lld sweep %lld ysum+t12 %lld corr+w %lld flush %lld\n", t2_acc[0], t2_acc[1], t2_acc[2], t2_acc[3], t2_acc[4], t2_acc[5]);
#endif
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

template <int NB, int NW>
static size_t t2_smem(int N) {
    return ((size_t)(2 * NB + 3 + NW) * N + 2 * NB + 32 ) * sizeof(double);
}

template <int NB, int NW, int RB>
static int t2_launch(mcosb_ctx* ctx, int S, double* A, double* d, double* e, double* tau) {
    const size_t smem = t2_smem<NB, NW>(ctx->N);
    MCOSB_CUDA(ctx, cudaFuncSetAttribute(tridiag2_kernel<NB, NW, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tridiag2_kernel<NB, NW, RB><<<S, NW * 32, smem, ctx->stream>>>(A, ctx->N, d, e, tau);
    return 1;
}

// Returns 1 if launched, 0 if this size is not supported (caller falls back to tridiag_kernel), <0 on error.
int mcosb_launch_tridiag2(mcosb_ctx* ctx, int S, double* A, double* d, double* e, double* tau) {
    const int N = ctx->N;
    if (N < 200) return 0;  // small matrices: the fused v1 kernel is faster (1.3 vs 2.9 us per matrix at N = 100)
    // measured on B200 at N = 500 (scratch/t2_time.cu): 16 warps x 8-row tiles 13.4 ms per 148 matrices, 8 warps x
    // 16-row tiles 13.7 ms; the smaller footprint (27 N doubles) carries N up to ~1050
    if (t2_smem<8, 16>(N) <= (size_t)ctx->smem_optin) return t2_launch<8, 16, 8>(ctx, S, A, d, e, tau);
    if (t2_smem<8, 8>(N) <= (size_t)ctx->smem_optin) return t2_launch<8, 8, 16>(ctx, S, A, d, e, tau);
    return 0;
}
