// K2: batched mean + sample covariance (FP64 tensor-core SYRK) with the Ledoit-Wolf and jackknife estimators as
// fused reductions.  Replaces observation_simulator.py:28 (LedoitWolf().fit), :40 (np.cov), :53-57 (jackknife loops).
//
//   mean_j   = sum_t X[t][j] / T
//   S        = Xc' Xc / (T-1)                         (MuCov; Jackknife collapses to the same thing)
//   S_b      = Xc' Xc / T, shrunk to (1-d) S_b + d m I (Ledoit-Wolf, SURVEY A.3 / sklearn _shrunk_covariance.py)
//
// HBM/L2 traffic per simulation: X read twice (mean pass + SYRK; 2 x 8TN bytes, three times for Ledoit-Wolf), cov
// written once (8N^2), plus one read-modify-write of cov for the Ledoit-Wolf shrink.  Flops: 2 T N^2 / 2 (lower tiles).
#include "common.cuh"
#include "dmma.cuh"

#define CM_COLS 32
#define CM_ROWGROUPS 8

// shift != nullptr: X holds draws minus `shift`; ybar (the mean of what is stored) goes to `mean`, shift + ybar to `mean_out`
__global__ void __launch_bounds__(CM_COLS* CM_ROWGROUPS) colmean_kernel(const double* __restrict__ X, int T, int N,
                                                                        double* __restrict__ mean,
                                                                        const double* __restrict__ shift,
                                                                        double* __restrict__ mean_out) {
    __shared__ double part[CM_ROWGROUPS][CM_COLS];
    const int s = blockIdx.y;
    const int col = blockIdx.x * CM_COLS + (threadIdx.x % CM_COLS);
    const int rg = threadIdx.x / CM_COLS;
    const double* x = X + (size_t)s * T * N;
    double acc = 0.0;
    if (col < N)
        for (int t = rg; t < T; t += CM_ROWGROUPS) acc += x[(size_t)t * N + col];
    part[rg][threadIdx.x % CM_COLS] = acc;
    __syncthreads();
    if (rg == 0 && col < N) {
        double tot = 0.0;
#pragma unroll
        for (int r = 0; r < CM_ROWGROUPS; ++r) tot += part[r][threadIdx.x];
        const double m = tot / (double)T;
        mean[(size_t)s * N + col] = m;
        if (shift) mean_out[(size_t)s * N + col] = shift[col] + m;
    }
}

// Lower-triangle tile p -> (I, J), I >= J
__device__ __forceinline__ void tile_ij(int p, int& I, int& J) {
    int i = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= p) ++i;
    while (i * (i + 1) / 2 > p) --i;
    I = i;
    J = p - i * (i + 1) / 2;
}

// grid = (2 x lower 128 x 128 tiles, S), 128 threads (4 warps as 2 x 2, each 64 x 32), two CTAs per SM so that the
// prologue / epilogue of one overlaps the main loop of the other.  An off-diagonal tile is split into its two 128 x 64
// column halves; a diagonal tile into the two warp sets of dmma_slab_diag (sub-blocks of the lower triangle: 4 x 16 and
// 2 x 16 + 2 x 20 DMMA per k-step).  lwpart[s][cta][2] receives (sum of squares of the scaled entries this CTA stands
// for in the full symmetric matrix, trace contribution).
// VEC: 16-byte cp.async (N even and X 16-byte aligned), otherwise 8-byte copies.
// CENTER = false: the data are already (nearly) centred -- the fused loop hands over Y = X - true mu -- and the products are
// taken raw, cov = (Y'Y - T ybar ybar') scale in the epilogue: the per-fragment subtractions of the centred form are DADDs on
// the pipe the DMMAs run on, 20 % of the kernel in the ncu source view (profiles/kernels_r02.md).
#define SY_THREADS 128
template <bool VEC, bool CENTER>
__global__ void __launch_bounds__(SY_THREADS, 2)
syrk_kernel(const double* __restrict__ X, const double* __restrict__ mean, int T, int N, double scale,
            double* __restrict__ cov, double* __restrict__ lwpart) {
    extern __shared__ double sm[];
    __shared__ double red[32];
#ifdef SYRK_TIMING
    const long long tk0 = clock64();
#endif
    const int s = blockIdx.y;
    int I, J;
    tile_ij(blockIdx.x >> 1, I, J);
    const int half = blockIdx.x & 1;
    const bool diag = (I == J);
    const double* x = X + (size_t)s * T * N;
    const double* mu = mean + (size_t)s * N;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int wm0 = (wid >> 1) * DM_WM, wn0 = (wid & 1) * DM_WN;    // off-diagonal: within the 128 x 64 half
    const int jcol0 = J * DM_BN + half * (DM_BN / 2);                 // first column of this CTA's B operand
    const int dwid = half * 4 + wid;                                  // diagonal: warp id of dmma_slab_diag's dealing
    // loader mapping (A: 128 columns, B: 64): VEC: column pair fixed, k = lk + LKS r; otherwise column fixed
    constexpr int LW = VEC ? 2 : 1, LKS = SY_THREADS * LW / DM_BM;
    const int lcol = (tid % (DM_BM / LW)) * LW, lk = tid / (DM_BM / LW);
    const int acol = I * DM_BM + lcol;
    const bool aok = acol < N;
    const double* asrc = x + (aok ? acol : 0);
    constexpr int LKSB = SY_THREADS * LW / (DM_BN / 2);
    const int lcolb = (tid % (DM_BN / 2 / LW)) * LW, lkb = tid / (DM_BN / 2 / LW);
    const int bcol = jcol0 + lcolb;
    const bool bok = bcol < N;
    const double* bsrc = x + (bok ? bcol : 0);

    // per-fragment column means (centring happens in registers, the tiles are copied raw).  Diagonal tiles use the
    // sub-tile decomposition of dmma_slab_diag: means of sub-tile U at [4 U .. 4 U + 3]
    double amean[DM_MB], bmean[DM_MB];
    int r0a = 0, c0a = 0, r0b = 0, c0b = 0;
    if (diag) dmma_diag_subtiles(dwid, r0a, c0a, r0b, c0b);
    {
        const int g = lane >> 2;
#pragma unroll
        for (int mb = 0; mb < DM_MB; ++mb) {
            int c = I * DM_BM + (diag ? (mb < 4 ? r0a : r0b - 32) : wm0) + mb * 8 + g;
            amean[mb] = (c < N) ? mu[c] : 0.0;
        }
#pragma unroll
        for (int nb = 0; nb < DM_MB; ++nb) {
            int c = (diag ? J * DM_BN + (nb < 4 ? c0a : c0b - 32) : jcol0 + wn0) + nb * 8 + g;
            bmean[nb] = (c < N && (diag || nb < DM_NB)) ? mu[c] : 0.0;
        }
    }
    double acc[DM_MB][DM_NB][2];
#pragma unroll
    for (int a = 0; a < DM_MB; ++a)
#pragma unroll
        for (int b = 0; b < DM_NB; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    auto issue = [&](int kt, int buf) {
#if defined(DM_ABLATE) && (DM_ABLATE & 4)   // timing experiment only: no global -> shared copies
        return;
#endif
        double* At = sm + (size_t)buf * 2 * DM_TILE_KMAJOR;
        double* Bt = At + DM_TILE_KMAJOR;
        const int kbase = kt * DM_BK;
#pragma unroll
        for (int r = 0; r < DM_BK / LKS; ++r) {
            const int kk = lk + LKS * r, k = kbase + kk;
            const bool ok = k < T;
            const size_t goff = (size_t)(ok ? k : 0) * N;
            if (VEC) cp_async16(&At[kk * DM_LD_KMAJOR + lcol], asrc + goff, ok && aok);
            else cp_async8(&At[kk * DM_LD_KMAJOR + lcol], asrc + goff, ok && aok);
        }
        if (!diag) {
#pragma unroll
            for (int r = 0; r < DM_BK / LKSB; ++r) {
                const int kk = lkb + LKSB * r, k = kbase + kk;
                const bool ok = k < T;
                const size_t goff = (size_t)(ok ? k : 0) * N;
                if (VEC) cp_async16(&Bt[kk * DM_LD_KMAJOR + lcolb], bsrc + goff, ok && bok);
                else cp_async8(&Bt[kk * DM_LD_KMAJOR + lcolb], bsrc + goff, ok && bok);
            }
        }
    };
    const int nkt = (T + DM_BK - 1) / DM_BK;
#ifdef SYRK_TIMING
    const long long tk1 = clock64();
    long long tk2 = 0;
#endif
#pragma unroll
    for (int st = 0; st < DM_STAGES - 1; ++st) {
        if (st < nkt) issue(st, st);
        cp_async_commit();
    }
    for (int kt = 0; kt < nkt; ++kt) {
        cp_async_wait<DM_STAGES - 2>();
        __syncthreads();
#ifdef SYRK_TIMING
        if (kt == 0) tk2 = clock64();
#endif
        const double* At = sm + (size_t)(kt % DM_STAGES) * 2 * DM_TILE_KMAJOR;
        const double* Bt = diag ? At : At + DM_TILE_KMAJOR;
        // the stage kt + 2 overwrites the buffer of slab kt - 1, which every warp has left (barrier above)
        auto next = [&] {
            if (kt + DM_STAGES - 1 < nkt) issue(kt + DM_STAGES - 1, (kt + DM_STAGES - 1) % DM_STAGES);
            cp_async_commit();
        };
        if (diag) dmma_slab_diag<true, CENTER>(At, dwid, lane, acc, next, amean, bmean, T - kt * DM_BK);
        else dmma_slab<true, true, CENTER>(At, Bt, wm0, wn0, lane, acc, next, amean, bmean, T - kt * DM_BK);
    }

#ifdef SYRK_TIMING
    const long long tk3 = clock64();
#endif
    // epilogue: scale, write the lower part and its mirror, accumulate Frobenius / trace partials
    const int g = lane >> 2, t4 = lane & 3;
    double* c = cov + (size_t)s * N * N;
    double frob = 0.0, tr = 0.0;
    const bool two = diag && dwid >= 6;
#pragma unroll
    for (int mb = 0; mb < DM_MB; ++mb)
#pragma unroll
        for (int nb = 0; nb < DM_NB; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                // diagonal tiles: accumulator rows 0..3 are sub-tile a, rows 4..7 sub-tile b (warps 6, 7 only)
                int row = I * DM_BM + (diag ? (mb < 4 ? r0a + mb * 8 : r0b + (mb - 4) * 8) : wm0 + mb * 8) + g;
                int col = (diag ? J * DM_BN + (mb < 4 ? c0a : c0b) : jcol0 + wn0) + nb * 8 + 2 * t4 + e;
                if (row < N && col < N && col <= row && (!diag || mb < 4 || two)) {
                    double v = CENTER ? acc[mb][nb][e] * scale
                                      : fma(-(double)T * mu[row], mu[col], acc[mb][nb][e]) * scale;
                    c[(size_t)row * N + col] = v;
                    if (col != row) {
                        c[(size_t)col * N + row] = v;
                        frob += 2.0 * v * v;
                    } else {
                        frob += v * v;
                        tr += v;
                    }
                }
            }
    if (lwpart) {
        frob = block_sum(frob, red);
        tr = block_sum(tr, red);
        if (tid == 0) {
            size_t o = ((size_t)s * gridDim.x + blockIdx.x) * 2;
            lwpart[o] = frob;
            lwpart[o + 1] = tr;
        }
    }
#ifdef SYRK_TIMING
    if (tid == 0 && s == 1 && blockIdx.x < 20)
        printf("syrk tile %d (I %d J %d): setup %lld fill %lld mainloop %lld epilogue %lld cycles\n", blockIdx.x, I, J,
               tk1 - tk0, tk2 - tk1, tk3 - tk2, clock64() - tk3);
#endif
}

// Ledoit-Wolf beta_raw = sum_t (sum_i xc_ti^2)^2, partial per chunk of rows.  grid = (row chunks, S)
#define RN_ROWS 64
__global__ void __launch_bounds__(256) lw_rownorm_kernel(const double* __restrict__ X,
                                                         const double* __restrict__ mean, int T, int N,
                                                         double* __restrict__ part) {
    __shared__ double red[32];
    const int s = blockIdx.y;
    const double* x = X + (size_t)s * T * N;
    const double* mu = mean + (size_t)s * N;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double acc = 0.0;
    for (int r = wid; r < RN_ROWS; r += 8) {
        int t = blockIdx.x * RN_ROWS + r;
        if (t >= T) break;
        double q = 0.0;
        for (int j = lane; j < N; j += 32) {
            double d = x[(size_t)t * N + j] - mu[j];
            q = fma(d, d, q);
        }
        q = warp_sum(q);
        acc += q * q;  // identical on all lanes
    }
    double tot = block_sum(lane == 0 ? acc : 0.0, red);
    if (threadIdx.x == 0) part[(size_t)s * gridDim.x + blockIdx.x] = tot;
}

// Finishes Ledoit-Wolf: shrinkage from the partials, then cov <- (1-d) cov + d m I in place.  grid = (S, 8)
__global__ void __launch_bounds__(256) lw_finish_kernel(double* __restrict__ cov, const double* __restrict__ lwpart,
                                                        int ntiles, const double* __restrict__ rnpart, int nchunks,
                                                        int T, int N, double* __restrict__ shrink_out) {
    __shared__ double s_shrink, s_m;
    const int s = blockIdx.x;
    if (threadIdx.x == 0) {   // every CTA of the matrix repeats this small sum in the same order
        double delta_raw = 0.0, tr = 0.0, beta_raw = 0.0;
        for (int i = 0; i < ntiles; ++i) {
            delta_raw += lwpart[((size_t)s * ntiles + i) * 2];
            tr += lwpart[((size_t)s * ntiles + i) * 2 + 1];
        }
        for (int i = 0; i < nchunks; ++i) beta_raw += rnpart[(size_t)s * nchunks + i];
        double m = tr / N;
        double beta = (beta_raw / T - delta_raw) / ((double)N * T);
        double delta = (delta_raw - N * m * m) / N;
        beta = fmin(beta, delta);
        double sh = (beta == 0.0) ? 0.0 : beta / delta;
        s_shrink = sh;
        s_m = m;
        if (shrink_out && blockIdx.y == 0) shrink_out[s] = sh;
    }
    __syncthreads();
    const double sh = s_shrink, m = s_m;
    double* c = cov + (size_t)s * N * N;
    // rows blockIdx.y, blockIdx.y + gridDim.y, ...: one warp per row (a single CTA per matrix streamed 4 MB through
    // 256 threads with a 64-bit div/mod per element and took 0.29 ms per 128 matrices)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = blockIdx.y * nw + wid; i < N; i += gridDim.y * nw) {
        double* row = c + (size_t)i * N;
        for (int j = lane; j < N; j += 32) {
            row[j] = lw_shrink(row[j], sh, m, j == i);
        }
    }
}

// The Ledoit-Wolf coefficients alone (same sums, same order as lw_finish_kernel): (shrinkage, mean variance) per
// simulation, for consumers that apply the shrink themselves (corr_kernel of the de-noiser).  One thread per simulation.
__global__ void lw_coef_kernel(const double* __restrict__ lwpart, int ntiles, const double* __restrict__ rnpart,
                               int nchunks, int T, int N, int S, double* __restrict__ coef) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double delta_raw = 0.0, tr = 0.0, beta_raw = 0.0;
    for (int i = 0; i < ntiles; ++i) {
        delta_raw += lwpart[((size_t)s * ntiles + i) * 2];
        tr += lwpart[((size_t)s * ntiles + i) * 2 + 1];
    }
    for (int i = 0; i < nchunks; ++i) beta_raw += rnpart[(size_t)s * nchunks + i];
    double m = tr / N;
    double beta = (beta_raw / T - delta_raw) / ((double)N * T);
    double delta = (delta_raw - N * m * m) / N;
    beta = fmin(beta, delta);
    coef[2 * (size_t)s] = (beta == 0.0) ? 0.0 : beta / delta;
    coef[2 * (size_t)s + 1] = m;
}

__global__ void fill_kernel(double* p, size_t n, double v) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

int mcosb_dev_moments(mcosb_ctx* ctx, int S, const double* dX, int kind, double* dmu, double* dcov, double* dshrink,
                      double* lwcoef_defer, const double* shift, const double* ybar_pre) {
    if (kind < 0 || kind > 2) return mcosb_fail(ctx, MCOSB_ERR_ARG, "unknown moments kind");
    if (S == 0) return MCOSB_OK;
    const int N = ctx->N, T = ctx->T;
    if (kind == MCOSB_MOMENTS_JACKKNIFE && T < 3)
        return mcosb_fail(ctx, MCOSB_ERR_ARG, "jackknife needs at least 3 observations");
    if (kind == MCOSB_MOMENTS_MUCOV && T < 2) return mcosb_fail(ctx, MCOSB_ERR_ARG, "covariance needs T >= 2");
    const bool lw = (kind == MCOSB_MOMENTS_LEDOIT_WOLF);

    // with a shift the statistics are taken of Y = X - shift: ybar in scratch, shift + ybar reported as the mean
    double* dmean = dmu;
    if (shift && ybar_pre) {
        dmean = const_cast<double*>(ybar_pre);      // the sampler summed the columns on the way: no mean pass over X
    } else {
        if (shift) MCOSB_TRY(mcosb_scratch_t(ctx, SL_COLMEAN, (size_t)S * N, &dmean));
        dim3 gm((N + CM_COLS - 1) / CM_COLS, S);
        colmean_kernel<<<gm, CM_COLS * CM_ROWGROUPS, 0, ctx->stream>>>(dX, T, N, dmean, shift, dmu);
        MCOSB_LAUNCHED(ctx, MK_COLMEAN);
    }

    const int nt = (N + DM_BM - 1) / DM_BM, ntiles = 2 * (nt * (nt + 1) / 2);   // two CTAs per 128 x 128 tile
    double* lwpart = nullptr;
    if (lw) MCOSB_TRY(mcosb_scratch_t(ctx, SL_LWPART, (size_t)S * ntiles * 2, &lwpart));
    const size_t smem = (size_t)DM_STAGES * 2 * DM_TILE_KMAJOR * sizeof(double);
    const double scale = lw ? 1.0 / T : 1.0 / (T - 1);
    const bool vec = N % 2 == 0 && (uintptr_t)dX % 16 == 0;
#define SY_LAUNCH(V, C)                                                                                               \
    do {                                                                                                              \
        MCOSB_CUDA(ctx, cudaFuncSetAttribute(syrk_kernel<V, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        syrk_kernel<V, C><<<dim3(ntiles, S), SY_THREADS, smem, ctx->stream>>>(dX, dmean, T, N, scale, dcov, lwpart);    \
    } while (0)
    if (vec && !shift) SY_LAUNCH(true, true);
    else if (vec) SY_LAUNCH(true, false);
    else if (!shift) SY_LAUNCH(false, true);
    else SY_LAUNCH(false, false);
#undef SY_LAUNCH
    MCOSB_LAUNCHED(ctx, MK_SYRK);

    if (lw) {
        const int nchunks = (T + RN_ROWS - 1) / RN_ROWS;
        double* rnpart = nullptr;
        MCOSB_TRY(mcosb_scratch_t(ctx, SL_LWSTAT, (size_t)S * nchunks, &rnpart));
        lw_rownorm_kernel<<<dim3(nchunks, S), 256, 0, ctx->stream>>>(dX, dmean, T, N, rnpart);
        MCOSB_LAUNCHED(ctx, MK_LW_ROWNORM);
        if (lwcoef_defer) {
            lw_coef_kernel<<<(S + 127) / 128, 128, 0, ctx->stream>>>(lwpart, ntiles, rnpart, nchunks, T, N, S, lwcoef_defer);
        } else {
            lw_finish_kernel<<<dim3(S, 8), 256, 0, ctx->stream>>>(dcov, lwpart, ntiles, rnpart, nchunks, T, N, dshrink);
        }
        MCOSB_LAUNCHED(ctx, MK_LW_FINISH);
    } else if (dshrink) {
        fill_kernel<<<(S + 255) / 256, 256, 0, ctx->stream>>>(dshrink, (size_t)S, 0.0);
        MCOSB_LAUNCHED(ctx, MK_FILL);
    }
    return MCOSB_OK;
}

extern "C" int mcosb_moments(mcosb_ctx* ctx, int S, const double* X, int kind, double* mu_hat, double* cov_hat,
                             double* shrink) {
    MCOSB_CHECK_CTX(ctx);
    if (S < 0 || !X || !mu_hat || !cov_hat) return mcosb_fail(ctx, MCOSB_ERR_ARG, "bad argument to mcosb_moments");
    const size_t N = ctx->N, T = ctx->T;
    Staging st(ctx);
    const double* dX;
    double *dmu, *dcov, *dsh;
    MCOSB_TRY(st.in(X, (size_t)S * T * N, SL_STAGE_IN0, &dX));
    MCOSB_TRY(st.out(mu_hat, (size_t)S * N, SL_STAGE_OUT0, &dmu));
    MCOSB_TRY(st.out(cov_hat, (size_t)S * N * N, SL_STAGE_OUT1, &dcov));
    MCOSB_TRY(st.out(shrink, (size_t)S, SL_STAGE_OUT2, &dsh));
    MCOSB_TRY(mcosb_dev_moments(ctx, S, dX, kind, dmu, dcov, dsh));
    return st.finish();
}
