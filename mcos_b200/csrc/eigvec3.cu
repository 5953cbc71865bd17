// D5 (wide): signal eigenvectors + clipped reconstruction, parallel ACROSS simulations instead of inside one CTA.
//
// eigvec_kernel (denoise.cu) does everything for one matrix in one CTA: its inverse iteration keeps ~10 lanes busy for
// thousands of dependent steps while the other 500 threads wait (12 us per matrix at N = 500).  A register-resident
// one-warp-per-vector variant was tried and was slower (23 us): the critical path is FP64 division / dependent-issue
// latency, which only more independent work hides.  So the stage is cut into three wide kernels:
//   inviter_kernel        one THREAD per (simulation, eigenvector): pivoted LU of (T - lam I) and three inverse-iteration
//                         sweeps; the per-thread arrays live in global scratch interleaved across threads (coalesced),
//                         all 32 lanes of a warp busy with different (simulation, vector) pairs;
//   backtransform_kernel  one WARP per (simulation, eigenvector): normalise, apply the stored Householder reflectors
//                         (iterate in registers, reflector rows read coalesced), write the eigenvector;
//   reconstruct_kernel    elementwise: cov_ij = clip((lbar d_ij + sum_s (lam_s - lbar) v_si v_sj) / sqrt(r_i r_j)) s_i s_j.
// Simulations with more than EV3_VMAX signal eigenvalues are left to eigvec_kernel (it skips the others).
#include "common.cuh"
#include "sbr.cuh"

#define EV3_VMAX 16

__device__ __forceinline__ double ev3_rcp(double x) {   // 1/x to ~1 ulp: hardware approximation + two Newton steps
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}

// W layout: W[a][i][g], a in 0..4 (dl, inverse pivot, du, du2, y), g = sim * EV3_VMAX + vec, stride G = S * EV3_VMAX
__global__ void __launch_bounds__(64) inviter_kernel(const double* __restrict__ dall, const double* __restrict__ eall,
                                                     const double* __restrict__ lamall, const int* __restrict__ nfall,
                                                     int N, int S, double* __restrict__ W,
                                                     unsigned char* __restrict__ piv) {
    const size_t G = (size_t)S * EV3_VMAX;
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g >= G) return;
    const int sim = (int)(g / EV3_VMAX), vec = (int)(g % EV3_VMAX);
    const int nf = nfall[sim];
    if (nf > EV3_VMAX || vec >= nf) return;
    const double* d = dall + (size_t)sim * N;
    const double* e = eall + (size_t)sim * N;
    const double* lam = lamall + (size_t)sim * N;
    double tnorm = 0.0;
    for (int i = 0; i < N; ++i) tnorm = fmax(tnorm, fabs(d[i]) + 2.0 * fabs(e[i]));
    const double tiny_piv = 2.2e-16 * fmax(tnorm, 1e-300);
    double lj = lam[vec];
    if (vec > 0 && lam[vec - 1] - lj < 10.0 * tiny_piv) lj -= 10.0 * tiny_piv * (double)(vec % 7 + 1);
    double* dl = W + 0 * (size_t)N * G + g;
    double* ip = W + 1 * (size_t)N * G + g;
    double* du = W + 2 * (size_t)N * G + g;
    double* du2 = W + 3 * (size_t)N * G + g;
    double* y = W + 4 * (size_t)N * G + g;
    unsigned char* pv = piv + g;
    // LU with partial pivoting (dgttrf recurrences)
    double dcur = d[0] - lj, ucur = (N > 1) ? e[0] : 0.0;
    for (int i = 0; i + 1 < N; ++i) {
        const double sub = e[i];
        const double dnext = d[i + 1] - lj;
        const double unext = (i + 2 < N) ? e[i + 1] : 0.0;
        if (fabs(dcur) >= fabs(sub)) {
            if (dcur == 0.0) dcur = tiny_piv;
            const double inv = ev3_rcp(dcur);
            const double fact = sub * inv;
            dl[i * G] = fact; ip[i * G] = inv; du[i * G] = ucur; du2[i * G] = 0.0; pv[i * G] = 0;
            dcur = dnext - fact * ucur;
            ucur = unext;
        } else {
            const double inv = ev3_rcp(sub);
            const double fact = dcur * inv;
            dl[i * G] = fact; ip[i * G] = inv; du[i * G] = dnext; du2[i * G] = unext; pv[i * G] = 1;
            dcur = ucur - fact * dnext;
            ucur = -fact * unext;
        }
    }
    if (dcur == 0.0) dcur = tiny_piv;
    ip[(size_t)(N - 1) * G] = ev3_rcp(dcur);
    // start vector: deterministic pseudo-random in (-1, 1)
    for (int i = 0; i < N; ++i) {
        uint32_t h = (uint32_t)(i * 2654435761u) ^ (uint32_t)((vec + 1) * 40503u);
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        y[i * G] = ((double)(h & 0xffffff) + 0.5) / 8388608.0 - 1.0;
    }
    double scale = 1.0;
    for (int sweep = 0; sweep < 3; ++sweep) {
        double bi = y[0] * scale;
        for (int i = 0; i + 1 < N; ++i) {
            const double bn = y[(size_t)(i + 1) * G] * scale;
            const double f = dl[i * G];
            if (pv[i * G] == 0) {
                y[i * G] = bi;
                bi = bn - f * bi;
            } else {
                y[i * G] = bn;
                bi = bi - f * bn;
            }
        }
        double x1 = bi * ip[(size_t)(N - 1) * G], x2 = 0.0, mx = fabs(x1);
        y[(size_t)(N - 1) * G] = x1;
        for (int i = N - 2; i >= 0; --i) {
            const double xi = (y[i * G] - du[i * G] * x1 - du2[i * G] * x2) * ip[i * G];
            y[i * G] = xi;
            x2 = x1;
            x1 = xi;
            mx = fmax(mx, fabs(xi));
        }
        scale = 1.0 / mx;
    }
    // the last sweep's overall scale is irrelevant: backtransform_kernel normalises
}

// One warp per (simulation, eigenvector), MAXM = ceil(N / 32) register slots per lane; a CTA holds 8 vectors of ONE
// simulation (two CTAs per simulation), so the reflector rows are staged in shared memory once per CTA, one reflector
// ahead of the one being applied (every warp used to read its own copy from L2 with the latency exposed at each step).
template <int MAXM>
__global__ void __launch_bounds__(256) backtransform_kernel(const double* __restrict__ Aall,
                                                            const double* __restrict__ tauall,
                                                            const int* __restrict__ nfall, int N, int S,
                                                            const double* __restrict__ W, double* __restrict__ V,
                                                            int* __restrict__ status, int off) {
    extern __shared__ double rowbuf[];           // [2][N]
    constexpr int PFN = (MAXM * 32 + 255) / 256;
    const size_t G = (size_t)S * EV3_VMAX;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int sim = blockIdx.x / (EV3_VMAX / 8), vec = (blockIdx.x % (EV3_VMAX / 8)) * 8 + wid;
    const size_t g = (size_t)sim * EV3_VMAX + vec;
    const int nf = nfall[sim];
    if (nf > EV3_VMAX || (vec - wid) >= nf) return;      // nothing for this CTA (uniform)
    const bool active = vec < nf;
    const double* A = Aall + (size_t)sim * N * N;
    const double* tau = tauall + (size_t)sim * N;
    const double* y = W + 4 * (size_t)N * G + g;
    double yr[MAXM];
    if (active) {
        double mx = 0.0;
#pragma unroll
        for (int m = 0; m < MAXM; ++m) {
            const int i = lane + 32 * m;
            yr[m] = (i < N) ? y[(size_t)i * G] : 0.0;
            mx = fmax(mx, fabs(yr[m]));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const double pre = 1.0 / mx;
        double nrm = 0.0;
#pragma unroll
        for (int m = 0; m < MAXM; ++m) { yr[m] *= pre; nrm = fma(yr[m], yr[m], nrm); }
        nrm = warp_sum(nrm);
        const double sc = 1.0 / sqrt(nrm);
        if (!isfinite(sc) || !isfinite(pre)) {
            if (lane == 0 && status) atomicOr(&status[sim], MCOSB_ST_EIGVEC);
        }
#pragma unroll
        for (int m = 0; m < MAXM; ++m) yr[m] *= sc;
    }
    // y <- H_0 ... H_{N-2-off} y ; reflector k: u[k+off] = 1, u[j] = A[k][j] for j > k+off
    // (off = 1: one-stage tridiagonalisation; off = 8: panel reflectors of sy2sb_kernel)
    const int K0 = N - 2 - off;
    if (K0 >= 0)
        for (int j = tid; j < N; j += 256) rowbuf[j] = (j > K0 + off) ? A[(size_t)K0 * N + j] : 0.0;
    __syncthreads();
    for (int k = K0; k >= 0; --k) {
        const int cur = (K0 - k) & 1;
        double pf[PFN];
#pragma unroll
        for (int r = 0; r < PFN; ++r) {
            const int j = tid + 256 * r;
            pf[r] = (k > 0 && j < N && j > k - 1 + off) ? A[(size_t)(k - 1) * N + j] : 0.0;
        }
        const double tk = tau[k];
        if (active && tk != 0.0) {
            const double* u = rowbuf + (size_t)cur * N;
            double ur[MAXM];
            double dp[4] = {0.0, 0.0, 0.0, 0.0};     // four chains: a dependent FP64 FMA costs ~25 cycles
#pragma unroll
            for (int m = 0; m < MAXM; ++m) {
                const int j = lane + 32 * m;
                ur[m] = (j > k + off && j < N) ? u[j] : ((j == k + off) ? 1.0 : 0.0);
                dp[m & 3] = fma(ur[m], yr[m], dp[m & 3]);
            }
            double dot = warp_sum((dp[0] + dp[1]) + (dp[2] + dp[3])) * tk;
#pragma unroll
            for (int m = 0; m < MAXM; ++m) yr[m] = fma(-dot, ur[m], yr[m]);
        }
        double* nb = rowbuf + (size_t)(cur ^ 1) * N;
#pragma unroll
        for (int r = 0; r < PFN; ++r) {
            const int j = tid + 256 * r;
            if (j < N) nb[j] = pf[r];
        }
        __syncthreads();
    }
    if (active) {
        double* out = V + g * (size_t)N;
#pragma unroll
        for (int m = 0; m < MAXM; ++m) {
            const int i = lane + 32 * m;
            if (i < N) out[i] = yr[m];
        }
    }
}

// grid = (ceil(N / 128), S), 256 threads: the CTA stages the signal eigenvectors in shared memory once (they are re-read
// N / 32 times per simulation instead of N / 8 from L2), each warp owns 4 output rows and a lane walks the columns, so one
// shared-memory load of v_sj feeds 4 FMAs.  r_i = lbar + sum_s c_s v_si^2 is the diagonal of the clipped matrix.
#define RC_ROWS 2
#ifndef RC_CTA_ROWS
#define RC_CTA_ROWS 128   // rows per CTA: the staging of the eigenvectors is amortised over 128 x N outputs
#endif
__global__ void __launch_bounds__(256) reconstruct_kernel(const double* __restrict__ V, const double* __restrict__ lamall,
                                                          const int* __restrict__ nfall,
                                                          const double* __restrict__ sigmaall, int N,
                                                          double* __restrict__ covall) {
    extern __shared__ double sm[];
    double* rs = sm;             // [N] 1 / sqrt(r_j)
    double* sg = sm + N;         // [N] sigma_j
    double* coef = sg + N;       // [EV3_VMAX]
    double* Vsm = coef + EV3_VMAX;   // [nf][N]
    __shared__ double red[32];
    const int sim = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nf = nfall[sim];
    if (nf > EV3_VMAX) return;   // left to eigvec_kernel
    const double* lam = lamall + (size_t)sim * N;
    const double* Vs = V + (size_t)sim * EV3_VMAX * N;
    double part = 0.0;
    for (int i = nf + tid; i < N; i += 256) part += lam[i];
    for (int idx = tid; idx < nf * N; idx += 256) Vsm[idx] = Vs[idx];
    for (int j = tid; j < N; j += 256) sg[j] = sigmaall[(size_t)sim * N + j];
    const double lbar = (nf < N) ? block_sum(part, red) / (double)(N - nf) : 0.0;
    if (tid < EV3_VMAX) coef[tid] = (tid < nf) ? lam[tid] - lbar : 0.0;
    __syncthreads();
    for (int j = tid; j < N; j += 256) {
        double r = lbar;
        for (int s = 0; s < nf; ++s) { const double v = Vsm[s * N + j]; r = fma(coef[s] * v, v, r); }
        rs[j] = 1.0 / sqrt(r);     // the division of the renormalisation becomes two multiplications per element
    }
    __syncthreads();
    double* out = covall + (size_t)sim * N * N;
    for (int rg = wid; rg < RC_CTA_ROWS / RC_ROWS; rg += 8) {
        const int i0 = blockIdx.x * RC_CTA_ROWS + rg * RC_ROWS;
        if (i0 >= N) break;
        double a[RC_ROWS][EV3_VMAX], rsi[RC_ROWS], si[RC_ROWS];
#pragma unroll
        for (int r = 0; r < RC_ROWS; ++r) {
            const int i = min(i0 + r, N - 1);
            rsi[r] = rs[i];
            si[r] = sg[i];
#pragma unroll
            for (int s = 0; s < EV3_VMAX; ++s) a[r][s] = (s < nf) ? coef[s] * Vsm[s * N + i] : 0.0;
        }
        for (int j = lane; j < N; j += 32) {
            double acc[RC_ROWS];
#pragma unroll
            for (int r = 0; r < RC_ROWS; ++r) acc[r] = (i0 + r == j) ? lbar : 0.0;
#pragma unroll
            for (int s = 0; s < EV3_VMAX; ++s) {
                if (s < nf) {
                    const double v = Vsm[s * N + j];
#pragma unroll
                    for (int r = 0; r < RC_ROWS; ++r) acc[r] = fma(a[r][s], v, acc[r]);
                }
            }
            const double rsj = rs[j], sj = sg[j];
#pragma unroll
            for (int r = 0; r < RC_ROWS; ++r) {
                if (i0 + r < N) {
                    double c = acc[r] * (rsi[r] * rsj);
                    c = c < -1.0 ? -1.0 : (c > 1.0 ? 1.0 : c);
                    out[(size_t)(i0 + r) * N + j] = c * (si[r] * sj);
                }
            }
        }
    }
}

// Launches the three kernels. Returns 1 if used (simulations with nf > EV3_VMAX still need eigvec_kernel), 0 if N is
// out of range for the register back-transformation, < 0 on error.
int mcosb_launch_eigvec3(mcosb_ctx* ctx, int S, const double* A, const double* d, const double* e, const double* tau,
                         const double* lam, const int* nf, const double* sigma, double* cov, int* status,
                         const double* Q2) {
    const int N = ctx->N;
    const size_t n = (size_t)N;
    if (N > 1024) return 0;
    const size_t G = (size_t)S * EV3_VMAX;
    double *W, *V;
    unsigned char* piv;
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_WORK, 5 * n * G, &W));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_VEC, n * G, &piv));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_VAR, n * G, &V));
    inviter_kernel<<<(unsigned)((G + 63) / 64), 64, 0, ctx->stream>>>(d, e, lam, nf, N, S, W, piv);
    MCOSB_LAUNCHED(ctx, MK_INVITER);
    const int off = Q2 ? SBR_B : 1;
    if (Q2) MCOSB_TRY(mcosb_launch_q2back(ctx, S, Q2, nf, W + 4 * n * G) == 1 ? MCOSB_OK : MCOSB_ERR_STATE);
    const unsigned bt_grid = (unsigned)((G + 7) / 8);
    if (N <= 128) backtransform_kernel<4><<<bt_grid, 256, 2 * n * sizeof(double), ctx->stream>>>(A, tau, nf, N, S, W, V, status, off);
    else if (N <= 256) backtransform_kernel<8><<<bt_grid, 256, 2 * n * sizeof(double), ctx->stream>>>(A, tau, nf, N, S, W, V, status, off);
    else if (N <= 512) backtransform_kernel<16><<<bt_grid, 256, 2 * n * sizeof(double), ctx->stream>>>(A, tau, nf, N, S, W, V, status, off);
    else if (N <= 768) backtransform_kernel<24><<<bt_grid, 256, 2 * n * sizeof(double), ctx->stream>>>(A, tau, nf, N, S, W, V, status, off);
    else backtransform_kernel<32><<<bt_grid, 256, 2 * n * sizeof(double), ctx->stream>>>(A, tau, nf, N, S, W, V, status, off);
    MCOSB_LAUNCHED(ctx, MK_BACKTRANSFORM);
    const size_t smem = (2 * n + EV3_VMAX + EV3_VMAX * n) * sizeof(double);
    MCOSB_CUDA(ctx, cudaFuncSetAttribute(reconstruct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    reconstruct_kernel<<<dim3((N + RC_CTA_ROWS - 1) / RC_CTA_ROWS, S), 256, smem, ctx->stream>>>(V, lam, nf, sigma, N, cov);
    MCOSB_LAUNCHED(ctx, MK_RECONSTRUCT);
    return 1;
}

// ---------------------------------------------------------------------------------------------------------------
// DetoneCovarianceTransformer.transform (covariance_transformer.py:222-250): remove the n_remove largest
// eigen-components of the correlation matrix, renormalise to unit diagonal, rescale by the input standard deviations.
// Reuses the eigen front end and the wide eigenvector kernels with n_facts := n_remove for every simulation.
// (The reference sorts eigenvalues by absolute value; for a positive semi-definite correlation matrix that is the same
// order.  corr_ii is taken as exactly 1; the reference's value is within one ulp of it.)
// ---------------------------------------------------------------------------------------------------------------
__global__ void fill_int_kernel(int* p, size_t n, int v) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// grid = (ceil(N / 8), S), one warp per row, in place on cov (each element is read before it is overwritten)
__global__ void __launch_bounds__(256) detone_kernel(const double* __restrict__ V, const double* __restrict__ lamall,
                                                     int n_remove, const double* __restrict__ sigmaall, int N,
                                                     double* __restrict__ covall) {
    extern __shared__ double sm[];
    double* rs = sm;             // [N] sqrt(diag(c2))
    const int sim = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double* lam = lamall + (size_t)sim * N;
    const double* sigma = sigmaall + (size_t)sim * N;
    const double* Vs = V + (size_t)sim * EV3_VMAX * N;
    for (int j = tid; j < N; j += 256) {
        double r = 1.0;
        for (int s = 0; s < n_remove; ++s) { const double v = Vs[(size_t)s * N + j]; r = fma(-lam[s] * v, v, r); }
        rs[j] = sqrt(r);
    }
    __syncthreads();
    const int i = blockIdx.x * 8 + wid;
    if (i >= N) return;
    double vi[EV3_VMAX];
#pragma unroll
    for (int s = 0; s < EV3_VMAX; ++s) vi[s] = (s < n_remove) ? lam[s] * Vs[(size_t)s * N + i] : 0.0;
    const double rsi = rs[i], si = sigma[i];
    double* row = covall + (size_t)sim * N * N + (size_t)i * N;
    for (int j = lane; j < N; j += 32) {
        const double sj = sigma[j];
        double c = row[j] / (si * sj);
        c = c < -1.0 ? -1.0 : (c > 1.0 ? 1.0 : c);
        if (i == j) c = 1.0;
#pragma unroll
        for (int s = 0; s < EV3_VMAX; ++s)
            if (s < n_remove) c = fma(-vi[s], Vs[(size_t)s * N + j], c);
        row[j] = (c / (rsi * rs[j])) * (si * sj);
    }
}

int mcosb_dev_detone(mcosb_ctx* ctx, int S, double* dcov, int n_remove, int* dstatus) {
    if (S == 0 || n_remove == 0) return MCOSB_OK;   // n_remove == 0 returns the input (covariance_transformer.py:223-224)
    const int N = ctx->N;
    const size_t n = (size_t)N;
    if (n_remove < 0 || n_remove > EV3_VMAX || n_remove > N)
        return mcosb_fail(ctx, MCOSB_ERR_ARG, "detone: n_remove must be between 0 and min(16, n_assets)");
    if (N > 1024) return mcosb_fail(ctx, MCOSB_ERR_ARG, "detone: n_assets too large");
    double *A, *sigma, *d, *e, *tau, *lam, *W, *V;
    unsigned char* piv;
    int* nf;
    const size_t G = (size_t)S * EV3_VMAX;
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_A, (size_t)S * n * n, &A));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_SIGMA, (size_t)S * n, &sigma));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_D, (size_t)S * n, &d));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_E, (size_t)S * n, &e));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_TAU, (size_t)S * n, &tau));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_LAM, (size_t)S * n, &lam));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_NF, (size_t)S, &nf));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_WORK, 5 * n * G, &W));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_VEC, n * G, &piv));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_VAR, n * G, &V));
    const double* Q2 = nullptr;
    MCOSB_TRY(mcosb_dev_eig_prepare(ctx, S, dcov, A, sigma, d, e, tau, lam, &Q2));
    const int off = Q2 ? SBR_B : 1;
    fill_int_kernel<<<(S + 255) / 256, 256, 0, ctx->stream>>>(nf, (size_t)S, n_remove);
    MCOSB_LAUNCHED(ctx, MK_FILL);
    inviter_kernel<<<(unsigned)((G + 63) / 64), 64, 0, ctx->stream>>>(d, e, lam, nf, N, S, W, piv);
    MCOSB_LAUNCHED(ctx, MK_INVITER);
    if (Q2) MCOSB_TRY(mcosb_launch_q2back(ctx, S, Q2, nf, W + 4 * n * G) == 1 ? MCOSB_OK : MCOSB_ERR_STATE);
    const unsigned bt_grid = (unsigned)((G + 7) / 8);
    if (N <= 128) backtransform_kernel<4><<<bt_grid, 256, 2 * n * sizeof(double), ctx->stream>>>(A, tau, nf, N, S, W, V, dstatus, off);
    else if (N <= 256) backtransform_kernel<8><<<bt_grid, 256, 2 * n * sizeof(double), ctx->stream>>>(A, tau, nf, N, S, W, V, dstatus, off);
    else if (N <= 512) backtransform_kernel<16><<<bt_grid, 256, 2 * n * sizeof(double), ctx->stream>>>(A, tau, nf, N, S, W, V, dstatus, off);
    else if (N <= 768) backtransform_kernel<24><<<bt_grid, 256, 2 * n * sizeof(double), ctx->stream>>>(A, tau, nf, N, S, W, V, dstatus, off);
    else backtransform_kernel<32><<<bt_grid, 256, 2 * n * sizeof(double), ctx->stream>>>(A, tau, nf, N, S, W, V, dstatus, off);
    MCOSB_LAUNCHED(ctx, MK_BACKTRANSFORM);
    detone_kernel<<<dim3((N + 7) / 8, S), 256, n * sizeof(double), ctx->stream>>>(V, lam, n_remove, sigma, N, dcov);
    MCOSB_LAUNCHED(ctx, MK_DETONE);
    return MCOSB_OK;
}

extern "C" int mcosb_detone(mcosb_ctx* ctx, int S, double* cov_inout, int n_remove, int* status) {
    MCOSB_CHECK_CTX(ctx);
    if (S < 0 || !cov_inout) return mcosb_fail(ctx, MCOSB_ERR_ARG, "bad argument to mcosb_detone");
    const size_t N = ctx->N;
    Staging st(ctx);
    double* dcov;
    int* dst;
    MCOSB_TRY(st.inout(cov_inout, (size_t)S * N * N, SL_STAGE_IN0, &dcov));
    MCOSB_TRY(st.out(status, (size_t)S, SL_STAGE_OUT3, &dst));
    if (dst) MCOSB_CUDA(ctx, cudaMemsetAsync(dst, 0, (size_t)S * sizeof(int), ctx->stream));
    MCOSB_TRY(mcosb_dev_detone(ctx, S, dcov, n_remove, dst));
    return st.finish();
}
