// RiskParityOptimizer.allocate (optimizer.py:290-359) -- SURVEY 8(f) "next" #1.
//
// The reference minimises J(w) = sum_i (RC_i - b_i sigma_p)^2, RC_i = w_i (Cw)_i / sigma_p, with scipy SLSQP from
// w0 = b (the risk budget, 1/N by default), sum(w) = 1, w >= 0, default ftol 1e-6.  Two regimes:
//   * on covariances of realistic scale (entries ~1e-2) J(w0) ~ 1e-6 and SLSQP's first convergence test
//     |g's| < ftol (s = the first QP step = -(g - mean g)) fires before any step is taken: it returns w0 UNCHANGED;
//   * on O(1) covariances (the reference's 4 x 4 golden) it iterates to the risk-parity point, to ~1e-4 accuracy.
// This kernel reproduces both ends: it evaluates the analytic gradient at w0 and returns w0 when ||g - mean g||^2 < 1e-6;
// otherwise it solves risk budgeting exactly through Spinu's convex formulation
//        min_y  1/2 y'Cy - sum_i b_i ln y_i ,   w = y / sum(y)
// with a damped Newton method (dense Cholesky of C + diag(b / y^2) in global scratch, one CTA per simulation).
// In between (SLSQP stopping after a few iterations, far from converged) the reference's output is an artefact of its
// stopping rule and is not reproduced; DESIGN.md says so.
#include "common.cuh"

#define RP_THREADS 256

// y = C x for the symmetric matrix C (thread i accumulates column i = row i, coalesced)
__device__ void rp_symv(const double* __restrict__ C, int N, const double* __restrict__ x, double* __restrict__ y) {
    for (int i = threadIdx.x; i < N; i += RP_THREADS) {
        double acc = 0.0;
        for (int j = 0; j < N; ++j) acc = fma(C[(size_t)j * N + i], x[j], acc);
        y[i] = acc;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(RP_THREADS) risk_parity_kernel(const double* __restrict__ cov_all,
                                                                 const double* __restrict__ budget, int N,
                                                                 double* __restrict__ H_all, double* __restrict__ w_all,
                                                                 int* __restrict__ iters_out, int* __restrict__ status) {
    extern __shared__ double sm[];
    double* b = sm;          // [N] budget
    double* w = b + N;       // [N] current point (w0, then y)
    double* m = w + N;       // [N] C w
    double* r = m + N;       // [N]
    double* g = r + N;       // [N] gradient / Newton step
    double* t = g + N;       // [N]
    __shared__ double red[32];
    const int sim = blockIdx.x, tid = threadIdx.x;
    const double* C = cov_all + (size_t)sim * N * N;
    double* H = H_all + (size_t)sim * N * N;
    for (int i = tid; i < N; i += RP_THREADS) {
        b[i] = budget ? budget[i] : 1.0 / N;
        w[i] = b[i];
    }
    __syncthreads();
    // ---- SLSQP's first convergence test at w0 -------------------------------------------------------------------------
    rp_symv(C, N, w, m);
    double part = 0.0;
    for (int i = tid; i < N; i += RP_THREADS) part = fma(w[i], m[i], part);
    const double var0 = block_sum(part, red);
    const double s0 = sqrt(var0);
    double p1 = 0.0, p2 = 0.0;
    for (int i = tid; i < N; i += RP_THREADS) {
        const double ri = w[i] * m[i] / s0 - b[i] * s0;
        r[i] = ri;
        t[i] = ri * w[i];
        p1 = fma(ri * w[i], m[i], p1);
        p2 = fma(ri, b[i], p2);
    }
    const double rwm = block_sum(p1, red);
    const double rb = block_sum(p2, red);
    rp_symv(C, N, t, g);   // g <- C (r o w)
    double gs = 0.0;
    for (int i = tid; i < N; i += RP_THREADS) {
        const double gi = 2.0 * (r[i] * m[i] / s0 + g[i] / s0 - m[i] * rwm / (s0 * s0 * s0) - m[i] * rb / s0);
        g[i] = gi;
        gs += gi;
    }
    const double gmean = block_sum(gs, red) / N;
    double h = 0.0;
    for (int i = tid; i < N; i += RP_THREADS) { const double dgi = g[i] - gmean; h = fma(dgi, dgi, h); }
    const double h1 = block_sum(h, red);
    if (h1 < 1e-6 || !(h1 == h1)) {
        for (int i = tid; i < N; i += RP_THREADS) w_all[(size_t)sim * N + i] = w[i];
        if (tid == 0) {
            if (iters_out) iters_out[sim] = 0;
            if (status && !(h1 == h1)) atomicOr(&status[sim], MCOSB_ST_NONFINITE);
        }
        return;
    }
    // ---- exact risk budgeting: damped Newton on 1/2 y'Cy - sum b ln y ----------------------------------------------
    for (int i = tid; i < N; i += RP_THREADS) w[i] = sqrt(b[i] / C[(size_t)i * N + i]);
    __syncthreads();
    int it = 0;
    bool bad = false;
    for (; it < 60; ++it) {
        rp_symv(C, N, w, m);
        double gmax = 0.0;
        for (int i = tid; i < N; i += RP_THREADS) {
            const double gi = m[i] - b[i] / w[i];
            g[i] = -gi;
            gmax = fmax(gmax, fabs(gi) * w[i] / b[i]);    // relative to the barrier term
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = gmax;
        __syncthreads();
        gmax = red[0];
        for (int k = 1; k < RP_THREADS / 32; ++k) gmax = fmax(gmax, red[k]);
        __syncthreads();
        if (gmax < 1e-13) break;
        // H = C + diag(b / y^2); Cholesky in place (lower), solve H d = -grad
        for (size_t idx = tid; idx < (size_t)N * N; idx += RP_THREADS) {
            const int i = (int)(idx / N), j = (int)(idx % N);
            H[idx] = C[idx] + ((i == j) ? b[i] / (w[i] * w[i]) : 0.0);
        }
        __syncthreads();
        for (int k = 0; k < N; ++k) {
            const double dk = H[(size_t)k * N + k];
            if (!(dk > 0.0)) bad = true;
            const double piv = (dk > 0.0) ? sqrt(dk) : 1.0;
            __syncthreads();
            for (int i = k + tid; i < N; i += RP_THREADS) H[(size_t)i * N + k] = (i == k) ? piv : H[(size_t)i * N + k] / piv;
            __syncthreads();
            const int mm = N - k - 1;
            for (int idx = tid; idx < mm * mm; idx += RP_THREADS) {
                const int i = k + 1 + idx / mm, j = k + 1 + idx % mm;
                if (j <= i) H[(size_t)i * N + j] -= H[(size_t)i * N + k] * H[(size_t)j * N + k];
            }
            __syncthreads();
        }
        if (tid < 32) {   // triangular solves by one warp: column-oriented, lanes over rows
            const int lane = tid;
            for (int j = 0; j < N; ++j) {
                __syncwarp();
                const double xj = g[j] / H[(size_t)j * N + j];
                for (int i = j + 1 + lane; i < N; i += 32) g[i] -= H[(size_t)i * N + j] * xj;
                __syncwarp();
                if (lane == 0) g[j] = xj;
            }
            for (int j = N - 1; j >= 0; --j) {
                __syncwarp();
                const double xj = g[j] / H[(size_t)j * N + j];
                for (int i = lane; i < j; i += 32) g[i] -= H[(size_t)j * N + i] * xj;
                __syncwarp();
                if (lane == 0) g[j] = xj;
            }
        }
        __syncthreads();
        // damping: stay strictly inside y > 0
        double amax = 1.0;
        for (int i = tid; i < N; i += RP_THREADS)
            if (g[i] < 0.0) amax = fmin(amax, 0.9 * w[i] / (-g[i]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmin(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        if ((tid & 31) == 0) red[tid >> 5] = amax;
        __syncthreads();
        amax = red[0];
        for (int k = 1; k < RP_THREADS / 32; ++k) amax = fmin(amax, red[k]);
        __syncthreads();
        for (int i = tid; i < N; i += RP_THREADS) w[i] = fma(amax, g[i], w[i]);
        __syncthreads();
    }
    part = 0.0;
    for (int i = tid; i < N; i += RP_THREADS) part += w[i];
    const double tot = block_sum(part, red);
    for (int i = tid; i < N; i += RP_THREADS) w_all[(size_t)sim * N + i] = w[i] / tot;
    if (tid == 0) {
        if (iters_out) iters_out[sim] = it;
        if (status) {
            if (bad) atomicOr(&status[sim], MCOSB_ST_NOT_PD);
            if (it >= 60) atomicOr(&status[sim], MCOSB_ST_QP_CAP);
        }
    }
}

int mcosb_dev_risk_parity(mcosb_ctx* ctx, int S, const double* dcov, const double* dbudget, double* dw, int* diters,
                          int* dstatus) {
    if (S == 0) return MCOSB_OK;
    const int N = ctx->N;
    const size_t n = (size_t)N;
    double* H;
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_NCO_WORK, (size_t)S * n * n, &H));
    const size_t smem = 6 * n * sizeof(double);
    if (smem > (size_t)ctx->smem_optin) return mcosb_fail(ctx, MCOSB_ERR_ARG, "n_assets too large for risk_parity_kernel");
    MCOSB_CUDA(ctx, cudaFuncSetAttribute(risk_parity_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    risk_parity_kernel<<<S, RP_THREADS, smem, ctx->stream>>>(dcov, dbudget, N, H, dw, diters, dstatus);
    MCOSB_LAUNCHED(ctx, MK_RISK_PARITY);
    return MCOSB_OK;
}

extern "C" int mcosb_risk_parity(mcosb_ctx* ctx, int S, const double* cov, const double* budget, double* w, int* iters,
                                 int* status) {
    MCOSB_CHECK_CTX(ctx);
    if (S < 0 || !cov || !w) return mcosb_fail(ctx, MCOSB_ERR_ARG, "bad argument to mcosb_risk_parity");
    const size_t N = ctx->N;
    Staging st(ctx);
    const double *dcov, *db;
    double* dw;
    int *dit, *dst;
    MCOSB_TRY(st.in(cov, (size_t)S * N * N, SL_STAGE_IN0, &dcov));
    MCOSB_TRY(st.in(budget, N, SL_STAGE_IN1, &db));
    MCOSB_TRY(st.out(w, (size_t)S * N, SL_STAGE_OUT0, &dw));
    MCOSB_TRY(st.out(iters, (size_t)S, SL_STAGE_OUT1, &dit));
    MCOSB_TRY(st.out(status, (size_t)S, SL_STAGE_OUT2, &dst));
    if (dst) MCOSB_CUDA(ctx, cudaMemsetAsync(dst, 0, (size_t)S * sizeof(int), ctx->stream));
    MCOSB_TRY(mcosb_dev_risk_parity(ctx, S, dcov, db, dw, dit, dst));
    return st.finish();
}
