// Context lifetime, truth upload (+ one-off Cholesky), and the error-estimator kernel (K5).
#include <cstdlib>

#include "common.cuh"

extern "C" int mcosb_version(void) { return 200; }

extern "C" int mcosb_create(mcosb_ctx** out, int device, int n_assets, int n_obs, uint64_t seed) {
    if (!out || n_assets < 1 || n_obs < 1) return MCOSB_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return MCOSB_ERR_CUDA;  // no CPU fallback, by design
    if (device < 0 || device >= count) return MCOSB_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return MCOSB_ERR_CUDA;
    mcosb_ctx* ctx = new mcosb_ctx();
    ctx->device = device;
    ctx->N = n_assets;
    ctx->T = n_obs;
    ctx->seed = seed;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return MCOSB_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    {
        int lo_pri = 0, hi_pri = 0;
        cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);
        if (cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, hi_pri) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->ev_ideal, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->ev_main, cudaEventDisableTiming) != cudaSuccess) {
            mcosb_destroy(ctx);
            return MCOSB_ERR_CUDA;
        }
    }
    cudaDeviceGetAttribute(&ctx->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device);
    size_t n = (size_t)n_assets;
    if (cudaMalloc(&ctx->d_mu, n * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&ctx->d_cov, n * n * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&ctx->d_chol, n * n * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&ctx->d_klist, ((n + 63) / 64) * ((n + 15) / 16) * sizeof(unsigned short)) != cudaSuccess ||
        cudaMalloc(&ctx->d_nk, ((n + 63) / 64 + 1) * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&ctx->d_hrp_slow, sizeof(int)) != cudaSuccess || cudaMemset(ctx->d_hrp_slow, 0, sizeof(int)) != cudaSuccess) {
        mcosb_destroy(ctx);
        return MCOSB_ERR_CUDA;
    }
    *out = ctx;
    return MCOSB_OK;
}

extern "C" int mcosb_destroy(mcosb_ctx* ctx) {
    if (!ctx) return MCOSB_ERR_ARG;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < SL_COUNT; ++i)
        if (ctx->slot_ptr[i]) cudaFree(ctx->slot_ptr[i]);
    if (ctx->d_mu) cudaFree(ctx->d_mu);
    if (ctx->d_cov) cudaFree(ctx->d_cov);
    if (ctx->d_chol) cudaFree(ctx->d_chol);
    if (ctx->d_klist) cudaFree(ctx->d_klist);
    if (ctx->d_nk) cudaFree(ctx->d_nk);
    if (ctx->d_hrp_slow) cudaFree(ctx->d_hrp_slow);
    if (ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); }
    if (ctx->ev_ideal) cudaEventDestroy(ctx->ev_ideal);
    if (ctx->ev_main) cudaEventDestroy(ctx->ev_main);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return MCOSB_OK;
}

extern "C" const char* mcosb_last_error(const mcosb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int mcosb_set_stream(mcosb_ctx* ctx, void* cuda_stream) {
    MCOSB_CHECK_CTX(ctx);
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return MCOSB_OK;
}

extern "C" int mcosb_synchronize(mcosb_ctx* ctx) {
    MCOSB_CHECK_CTX(ctx);
    MCOSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MCOSB_OK;
}

extern "C" long long mcosb_launch_count(const mcosb_ctx* ctx) { return ctx ? ctx->launches : -1; }

extern "C" int mcosb_last_stage_ms(mcosb_ctx* ctx, double* stage_ms) {
    if (!ctx || !stage_ms) return MCOSB_ERR_ARG;
    for (int i = 0; i < MCOSB_N_STAGES; ++i) stage_ms[i] = ctx->stage_ms[i];
    return MCOSB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Cholesky of the true covariance (single CTA, in place on a copy), once per truth -- but the end-to-end path uploads
// the truth with every call, so it is not free.  Left-looking by columns: row k of L is staged in shared memory, every
// warp forms the dot products L[i][:k] . L[k][:k] of its rows with coalesced reads (N^3 / 6 FMA, no trailing update
// through global memory: the right-looking version took 11 ms at N = 500).
// flag[0] is set to 1 when a pivot is not positive.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) chol_truth_kernel(double* __restrict__ L, int n, int* __restrict__ flag) {
    extern __shared__ double Lk[];   // [n] row k of L
    __shared__ double s_piv;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int k = 0; k < n; ++k) {
        for (int j = tid; j < k; j += 1024) Lk[j] = L[(size_t)k * n + j];
        __syncthreads();
        for (int i = k + w; i < n; i += 32) {
            const double* row = L + (size_t)i * n;
            double dot = 0.0;
            for (int j = lane; j < k; j += 32) dot = fma(row[j], Lk[j], dot);
            dot = warp_sum(dot);
            if (lane == 0) {
                double v = row[k] - dot;
                if (i == k) {
                    if (!(v > 0.0)) { flag[0] = 1; v = 1.0; }
                    v = sqrt(v);
                    s_piv = v;
                }
                L[(size_t)i * n + k] = v;
            }
        }
        __syncthreads();
        const double piv = s_piv;
        for (int i = k + 1 + tid; i < n; i += 1024) L[(size_t)i * n + k] /= piv;
        __syncthreads();
    }
    for (int idx = tid; idx < n * n; idx += 1024) {
        int i = idx / n, j = idx % n;
        if (j > i) L[idx] = 0.0;
    }
}

// Blocked left-looking form of the same factorisation (the end-to-end path re-factorises the truth with every call: the
// one-CTA kernel above is 5.1 ms at N = 500, 1500 block-wide barriers and 500 dependent L2 round trips).  Per panel of
// CHP = 32 columns starting at p:
//   chol_update_kernel   A[i][p + c] -= sum_{k < p} L[i][k] L[p + c][k] for the rows i >= p, 32 x 32 outputs per CTA, k in
//                        chunks of 32 through shared memory (both operand tiles read coalesced along k);
//   chol_panel_kernel    one CTA: the updated panel (N - p rows x 32 columns) in shared memory, 32 column steps (pivot,
//                        scale, update of the columns to its right), written back; the upper triangle of its rows zeroed.
// Exact zeros stay exact zeros (sums of products with zero), which is what the TRMM's slab list relies on.
#define CHP 32
__global__ void __launch_bounds__(CHP * CHP) chol_update_kernel(double* __restrict__ L, int n, int p) {
    __shared__ double Li[CHP][CHP + 1], Lc[CHP][CHP + 1];
    const int tc = threadIdx.x & (CHP - 1), ti = threadIdx.x / CHP;
    const int i0 = p + blockIdx.x * CHP;
    double acc = 0.0;
    for (int k0 = 0; k0 < p; k0 += CHP) {
        const int k = k0 + tc;            // lanes along k: coalesced
        Li[ti][tc] = (i0 + ti < n && k < p) ? L[(size_t)(i0 + ti) * n + k] : 0.0;
        Lc[ti][tc] = (p + ti < n && k < p) ? L[(size_t)(p + ti) * n + k] : 0.0;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < CHP; ++kk) acc = fma(Li[ti][kk], Lc[tc][kk], acc);
        __syncthreads();
    }
    const int i = i0 + ti, j = p + tc;
    if (i < n && j < n && j <= i) L[(size_t)i * n + j] -= acc;
}

__global__ void __launch_bounds__(1024) chol_panel_kernel(double* __restrict__ L, int n, int p, int* __restrict__ flag) {
    extern __shared__ double P[];         // [(n - p)][CHP + 1]
    __shared__ double s_piv;
    const int m = n - p, w = min(CHP, n - p), tid = threadIdx.x;
    for (int idx = tid; idx < m * CHP; idx += 1024) {
        const int r = idx / CHP, c = idx % CHP;
        P[r * (CHP + 1) + c] = (c < w && c <= r) ? L[(size_t)(p + r) * n + p + c] : 0.0;
    }
    __syncthreads();
    for (int c = 0; c < w; ++c) {
        if (tid == 0) {
            double v = P[c * (CHP + 1) + c];
            if (!(v > 0.0)) { flag[0] = 1; v = 1.0; }
            v = sqrt(v);
            P[c * (CHP + 1) + c] = v;
            s_piv = v;
        }
        __syncthreads();
        const double piv = s_piv;
        for (int r = c + 1 + tid; r < m; r += 1024) P[r * (CHP + 1) + c] /= piv;
        __syncthreads();
        // columns c + 1 .. w - 1 of the rows below: P[r][c2] -= P[r][c] P[c2][c]
        const int nc = w - c - 1;
        for (int idx = tid; idx < (m - c - 1) * nc; idx += 1024) {
            const int r = c + 1 + idx / nc, c2 = c + 1 + idx % nc;
            if (c2 <= r) P[r * (CHP + 1) + c2] = fma(-P[r * (CHP + 1) + c], P[c2 * (CHP + 1) + c], P[r * (CHP + 1) + c2]);
        }
        __syncthreads();
    }
    for (int idx = tid; idx < m * CHP; idx += 1024) {
        const int r = idx / CHP, c = idx % CHP;
        if (c < w && c <= r) L[(size_t)(p + r) * n + p + c] = P[r * (CHP + 1) + c];
    }
    // upper triangle of the panel's rows
    for (int idx = tid; idx < w * n; idx += 1024) {
        const int r = idx / n, j = idx % n;
        if (j > p + r) L[(size_t)(p + r) * n + j] = 0.0;
    }
}

// The k-slabs of L a TRMM column tile has to visit (ctx->d_klist): one CTA per tile of 64 rows of L, one thread per slab of
// 16 columns up to the diagonal; the non-empty ones are listed in ascending order.  Dense factors list every slab up to
// the diagonal, which is what the TRMM walked before the list existed.
__global__ void __launch_bounds__(128) chol_slab_list_kernel(const double* __restrict__ L, int n,
                                                             unsigned short* __restrict__ klist, int* __restrict__ nk) {
    __shared__ unsigned short flags[128];
    const int J = blockIdx.x, KL = (n + 15) / 16;
    const int kdiag = (min(n, (J + 1) * 64) + 15) / 16;   // slabs at or below the diagonal of the tile's last row
    for (int base = 0, cnt = 0; base < kdiag; base += 128) {
        const int kt = base + threadIdx.x;
        unsigned short f = 0;
        if (kt < kdiag) {
            for (int r = 0; r < 64; ++r) {
                const int j = J * 64 + r;
                if (j >= n) break;
                const double* row = L + (size_t)j * n;
                bool nz = false;
                for (int k = kt * 16; k < min(n, kt * 16 + 16); ++k) nz |= (row[k] != 0.0);
                if (nz) f |= (r < 32) ? 0x4000 : 0x8000;
            }
        }
        flags[threadIdx.x] = f;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 0; i < 128 && base + i < kdiag; ++i)
                if (flags[i]) klist[(size_t)J * KL + cnt++] = (unsigned short)((base + i) | flags[i]);
            if (base + 128 >= kdiag) nk[J] = cnt;
            // half-slabs (32 rows x 16 columns) at or below the diagonal that turned out empty: any of them makes the list
            // worth reading
            int skipped = 0;
            for (int i = 0; i < 128 && base + i < kdiag; ++i)
                for (int hf = 0; hf < 2; ++hf) {
                    const int jlast = min(n, J * 64 + 32 * hf + 32) - 1;   // last row of the half
                    if (jlast >= J * 64 + 32 * hf && (base + i) * 16 <= jlast && !(flags[i] & (hf ? 0x8000 : 0x4000))) ++skipped;
                }
            if (skipped) atomicAdd(&nk[gridDim.x], skipped);
        }
        // cnt lives in thread 0 only; the other threads never read it
        __syncthreads();
    }
}

extern "C" int mcosb_set_truth(mcosb_ctx* ctx, const double* mu, const double* cov) {
    MCOSB_CHECK_CTX(ctx);
    if (!mu || !cov) return mcosb_fail(ctx, MCOSB_ERR_ARG, "mu/cov must not be null");
    size_t n = (size_t)ctx->N;
    cudaMemcpyKind kmu = mcosb_is_device_ptr(mu) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    cudaMemcpyKind kcov = mcosb_is_device_ptr(cov) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    MCOSB_CUDA(ctx, cudaMemcpyAsync(ctx->d_mu, mu, n * sizeof(double), kmu, ctx->stream));
    MCOSB_CUDA(ctx, cudaMemcpyAsync(ctx->d_cov, cov, n * n * sizeof(double), kcov, ctx->stream));
    MCOSB_CUDA(ctx, cudaMemcpyAsync(ctx->d_chol, ctx->d_cov, n * n * sizeof(double), cudaMemcpyDeviceToDevice,
                                    ctx->stream));
    int* dflag = nullptr;
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_STATUS, 1, &dflag));
    MCOSB_CUDA(ctx, cudaMemsetAsync(dflag, 0, sizeof(int), ctx->stream));
    const size_t smem_panel = n * (CHP + 1) * sizeof(double);
    static const bool chol_one_cta = getenv("MCOSB_CHOL_ONE_CTA") != nullptr;   // the one-CTA kernel (A/B runs)
    if (!chol_one_cta && smem_panel <= (size_t)ctx->smem_optin) {
        MCOSB_CUDA(ctx, cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_panel));
        for (int p = 0; p < ctx->N; p += CHP) {
            if (p > 0) {
                chol_update_kernel<<<(ctx->N - p + CHP - 1) / CHP, CHP * CHP, 0, ctx->stream>>>(ctx->d_chol, ctx->N, p);
                MCOSB_LAUNCHED(ctx, MK_CHOL_TRUTH);
            }
            chol_panel_kernel<<<1, 1024, (size_t)(ctx->N - p) * (CHP + 1) * sizeof(double), ctx->stream>>>(ctx->d_chol, ctx->N, p, dflag);
            MCOSB_LAUNCHED(ctx, MK_CHOL_TRUTH);
        }
    } else {
        chol_truth_kernel<<<1, 1024, n * sizeof(double), ctx->stream>>>(ctx->d_chol, ctx->N, dflag);
        MCOSB_LAUNCHED(ctx, MK_CHOL_TRUTH);
    }
    const int n_tiles = (ctx->N + 63) / 64;
    MCOSB_CUDA(ctx, cudaMemsetAsync(ctx->d_nk + n_tiles, 0, sizeof(int), ctx->stream));
    chol_slab_list_kernel<<<n_tiles, 128, 0, ctx->stream>>>(ctx->d_chol, ctx->N, ctx->d_klist, ctx->d_nk);
    MCOSB_LAUNCHED(ctx, MK_CHOL_TRUTH);
    int flag = 0, skipped = 0;
    MCOSB_CUDA(ctx, cudaMemcpyAsync(&flag, dflag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MCOSB_CUDA(ctx, cudaMemcpyAsync(&skipped, ctx->d_nk + n_tiles, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MCOSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->chol_sparse = skipped > 0;
    ctx->ideal_key.clear();  // ideal allocations belong to the previous truth
    ctx->have_truth = true;  // mu/cov are usable by the estimators even if sampling is not
    if (flag) return mcosb_fail(ctx, MCOSB_ERR_NOT_PD, "true covariance is not positive definite");
    return MCOSB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// K5: error estimators (error_estimator.py:21-49).  delta = w* - w;  EO = delta.mu;  Var = delta' Cov delta.
// One CTA handles ERR_SB simulations so each row of the true covariance read from L2 is reused ERR_SB times.
// ---------------------------------------------------------------------------------------------------------------
#define ERR_SB 8
#define ERR_THREADS 256

__global__ void __launch_bounds__(ERR_THREADS) errors_kernel(const double* __restrict__ w,
                                                             const double* __restrict__ wstar, int batched,
                                                             const double* __restrict__ mu,
                                                             const double* __restrict__ cov, int n, int S, int nk,
                                                             int k0, int k1, int k2, double* __restrict__ err, int ld) {
    // the kinds wanted, in output order: err[s * ld + k] = estimator kinds[k]; Var / EO are formed only if some kind
    // needs them, with the same instructions whatever else is asked for (bit-equal to the single-kind call)
    const int kinds[3] = {k0, k1, k2};
    bool need_var = false, need_eo = false;
    for (int k = 0; k < nk; ++k) {
        need_var |= kinds[k] != MCOSB_ERR_EXPECTED_OUTCOME;
        need_eo |= kinds[k] != MCOSB_ERR_VARIANCE;
    }
    extern __shared__ double sm[];
    double* delta = sm;                  // [ERR_SB][n]
    double* red = sm + (size_t)ERR_SB * n;  // [ERR_THREADS/32][ERR_SB][2]
    const int s0 = blockIdx.x * ERR_SB;
    const int ns = min(ERR_SB, S - s0);
    for (int idx = threadIdx.x; idx < ERR_SB * n; idx += blockDim.x) {
        int s = idx / n, j = idx % n;
        double d = 0.0;
        if (s < ns) {
            double ws = batched ? wstar[(size_t)(s0 + s) * n + j] : wstar[j];
            d = ws - w[(size_t)(s0 + s) * n + j];
        }
        delta[idx] = d;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double var_acc[ERR_SB], eo_acc[ERR_SB];
#pragma unroll
    for (int s = 0; s < ERR_SB; ++s) var_acc[s] = eo_acc[s] = 0.0;
    if (need_var) {
        for (int i = wid; i < n; i += nw) {
            double t[ERR_SB];
#pragma unroll
            for (int s = 0; s < ERR_SB; ++s) t[s] = 0.0;
            const double* row = cov + (size_t)i * n;
            for (int j = lane; j < n; j += 32) {
                double c = row[j];
#pragma unroll
                for (int s = 0; s < ERR_SB; ++s) t[s] = fma(c, delta[s * n + j], t[s]);
            }
#pragma unroll
            for (int s = 0; s < ERR_SB; ++s) {
                double v = warp_sum(t[s]);
                var_acc[s] = fma(delta[s * n + i], v, var_acc[s]);  // identical on all lanes
            }
        }
    }
    if (need_eo) {
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            double m = mu[j];
#pragma unroll
            for (int s = 0; s < ERR_SB; ++s) eo_acc[s] = fma(delta[s * n + j], m, eo_acc[s]);
        }
#pragma unroll
        for (int s = 0; s < ERR_SB; ++s) eo_acc[s] = warp_sum(eo_acc[s]);
    }
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < ERR_SB; ++s) {
            red[(wid * ERR_SB + s) * 2 + 0] = var_acc[s];
            red[(wid * ERR_SB + s) * 2 + 1] = eo_acc[s];
        }
    }
    __syncthreads();
    if (threadIdx.x < ns) {
        int s = threadIdx.x;
        double var = 0.0, eo = 0.0;
        for (int k = 0; k < nw; ++k) {
            var += red[(k * ERR_SB + s) * 2 + 0];
            eo += red[(k * ERR_SB + s) * 2 + 1];
        }
        for (int k = 0; k < nk; ++k) {
            double out;
            if (kinds[k] == MCOSB_ERR_EXPECTED_OUTCOME) out = eo;
            else if (kinds[k] == MCOSB_ERR_VARIANCE) out = var;
            else out = eo / sqrt(var);
            err[(size_t)(s0 + s) * ld + k] = out;
        }
    }
}

int mcosb_dev_errors(mcosb_ctx* ctx, int S, const double* dw, const double* dwstar, int batched, int nk,
                     const int* kinds, double* derr, int ld) {
    if (!ctx->have_truth) return mcosb_fail(ctx, MCOSB_ERR_STATE, "mcosb_set_truth has not been called");
    if (nk < 1 || nk > 3 || !kinds) return mcosb_fail(ctx, MCOSB_ERR_ARG, "between 1 and 3 error estimator kinds");
    int kk[3] = {0, 0, 0};
    for (int k = 0; k < nk; ++k) {
        if (kinds[k] < 0 || kinds[k] > 2) return mcosb_fail(ctx, MCOSB_ERR_ARG, "unknown error estimator kind");
        kk[k] = kinds[k];
    }
    if (S == 0) return MCOSB_OK;
    size_t smem = ((size_t)ERR_SB * ctx->N + (ERR_THREADS / 32) * ERR_SB * 2) * sizeof(double);
    if (smem > 48 * 1024)
        MCOSB_CUDA(ctx, cudaFuncSetAttribute(errors_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = (S + ERR_SB - 1) / ERR_SB;
    errors_kernel<<<grid, ERR_THREADS, smem, ctx->stream>>>(dw, dwstar, batched, ctx->d_mu, ctx->d_cov, ctx->N, S,
                                                           nk, kk[0], kk[1], kk[2], derr, ld);
    MCOSB_LAUNCHED(ctx, MK_ERRORS);
    return MCOSB_OK;
}

extern "C" int mcosb_errors(mcosb_ctx* ctx, int S, const double* w, const double* w_star, int w_star_batched,
                            int kind, double* err) {
    MCOSB_CHECK_CTX(ctx);
    if (S < 0 || !w || !w_star || !err) return mcosb_fail(ctx, MCOSB_ERR_ARG, "bad argument to mcosb_errors");
    size_t n = (size_t)ctx->N;
    Staging st(ctx);
    const double *dw, *dws;
    double* derr;
    MCOSB_TRY(st.in(w, (size_t)S * n, SL_STAGE_IN0, &dw));
    MCOSB_TRY(st.in(w_star, w_star_batched ? (size_t)S * n : n, SL_STAGE_IN1, &dws));
    MCOSB_TRY(st.out(err, (size_t)S, SL_STAGE_OUT0, &derr));
    MCOSB_TRY(mcosb_dev_errors(ctx, S, dw, dws, w_star_batched, 1, &kind, derr, 1));
    return st.finish();
}

extern "C" int mcosb_errors_multi(mcosb_ctx* ctx, int S, const double* w, const double* w_star, int w_star_batched,
                                  int n_kinds, const int* kinds, double* err) {
    MCOSB_CHECK_CTX(ctx);
    if (S < 0 || !w || !w_star || !err || !kinds || n_kinds < 1 || n_kinds > 3)
        return mcosb_fail(ctx, MCOSB_ERR_ARG, "bad argument to mcosb_errors_multi");
    size_t n = (size_t)ctx->N;
    Staging st(ctx);
    const double *dw, *dws;
    double* derr;
    MCOSB_TRY(st.in(w, (size_t)S * n, SL_STAGE_IN0, &dw));
    MCOSB_TRY(st.in(w_star, w_star_batched ? (size_t)S * n : n, SL_STAGE_IN1, &dws));
    MCOSB_TRY(st.out(err, (size_t)S * n_kinds, SL_STAGE_OUT0, &derr));
    MCOSB_TRY(mcosb_dev_errors(ctx, S, dw, dws, w_star_batched, n_kinds, kinds, derr, n_kinds));
    return st.finish();
}

// ---------------------------------------------------------------------------------------------------------------
// Per-kernel timing: with profiling on, every launch is followed by an event on the context's stream; a kernel's
// duration is the gap between its event and the previous one.
// ---------------------------------------------------------------------------------------------------------------
extern "C" int mcosb_set_profiling(mcosb_ctx* ctx, int on) {
    MCOSB_CHECK_CTX(ctx);
    MCOSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& k : ctx->kev) cudaEventDestroy(k.second);
    ctx->kev.clear();
    for (int i = 0; i < MK_COUNT; ++i) { ctx->kernel_ms[i] = 0.0; ctx->kernel_launches[i] = 0; }
    ctx->profile = on != 0;
    MCOSB_PROFILE_MARK(ctx);   // anchor: the first kernel launched after this call has a predecessor to be timed against
    return MCOSB_OK;
}

extern "C" int mcosb_kernel_count(void) { return MK_COUNT; }

extern "C" const char* mcosb_kernel_name(int id) { return (id >= 0 && id < MK_COUNT) ? MCOSB_KERNEL_NAMES[id] : ""; }

extern "C" int mcosb_kernel_times(mcosb_ctx* ctx, double* ms, long long* launches) {
    MCOSB_CHECK_CTX(ctx);
    MCOSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t i = 1; i < ctx->kev.size(); ++i) {
        if (ctx->kev[i].first < 0) continue;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, ctx->kev[i - 1].second, ctx->kev[i].second) == cudaSuccess)
            ctx->kernel_ms[ctx->kev[i].first] += t;
    }
    if (!ctx->kev.empty()) {  // keep the last event as the anchor for what follows
        for (size_t i = 0; i + 1 < ctx->kev.size(); ++i) cudaEventDestroy(ctx->kev[i].second);
        auto last = ctx->kev.back();
        last.first = -1;
        ctx->kev.clear();
        ctx->kev.push_back(last);
    }
    for (int i = 0; i < MK_COUNT; ++i) {
        if (ms) ms[i] = ctx->kernel_ms[i];
        if (launches) launches[i] = ctx->kernel_launches[i];
    }
    return MCOSB_OK;
}
