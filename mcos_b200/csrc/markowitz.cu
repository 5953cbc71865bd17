// K4a: MarkowitzOptimizer.allocate (optimizer.py:44-49 + PyPortfolioOpt 0.5.2 EfficientFrontier.max_sharpe /
// clean_weights), batched, one CTA per simulation.
//
// The reference hands  max (w.mu - rf)/sqrt(w'Cw), sum w = 1, 0 <= w <= 1  to scipy SLSQP (ftol 1e-6), whose answer is
// 1e-5..1.6e-4 away from the optimum of its own problem.  This kernel solves the same problem exactly: with
// c = mu - rf and some c_i > 0 it is the convex QP
//        min y'Cy   s.t.  c'y = 1, y >= 0,      w = y / sum(y)
// solved by a primal active-set method (one index enters or leaves the free set F per iteration).  Per iteration:
//   - C_FF z = c_F by Cholesky (factor kept in shared memory; appended row by row, refactored after a removal),
//   - grad = C[:,F] y_F   (rows of C are read coalesced from L2: |F| x N doubles),
//   - multipliers lam_i = grad_i - nu c_i, nu = y'Cy; most negative one enters; ratio test on the way to the target.
// The logic mirrors oracle/mcos_port.py::markowitz_exact_qp step for step (same start vertex, same tie-breaks).
#include "common.cuh"

#ifndef MK_THREADS
#define MK_THREADS 128   // measured: 128 -> 2.08, 256 -> 2.30, 512 -> 2.81 us per simulation at N = 500
#endif

struct MkSmem {
    double* c;      // [N]   mu - rf
    double* y;      // [N]   current iterate (zero off the free set)
    double* grad;   // [N]   C y
    double* z;      // [fcap] solve workspace / target on the free set
    double* b;      // [fcap] right-hand side workspace
    double* L;      // packed lower-triangular Cholesky factor of C_FF, row-major: L(i,j) at i(i+1)/2 + j
    int* F;         // [fcap] free set in insertion order
    double* red;    // [32]
    int* ired;      // [64]
};

__device__ __forceinline__ double& Lat(double* L, int i, int j) { return L[(size_t)i * (i + 1) / 2 + j]; }

// argmin/argmax helpers over block: value + lowest index on ties. Results identical on all threads.
__device__ void block_argmin(double v, int idx, MkSmem& s, double& vout, int& iout) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    __syncthreads();
    if (lane == 0) { s.red[wid] = v; s.ired[wid] = idx; }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    double bv = s.red[0];
    int bi = s.ired[0];
    for (int k = 1; k < nw; ++k) {
        double ov = s.red[k];
        int oi = s.ired[k];
        if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    vout = bv;
    iout = bi;
}

// Solve L L' z = rhs for the current f x f factor, by warp 0; result in s.z. Other warps wait at the barrier.
__device__ void chol_solve(MkSmem& s, int f) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        // forward: L t = b  (t overwrites b)
        for (int j = 0; j < f; ++j) {
            __syncwarp();
            const double xj = s.b[j] / Lat(s.L, j, j);
            for (int i = j + 1 + lane; i < f; i += 32) s.b[i] -= Lat(s.L, i, j) * xj;
            __syncwarp();
            if (lane == 0) s.b[j] = xj;
        }
        __syncwarp();
        // backward: L' z = t
        for (int j = f - 1; j >= 0; --j) {
            __syncwarp();
            const double xj = s.b[j] / Lat(s.L, j, j);
            for (int i = lane; i < j; i += 32) s.b[i] -= Lat(s.L, j, i) * xj;
            __syncwarp();
            if (lane == 0) s.z[j] = xj;
        }
    }
    __syncthreads();
}

// Appends index a (already stored in F[f]) as row f of the factor. Returns false if the new pivot is not positive.
__device__ bool chol_append(MkSmem& s, const double* __restrict__ C, int N, int f, int a) {
    for (int i = threadIdx.x; i < f; i += blockDim.x) s.b[i] = C[(size_t)s.F[i] * N + a];
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        for (int j = 0; j < f; ++j) {
            __syncwarp();
            const double xj = s.b[j] / Lat(s.L, j, j);
            for (int i = j + 1 + lane; i < f; i += 32) s.b[i] -= Lat(s.L, i, j) * xj;
            __syncwarp();
            if (lane == 0) s.b[j] = xj;
        }
    }
    __syncthreads();
    double part = 0.0;
    for (int i = threadIdx.x; i < f; i += blockDim.x) {
        double l = s.b[i];
        Lat(s.L, f, i) = l;
        part = fma(l, l, part);
    }
    const double ss = block_sum(part, s.red);
    const double piv = C[(size_t)a * N + a] - ss;
    const bool ok = piv > 0.0;
    if (threadIdx.x == 0) Lat(s.L, f, f) = ok ? sqrt(piv) : 1.0;
    __syncthreads();
    return ok;
}

// Full refactorisation of C_FF (after a removal). Returns false if a pivot is not positive.
__device__ bool chol_refactor(MkSmem& s, const double* __restrict__ C, int N, int f) {
    for (int idx = threadIdx.x; idx < f * (f + 1) / 2; idx += blockDim.x) {
        int i = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
        while ((i + 1) * (i + 2) / 2 <= idx) ++i;
        while (i * (i + 1) / 2 > idx) --i;
        int j = idx - i * (i + 1) / 2;
        s.L[idx] = C[(size_t)s.F[i] * N + s.F[j]];
    }
    __syncthreads();
    bool ok = true;
    for (int k = 0; k < f; ++k) {
        const double d = Lat(s.L, k, k);
        if (!(d > 0.0)) ok = false;
        const double piv = ok ? sqrt(d) : 1.0;
        __syncthreads();
        for (int i = k + threadIdx.x; i < f; i += blockDim.x) Lat(s.L, i, k) = (i == k) ? piv : Lat(s.L, i, k) / piv;
        __syncthreads();
        const int m = f - k - 1;
        for (int idx = threadIdx.x; idx < m * m; idx += blockDim.x) {
            int i = k + 1 + idx / m, j = k + 1 + idx % m;
            if (j <= i) Lat(s.L, i, j) -= Lat(s.L, i, k) * Lat(s.L, j, k);
        }
        __syncthreads();
    }
    return ok;
}

__global__ void __launch_bounds__(MK_THREADS) markowitz_kernel(const double* __restrict__ mu_all,
                                                               const double* __restrict__ cov_all, int N, int fcap,
                                                               double rf, int clean, int max_iter,
                                                               double* __restrict__ w_all, int* __restrict__ iters_out,
                                                               int* __restrict__ status, int* __restrict__ capflag,
                                                               int pass) {
    // two tiers: pass 0 runs everything with a small free-set capacity (30 KB of shared memory -> 7 CTAs per SM instead
    // of 2; the free set is ~50 at N = 500), pass 1 re-solves with the full capacity only what overflowed
    if (pass == 1 && capflag[blockIdx.x] == 0) return;
    extern __shared__ double sm[];
    MkSmem s;
    s.c = sm;
    s.y = s.c + N;
    s.grad = s.y + N;
    s.z = s.grad + N;
    s.b = s.z + fcap;
    s.L = s.b + fcap;
    s.red = s.L + (size_t)fcap * (fcap + 1) / 2;
    s.F = reinterpret_cast<int*>(s.red + 32);
    s.ired = s.F + fcap;
    const int sim = blockIdx.x, tid = threadIdx.x;
    const double* mu = mu_all + (size_t)sim * N;
    const double* C = cov_all + (size_t)sim * N * N;
    double* w = w_all + (size_t)sim * N;
    int st = 0;
    bool redo = false;

    // start vertex: best stand-alone Sharpe ratio among assets with c_i > 0
    double best = INFINITY;
    int besti = N;
    for (int i = tid; i < N; i += MK_THREADS) {
        double ci = mu[i] - rf;
        s.c[i] = ci;
        s.y[i] = 0.0;
        if (ci > 0.0) {
            double score = -(ci / sqrt(C[(size_t)i * N + i]));  // argmax via argmin of the negative
            if (score < best) { best = score; besti = i; }
        } else if (ci != ci) {
            st |= MCOSB_ST_NONFINITE;
        }
    }
    double bv;
    int start;
    block_argmin(best, besti, s, bv, start);
    if (__syncthreads_or(st)) {
        for (int i = tid; i < N; i += MK_THREADS) w[i] = NAN;
        if (tid == 0) {
            if (iters_out) iters_out[sim] = 0;
            if (status) status[sim] |= MCOSB_ST_NONFINITE;
        }
        return;
    }
    if (start >= N || !(bv < 0.0)) {
        // No asset beats the risk-free rate.  The QP reformulation does not apply; the stated problem
        // max (w.mu - rf) / sqrt(w'Cw) on the simplex then has a non-positive numerator everywhere, its objective is
        // quasi-convex and the maximum sits at a vertex: the single asset with the largest (least negative) stand-alone
        // Sharpe ratio.  (The reference hands this case to SLSQP, which also stops at a vertex -- a local one.)  The
        // simulation is flagged MCOSB_ST_NO_EXCESS but keeps a defined allocation.
        best = INFINITY;
        besti = N;
        for (int i = tid; i < N; i += MK_THREADS) {
            const double d = C[(size_t)i * N + i];
            if (d > 0.0) {
                const double score = -(s.c[i] / sqrt(d));
                if (score < best) { best = score; besti = i; }
            }
        }
        block_argmin(best, besti, s, bv, start);
        for (int i = tid; i < N; i += MK_THREADS) w[i] = (start < N) ? (i == start ? 1.0 : 0.0) : NAN;
        if (tid == 0) {
            if (iters_out) iters_out[sim] = 0;
            if (status) status[sim] |= MCOSB_ST_NO_EXCESS | (start < N ? 0 : MCOSB_ST_NOT_PD);
        }
        return;
    }
    int f = 1;
    if (tid == 0) {
        s.F[0] = start;
        double d = C[(size_t)start * N + start];
        s.L[0] = sqrt(d);
        s.y[start] = 1.0 / s.c[start];
    }
    __syncthreads();
    if (!(C[(size_t)start * N + start] > 0.0)) st |= MCOSB_ST_NOT_PD;

    int iters = 0;
    while (!st) {
        if (++iters > max_iter) { st |= MCOSB_ST_QP_CAP; break; }
        // target = minimiser of y'Cy on the free set subject to c'y = 1
        for (int i = tid; i < f; i += MK_THREADS) s.b[i] = s.c[s.F[i]];
        __syncthreads();
        chol_solve(s, f);
        double part = 0.0;
        for (int i = tid; i < f; i += MK_THREADS) part = fma(s.c[s.F[i]], s.z[i], part);
        const double cz = block_sum(part, s.red);
        bool neg = false;
        for (int i = tid; i < f; i += MK_THREADS) {
            double t = s.z[i] / cz;
            s.z[i] = t;
            if (!(t > 0.0)) neg = true;
        }
        const bool any_neg = __syncthreads_or(neg);
        if (!any_neg) {
            // accept the target, test optimality
            for (int i = tid; i < f; i += MK_THREADS) s.y[s.F[i]] = s.z[i];
            __syncthreads();
            for (int i = tid; i < N; i += MK_THREADS) {
                double g = 0.0;
                for (int k = 0; k < f; ++k) g = fma(C[(size_t)s.F[k] * N + i], s.z[k], g);
                s.grad[i] = g;
            }
            __syncthreads();
            part = 0.0;
            for (int i = tid; i < f; i += MK_THREADS) part = fma(s.z[i], s.grad[s.F[i]], part);
            const double nu = block_sum(part, s.red);
            double lmin = INFINITY;
            int imin = N;
            for (int i = tid; i < N; i += MK_THREADS) {
                if (s.y[i] == 0.0) {  // off the free set (free entries are strictly positive here)
                    double lam = s.grad[i] - nu * s.c[i];
                    if (lam < lmin) { lmin = lam; imin = i; }
                }
            }
            double lv;
            int enter;
            block_argmin(lmin, imin, s, lv, enter);
            if (enter >= N || lv >= -1e-12 * fmax(1.0, fabs(nu))) break;  // KKT satisfied
            if (f >= fcap) {
                if (pass == 0 && capflag) { if (tid == 0) capflag[sim] = 1; redo = true; }
                else st |= MCOSB_ST_QP_CAP;
                break;
            }
            if (tid == 0) s.F[f] = enter;
            __syncthreads();
            if (!chol_append(s, C, N, f, enter)) st |= MCOSB_ST_NOT_PD;
            ++f;
        } else {
            // move towards the target until the first free variable reaches zero, drop it
            double rmin = INFINITY;
            int kmin = fcap;
            for (int i = tid; i < f; i += MK_THREADS) {
                double cur = s.y[s.F[i]], step = s.z[i] - cur;
                if (step < 0.0) {
                    double r = cur / (-step);
                    if (r < rmin) { rmin = r; kmin = i; }
                }
            }
            double rv;
            int k;
            block_argmin(rmin, kmin, s, rv, k);
            const double alpha = fmin(1.0, rv);
            const int dropped = s.F[k];
            __syncthreads();
            for (int i = tid; i < f; i += MK_THREADS) {
                int a = s.F[i];
                s.y[a] = (i == k) ? 0.0 : s.y[a] + alpha * (s.z[i] - s.y[a]);
            }
            __syncthreads();
            if (tid == 0) {
                for (int i = k; i + 1 < f; ++i) s.F[i] = s.F[i + 1];
                s.y[dropped] = 0.0;
            }
            --f;
            __syncthreads();
            if (f == 0) { st |= MCOSB_ST_QP_CAP; break; }
            if (!chol_refactor(s, C, N, f)) st |= MCOSB_ST_NOT_PD;
        }
    }
    if (redo) return;   // pass 1 solves this simulation again with the full capacity
    // w = y / sum(y); clean_weights: |w| < 1e-4 -> 0, round half-even to 5 decimals, no renormalisation
    double part = 0.0;
    for (int i = tid; i < N; i += MK_THREADS) part += s.y[i];
    const double tot = block_sum(part, s.red);
    for (int i = tid; i < N; i += MK_THREADS) {
        double wi = s.y[i] / tot;
        if (clean) {
            if (fabs(wi) < 1e-4) wi = 0.0;
            wi = rint(wi * 1e5) / 1e5;
        }
        w[i] = wi;
    }
    if (tid == 0) {
        if (iters_out) iters_out[sim] = iters;
        if (status && st) status[sim] |= st;
    }
}

static int mk_fcap(int N) { return N <= 128 ? N : (N <= 600 ? 128 : 192); }

int mcosb_dev_markowitz(mcosb_ctx* ctx, int S, const double* dmu, const double* dcov, double rf, int clean,
                        double* dw, int* diters, int* dstatus) {
    if (S == 0) return MCOSB_OK;
    const int N = ctx->N;
    const int fcap = mk_fcap(N);
    auto smem_of = [&](int cap) {
        return (3 * (size_t)N + 2 * cap + (size_t)cap * (cap + 1) / 2 + 32) * sizeof(double) + ((size_t)cap + 64) * sizeof(int);
    };
    const size_t smem = smem_of(fcap);
    if (smem > (size_t)ctx->smem_optin) return mcosb_fail(ctx, MCOSB_ERR_ARG, "n_assets too large for markowitz_kernel");
    MCOSB_CUDA(ctx, cudaFuncSetAttribute(markowitz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int fcap0 = 64;
    if (fcap > fcap0 && S > 1) {
        int* capflag;
        MCOSB_TRY(mcosb_scratch_t(ctx, SL_MK_ITERS, (size_t)S, &capflag));
        MCOSB_CUDA(ctx, cudaMemsetAsync(capflag, 0, (size_t)S * sizeof(int), ctx->stream));
        markowitz_kernel<<<S, MK_THREADS, smem_of(fcap0), ctx->stream>>>(dmu, dcov, N, fcap0, rf, clean, 50 * N + 100, dw,
                                                                        diters, dstatus, capflag, 0);
        MCOSB_LAUNCHED(ctx, MK_MARKOWITZ);
        markowitz_kernel<<<S, MK_THREADS, smem, ctx->stream>>>(dmu, dcov, N, fcap, rf, clean, 50 * N + 100, dw, diters,
                                                              dstatus, capflag, 1);
    } else {
        markowitz_kernel<<<S, MK_THREADS, smem, ctx->stream>>>(dmu, dcov, N, fcap, rf, clean, 50 * N + 100, dw, diters,
                                                              dstatus, nullptr, 2);
    }
    MCOSB_LAUNCHED(ctx, MK_MARKOWITZ);
    return MCOSB_OK;
}

extern "C" int mcosb_markowitz(mcosb_ctx* ctx, int S, const double* mu, const double* cov, double rf, int clean,
                               double* w, int* iters, int* status) {
    MCOSB_CHECK_CTX(ctx);
    if (S < 0 || !mu || !cov || !w) return mcosb_fail(ctx, MCOSB_ERR_ARG, "bad argument to mcosb_markowitz");
    const size_t N = ctx->N;
    Staging st(ctx);
    const double *dmu, *dcov;
    double* dw;
    int *dit, *dst;
    MCOSB_TRY(st.in(mu, (size_t)S * N, SL_STAGE_IN0, &dmu));
    MCOSB_TRY(st.in(cov, (size_t)S * N * N, SL_STAGE_IN1, &dcov));
    MCOSB_TRY(st.out(w, (size_t)S * N, SL_STAGE_OUT0, &dw));
    MCOSB_TRY(st.out(iters, (size_t)S, SL_STAGE_OUT1, &dit));
    MCOSB_TRY(st.out(status, (size_t)S, SL_STAGE_OUT2, &dst));
    if (dst) MCOSB_CUDA(ctx, cudaMemsetAsync(dst, 0, (size_t)S * sizeof(int), ctx->stream));
    MCOSB_TRY(mcosb_dev_markowitz(ctx, S, dmu, dcov, rf, clean, dw, dit, dst));
    return st.finish();
}
