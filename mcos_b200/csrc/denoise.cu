// K3: DeNoiserCovarianceTransformer.transform (covariance_transformer.py:62-208), batched, one CTA per matrix.
//
//   corr_kernel        cov -> corr (clipped), sigma                               (cov_to_corr, :9-18)
//   tridiag_kernel     Householder tridiagonalisation Q' R Q = T                  (first half of np.linalg.eigh, :105)
//   eigvals_kernel     all N eigenvalues of T by Sturm bisection, sorted descending (:105-107)
//   mpfit_kernel       Marcenko-Pastur / KDE fit: L-BFGS-B in one variable, analytic gradient (:111-192)
//   eigvec_kernel      the n_facts signal eigenvectors by inverse iteration + back-transformation, the clipped
//                      reconstruction in low-rank form, renormalisation and rescaling  (:194-208, :96)
//
// Why not Jacobi: at N = 500 a Jacobi eigensolver costs ~6 N^3 flops per sweep x 8 sweeps (6e9 flops) and needs A and V
// (4 MB) resident; Householder + bisection costs 4/3 N^3 (1.7e8) and only the n_facts (~10) eigenvectors the
// reconstruction needs are ever formed:  V L' V' = lbar I + V_s (L_s - lbar I) V_s'  because V V' = I.
#include "common.cuh"
#include "sbr.cuh"

int mcosb_launch_tridiag2(mcosb_ctx* ctx, int S, double* A, double* d, double* e, double* tau);
int mcosb_launch_eigvec3(mcosb_ctx* ctx, int S, const double* A, const double* d, const double* e, const double* tau,
                         const double* lam, const int* nf, const double* sigma, double* cov, int* status,
                         const double* Q2);

// ---------------------------------------------------------------------------------------------------------------
#ifndef CORR_ROWS
#define CORR_ROWS 32   // rows per CTA (a multiple of 8)
#endif
#ifndef CORR_UNROLL
#define CORR_UNROLL 8  // independent row loads per lane
#endif
// D1: correlation matrix.  grid = (ceil(N / CORR_ROWS), S), 256 threads, CORR_ROWS / 8 rows per warp; the CTA takes the N square roots of
// the diagonal once (strided reads) into shared memory instead of once per row.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) corr_kernel(const double* __restrict__ cov, int N, double* __restrict__ A,
                                                   double* __restrict__ sigma, const double* __restrict__ lwcoef) {
    extern __shared__ double sg[];   // [N]
    const int s = blockIdx.y;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double* c = cov + (size_t)s * N * N;
    double* a = A + (size_t)s * N * N;
    // a deferred Ledoit-Wolf shrink (lw_coef_kernel) is applied on the fly, with lw_finish_kernel's arithmetic:
    // cov_ij <- (1 - sh) cov_ij (+ sh m on the diagonal); the shrunk covariance itself is never written
    const double sh = lwcoef ? lwcoef[2 * (size_t)s] : 0.0, m = lwcoef ? lwcoef[2 * (size_t)s + 1] : 0.0;
    for (int j = threadIdx.x; j < N; j += 256) {
        double v = c[(size_t)j * N + j];
        if (lwcoef) v = lw_shrink(v, sh, m, true);
        sg[j] = sqrt(v);
    }
    __syncthreads();
    for (int r = 0; r < CORR_ROWS / 8; ++r) {
        const int row = blockIdx.x * CORR_ROWS + wid * (CORR_ROWS / 8) + r;
        if (row >= N) return;
        const double si = sg[row];
        if (lane == 0) sigma[(size_t)s * N + row] = si;
        // 8 loads in flight per lane: with one, the 64 warps of an SM keep ~24 KB outstanding and the kernel sits at 3.8 TB/s
        for (int j0 = lane; j0 < N; j0 += 32 * CORR_UNROLL) {
            double x[CORR_UNROLL];
#pragma unroll
            for (int u = 0; u < CORR_UNROLL; ++u) {
                const int j = j0 + 32 * u;
                x[u] = (j < N) ? __ldcs(c + (size_t)row * N + j) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < CORR_UNROLL; ++u) {
                const int j = j0 + 32 * u;
                if (j < N) {
                    double v = x[u];
                    if (lwcoef) v = lw_shrink(v, sh, m, j == row);
                    v = v / (si * sg[j]);
                    v = v < -1.0 ? -1.0 : (v > 1.0 ? 1.0 : v);  // NaN passes through, as in numpy
                    a[(size_t)row * N + j] = v;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// D2: Householder tridiagonalisation of the symmetric matrix A (full storage, both triangles kept current).
// One CTA per matrix; A lives in global memory (L2-resident while it is being worked on).  The rank-2 update of
// step k-1 is fused with the symmetric matrix-vector product of step k, so each step makes ONE read+write pass over
// the trailing block:  traffic = sum_k 16 (N-k)^2 bytes ~ 16 N^3 / 3.
// Outputs: d[N], e[N-1] (tridiagonal), tau[N]; reflector k is stored in row k: u_k = (1, A[k][k+2:]) on indices k+1...
// ---------------------------------------------------------------------------------------------------------------
#define TD_THREADS 512
#define TD_WARPS (TD_THREADS / 32)

__global__ void __launch_bounds__(TD_THREADS) tridiag_kernel(double* __restrict__ Aall, int N, double* __restrict__ dall,
                                                             double* __restrict__ eall, double* __restrict__ tauall) {
    extern __shared__ double sm[];
    double* vp = sm;            // pending reflector (step k-1), indexed by global row
    double* wp = vp + N;        // pending w
    double* v = wp + N;         // current reflector
    double* p = v + N;          // symv result / w under construction
    double* ppart = p + N;      // [TD_WARPS][32] row-group partials
    __shared__ double red[32];
    __shared__ double s_tau, s_beta;

    const int s = blockIdx.x;
    double* A = Aall + (size_t)s * N * N;
    double* d = dall + (size_t)s * N;
    double* e = eall + (size_t)s * N;
    double* tau = tauall + (size_t)s * N;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    bool pending = false;
    for (int k = 0; k + 2 < N; ++k) {
        // ---- bring row k up to date (pending update acts on indices >= k), read x = A[k][k+1:] -----------------
        double* rowk = A + (size_t)k * N;
        double xsq = 0.0;
        for (int j = k + tid; j < N; j += TD_THREADS) {
            double a = rowk[j];
            if (pending) a -= vp[k] * wp[j] + wp[k] * vp[j];
            if (j == k) d[k] = a;
            else {
                v[j] = a;  // x, to be scaled into the reflector
                if (j > k + 1) xsq += a * a;
            }
        }
        xsq = block_sum(xsq, red);  // contains the barrier that publishes v[]
        if (tid == 0) {
            double alpha = v[k + 1];
            double t = 0.0, beta = alpha;
            if (xsq != 0.0) {
                beta = -copysign(sqrt(alpha * alpha + xsq), alpha);
                t = (beta - alpha) / beta;
            }
            s_tau = t;
            s_beta = beta;
            e[k] = beta;
            tau[k] = t;
        }
        __syncthreads();
        const double tk = s_tau;
        {
            const double alpha = v[k + 1];
            const double scal = (tk != 0.0) ? 1.0 / (alpha - s_beta) : 0.0;
            __syncthreads();
            for (int j = k + 1 + tid; j < N; j += TD_THREADS) {
                double vj = (j == k + 1) ? 1.0 : v[j] * scal;
                v[j] = vj;
                if (j > k + 1) rowk[j] = vj;  // keep the reflector for the back-transformation
            }
        }
        __syncthreads();
        // ---- fused pass over rows/cols k+1..N-1: apply the pending update, accumulate p = A v --------------------
        const int base = k + 1, m = N - base;
        const int nch = (m + 31) >> 5;
        int nrg = 1;
        if (nch < TD_WARPS) nrg = TD_WARPS / nch;
        if (nch >= TD_WARPS || wid < nch * nrg) {
            const int rg = (nch >= TD_WARPS) ? 0 : wid / nch;
            for (int ch = (nch >= TD_WARPS) ? wid : wid % nch; ch < nch; ch += TD_WARPS) {
                const int j = base + ch * 32 + lane;
                const bool act = j < N;
                const double vpj = (act && pending) ? vp[j] : 0.0, wpj = (act && pending) ? wp[j] : 0.0;
                double acc = 0.0;
                double* col = A + j;
                int i = base + rg;
                if (pending) {
                    for (; i + 3 * nrg < N; i += 4 * nrg) {
                        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
                        if (act) {
                            a0 = col[(size_t)i * N];
                            a1 = col[(size_t)(i + nrg) * N];
                            a2 = col[(size_t)(i + 2 * nrg) * N];
                            a3 = col[(size_t)(i + 3 * nrg) * N];
                        }
                        a0 -= vp[i] * wpj + wp[i] * vpj;
                        a1 -= vp[i + nrg] * wpj + wp[i + nrg] * vpj;
                        a2 -= vp[i + 2 * nrg] * wpj + wp[i + 2 * nrg] * vpj;
                        a3 -= vp[i + 3 * nrg] * wpj + wp[i + 3 * nrg] * vpj;
                        if (act) {
                            col[(size_t)i * N] = a0;
                            col[(size_t)(i + nrg) * N] = a1;
                            col[(size_t)(i + 2 * nrg) * N] = a2;
                            col[(size_t)(i + 3 * nrg) * N] = a3;
                        }
                        acc = fma(a0, v[i], acc);
                        acc = fma(a1, v[i + nrg], acc);
                        acc = fma(a2, v[i + 2 * nrg], acc);
                        acc = fma(a3, v[i + 3 * nrg], acc);
                    }
                    for (; i < N; i += nrg) {
                        double a0 = act ? col[(size_t)i * N] : 0.0;
                        a0 -= vp[i] * wpj + wp[i] * vpj;
                        if (act) col[(size_t)i * N] = a0;
                        acc = fma(a0, v[i], acc);
                    }
                } else {
                    for (; i + 3 * nrg < N; i += 4 * nrg) {
                        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
                        if (act) {
                            a0 = col[(size_t)i * N];
                            a1 = col[(size_t)(i + nrg) * N];
                            a2 = col[(size_t)(i + 2 * nrg) * N];
                            a3 = col[(size_t)(i + 3 * nrg) * N];
                        }
                        acc = fma(a0, v[i], acc);
                        acc = fma(a1, v[i + nrg], acc);
                        acc = fma(a2, v[i + 2 * nrg], acc);
                        acc = fma(a3, v[i + 3 * nrg], acc);
                    }
                    for (; i < N; i += nrg) {
                        double a0 = act ? col[(size_t)i * N] : 0.0;
                        acc = fma(a0, v[i], acc);
                    }
                }
                if (nch >= TD_WARPS) {
                    if (act) p[j] = acc;
                } else {
                    ppart[(rg * nch + ch) * 32 + lane] = acc;
                }
            }
        }
        __syncthreads();
        if (nch < TD_WARPS) {
            for (int c = tid; c < m; c += TD_THREADS) {
                double t = 0.0;
                for (int r = 0; r < nrg; ++r) t += ppart[(r * nch + (c >> 5)) * 32 + (c & 31)];
                p[base + c] = t;
            }
            __syncthreads();
        }
        // ---- w = tau p - (tau/2)(tau p . v) v ---------------------------------------------------------------------
        double dot = 0.0;
        for (int j = base + tid; j < N; j += TD_THREADS) dot += p[j] * v[j];
        dot = block_sum(dot, red);
        const double a2 = -0.5 * tk * (tk * dot);
        for (int j = base + tid; j < N; j += TD_THREADS) {
            wp[j] = tk * p[j] + a2 * v[j];
            vp[j] = v[j];
        }
        pending = true;
        __syncthreads();
    }
    // ---- trailing 1x1 / 2x2 block ----------------------------------------------------------------------------------
    if (tid == 0) {
        if (N == 1) {
            d[0] = A[0];
        } else {
            const int k = N - 2;
            double a00 = A[(size_t)k * N + k], a01 = A[(size_t)k * N + k + 1], a11 = A[(size_t)(k + 1) * N + k + 1];
            if (pending) {
                a00 -= 2.0 * vp[k] * wp[k];
                a01 -= vp[k] * wp[k + 1] + wp[k] * vp[k + 1];
                a11 -= 2.0 * vp[k + 1] * wp[k + 1];
            }
            d[k] = a00;
            e[k] = a01;
            d[k + 1] = a11;
            tau[k] = 0.0;
        }
        e[N - 1] = 0.0;
        tau[N - 1] = 0.0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// D3: eigenvalues of the symmetric tridiagonal (d, e) by bisection on the Sturm sequence, one thread per eigenvalue.
// The Sturm count uses the division-free three-term recurrence p_{i+1} = (d_i - x) p_i - e_{i-1}^2 p_{i-1} (3 FP64
// instructions per step) with exponent rescaling every 8 steps; #sign changes of (p_0..p_N) = #eigenvalues < x.
// Output lam[N] sorted DESCENDING (the order _get_PCA produces, covariance_transformer.py:106-107).
// ---------------------------------------------------------------------------------------------------------------
#ifndef EV_THREADS
#define EV_THREADS 256
#endif

// sde[i] = (d_i, e_{i-1}^2), e_{-1} = 0.  Only the recurrence itself (DADD, DMUL, DFMA) runs on the FP64 pipe: the sign
// changes are tracked on the integer pipe from the high words (one funnel shift per step collects the change bits, one
// popc per 8 steps counts them), and an exact zero takes the sign opposite to its predecessor (Wilkinson) by a select on
// the sign word instead of a substituted tiny value.
__device__ __noinline__ int sturm_count_exact(const double2* __restrict__ sde, int n, double x) {
    double pm = 0.0, pc = 1.0;
    int hc = 0, cnt = 0;          // hc: high word carrying the effective sign of p_i
    unsigned acc = 0;
#define STURM_STEP(idx)                                                        \
    {                                                                          \
        const double2 de = sde[idx];                                           \
        const double pn = fma(de.x - x, pc, -de.y * pm);                       \
        const int hn = __double2hiint(pn);                                     \
        const bool zero = ((hn & 0x7fffffff) | __double2loint(pn)) == 0;       \
        const int he = zero ? ~hc : hn;                                        \
        acc = __funnelshift_l((unsigned)(he ^ hc), acc, 1);                    \
        hc = he;                                                               \
        pm = pc;                                                               \
        pc = pn;                                                               \
    }
    int i = 0;
    for (; i + 8 <= n; i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) STURM_STEP(i + k)
        cnt += __popc(acc);
        acc = 0;
        int ex = ((__double2hiint(pc) >> 20) & 0x7ff) - 1023;
        int exm = ((__double2hiint(pm) >> 20) & 0x7ff) - 1023;
        ex = max(ex, exm);
        if (ex > 300 || ex < -300) {
            const double sc = __hiloint2double((1023 - ex) << 20, 0);
            pc *= sc;
            pm *= sc;
        }
    }
    for (; i < n; ++i) STURM_STEP(i)
#undef STURM_STEP
    return cnt + __popc(acc);
}

// Fast path of the same count: the sign bits of p_1 .. p_n are packed raw (one funnel shift per step) and the changes are
// counted per 8 steps from the packed word; an exact zero anywhere (one DSETP per step, OR-ed into a predicate) sends
// the evaluation to sturm_count_exact, which applies Wilkinson's sign rule.  6 instead of 9 instructions per step.
__device__ __forceinline__ int sturm_count(const double2* __restrict__ sde, int n, double x) {
    double pm = 0.0, pc = 1.0;
    unsigned acc = 0, prev = 0;   // prev: sign bit of the last p of the previous group (p_0 = 1 > 0)
    int cnt = 0;
    bool anyzero = false;
#define STURM_FAST(idx)                                                        \
    {                                                                          \
        const double2 de = sde[idx];                                           \
        const double pn = fma(de.x - x, pc, -de.y * pm);                       \
        anyzero |= (pn == 0.0);                                                \
        acc = __funnelshift_l((unsigned)__double2hiint(pn), acc, 1);           \
        pm = pc;                                                               \
        pc = pn;                                                               \
    }
    int i = 0;
    for (; i + 8 <= n; i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) STURM_FAST(i + k)
        // bit k of acc = sign of the (8 - k)-th value of the group; its predecessor sits one bit higher (prev for bit 7)
        cnt += __popc((acc ^ ((acc >> 1) | (prev << 7))) & 0xffu);
        prev = acc & 1u;
        acc = 0;
        int ex = ((__double2hiint(pc) >> 20) & 0x7ff) - 1023;
        int exm = ((__double2hiint(pm) >> 20) & 0x7ff) - 1023;
        ex = max(ex, exm);
        if (ex > 300 || ex < -300) {
            const double sc = __hiloint2double((1023 - ex) << 20, 0);
            pc *= sc;
            pm *= sc;
        }
    }
    const int rem = n - i;
    if (rem > 0) {
        for (; i < n; ++i) STURM_FAST(i)
        const unsigned m = (1u << rem) - 1u;
        cnt += __popc((acc ^ ((acc >> 1) | (prev << (rem - 1)))) & m);
    }
#undef STURM_FAST
    if (anyzero) return sturm_count_exact(sde, n, x);
    return cnt;
}

// The count together with p_N(x) and its derivative (same scaling, so that -p / dp is Newton's correction): the derivative
// obeys p'_{i+1} = (d_i - x) p'_i - p_i - e_{i-1}^2 p'_{i-1}, three more FP64 instructions per step.  dp = 0 tells the caller
// that no correction is available (an exact zero in the sequence: the count then comes from the exact path).
__device__ __forceinline__ int sturm_newton(const double2* __restrict__ sde, int n, double x, double& p, double& dp) {
    double pm = 0.0, pc = 1.0, qm = 0.0, qc = 0.0;
    unsigned acc = 0, prev = 0;
    int cnt = 0;
    bool anyzero = false;
#define STURM_NEWTON(idx)                                                      \
    {                                                                          \
        const double2 de = sde[idx];                                           \
        const double t = de.x - x;                                             \
        const double pn = fma(t, pc, -de.y * pm);                              \
        const double qn = fma(t, qc, fma(-de.y, qm, -pc));                     \
        anyzero |= (pn == 0.0);                                                \
        acc = __funnelshift_l((unsigned)__double2hiint(pn), acc, 1);           \
        pm = pc;                                                               \
        pc = pn;                                                               \
        qm = qc;                                                               \
        qc = qn;                                                               \
    }
    int i = 0;
    for (; i + 8 <= n; i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) STURM_NEWTON(i + k)
        cnt += __popc((acc ^ ((acc >> 1) | (prev << 7))) & 0xffu);
        prev = acc & 1u;
        acc = 0;
        int ex = (__double2hiint(pc) >> 20) & 0x7ff;
        ex = max(ex, (__double2hiint(pm) >> 20) & 0x7ff);
        ex = max(ex, (__double2hiint(qc) >> 20) & 0x7ff);
        ex = max(ex, (__double2hiint(qm) >> 20) & 0x7ff) - 1023;
        if (ex > 300 || ex < -300) {
            const double sc = __hiloint2double((1023 - ex) << 20, 0);
            pc *= sc;
            pm *= sc;
            qc *= sc;
            qm *= sc;
        }
    }
    const int rem = n - i;
    if (rem > 0) {
        for (; i < n; ++i) STURM_NEWTON(i)
        const unsigned m = (1u << rem) - 1u;
        cnt += __popc((acc ^ ((acc >> 1) | (prev << (rem - 1)))) & m);
    }
#undef STURM_NEWTON
    p = pc;
    dp = qc;
    if (anyzero) {
        dp = 0.0;
        return sturm_count_exact(sde, n, x);
    }
    return cnt;
}

__global__ void __launch_bounds__(EV_THREADS) eigvals_kernel(const double* __restrict__ dall,
                                                             const double* __restrict__ eall, int N,
                                                             double* __restrict__ lamall) {
    extern __shared__ double2 sde[];   // [N] (d_i, e_{i-1}^2)
    __shared__ double red_lo[EV_THREADS / 32], red_hi[EV_THREADS / 32];
    const int s = blockIdx.x, tid = threadIdx.x;
    const double* d = dall + (size_t)s * N;
    const double* e = eall + (size_t)s * N;
    double lo = 1e300, hi = -1e300;
    int bad = 0;
    for (int i = tid; i < N; i += EV_THREADS) {
        double di = d[i], ei = (i < N - 1) ? e[i] : 0.0, em = (i > 0) ? e[i - 1] : 0.0;
        sde[i] = make_double2(di, em * em);
        bad |= !(isfinite(di) && isfinite(ei));
        double r = fabs(ei) + fabs(em);
        lo = fmin(lo, di - r);
        hi = fmax(hi, di + r);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((tid & 31) == 0) { red_lo[tid >> 5] = lo; red_hi[tid >> 5] = hi; }
    if (__syncthreads_or(bad)) {   // a non-finite tridiagonal (NaN in the input): NaN eigenvalues, flagged downstream
        for (int j = tid; j < N; j += EV_THREADS) lamall[(size_t)s * N + j] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    lo = red_lo[0]; hi = red_hi[0];
    for (int i = 1; i < EV_THREADS / 32; ++i) { lo = fmin(lo, red_lo[i]); hi = fmax(hi, red_hi[i]); }
    const double span = fmax(fabs(lo), fabs(hi));
    const double pad = 2.2e-16 * span * (2.0 * N) + 1e-300;
    lo -= pad;
    hi += pad;
    // one multisection step shared by the whole matrix: Sturm counts on a uniform grid of N points (one per thread and
    // round) give every eigenvalue a bracket of width span / N for one evaluation instead of log2(N) bisection steps
    int* cnts = reinterpret_cast<int*>(sde + N);   // [N] number of eigenvalues below grid point i + 1
    __syncthreads();
    const double h = (hi - lo) / (double)N;
    for (int i = tid; i < N; i += EV_THREADS) cnts[i] = sturm_count(sde, N, fma(h, (double)(i + 1), lo));
    __syncthreads();
    // below ~4 ulp of the spectrum's scale the tridiagonal form itself is not accurate; stop there
    const double tol = 8.9e-16 * fmax(span, 1e-300);
    // second multisection step, inside the cells: the m eigenvalues that share a cell (the bulk of a Marcenko-Pastur spectrum
    // puts ~9 into a cell of width span / N) probe m evenly spaced interior points, one each, and every one of them takes the
    // tightest bracket the m probes allow -- one cheap count instead of ~log2(m) Newton-capable evaluations each
    double* xs2 = reinterpret_cast<double*>(cnts + N + (N & 1));   // [N] probe of eigenvalue j
    int* cs2 = reinterpret_cast<int*>(xs2 + N);                    // [N] its count
    for (int j = tid; j < N; j += EV_THREADS) {
        int l0 = 0, l1 = N - 1;                  // smallest i with cnts[i] > j (cnts[N-1] = N)
        while (l0 < l1) {
            const int mid = (l0 + l1) >> 1;
            if (cnts[mid] > j) l1 = mid; else l0 = mid + 1;
        }
        const int ca = l0 > 0 ? cnts[l0 - 1] : 0, m = cnts[l0] - ca;
        const double a = fma(h, (double)l0, lo);
        const double x = fma(h * ((double)(j - ca) + 0.5), 1.0 / (double)m, a);
        xs2[j] = x;
        cs2[j] = (m > 1) ? sturm_count(sde, N, x) : -1;
    }
    __syncthreads();
    for (int j = tid; j < N; j += EV_THREADS) {  // j-th smallest eigenvalue
        int l0 = 0, l1 = N - 1;
        while (l0 < l1) {
            const int mid = (l0 + l1) >> 1;
            if (cnts[mid] > j) l1 = mid; else l0 = mid + 1;
        }
        double a = fma(h, (double)l0, lo), b = fma(h, (double)(l0 + 1), lo);
        int ca = l0 > 0 ? cnts[l0 - 1] : 0, cb = cnts[l0];
        if (cb - ca > 1) {
            const int c0 = ca, c1 = cb;          // the cell's eigenvalues are c0 .. c1 - 1; their probes ascend with the index
            for (int jj = c0; jj < c1; ++jj) {
                const double x = xs2[jj];
                const int c = cs2[jj];
                if (c > j) { if (x < b) { b = x; cb = c; } }
                else if (x > a) { a = x; ca = c; }
            }
        }
        // Bisection on the count until the bracket isolates eigenvalue j (counts j and j + 1 at its ends), then Newton on
        // p_N from the last point evaluated, kept only while it stays inside the bracket and at least halves the step;
        // the count of every evaluation still tightens the bracket, so a rejected step costs nothing.  Newton's iterate is
        // accepted once its correction is below 32 tol, or once the observed quadratic contraction |dx_k| / dx_{k-1}^2
        // puts the next error below tol / 10.  Clustered eigenvalues never isolate and stay on bisection.
        // (scratch/eig_newton_proto.py: ~10 evaluations per eigenvalue instead of 43 at N = 500.)
        double x = 0.5 * (a + b), dxp = b - a, lam = x;
        bool nw = false;
        for (int it = 0; it < 200; ++it) {
            if (!(x > a && x < b) || b - a <= tol) { lam = 0.5 * (a + b); break; }
            double p, dp;
            const int c = sturm_newton(sde, N, x, p, dp);
            if (c > j) { b = x; cb = c; } else { a = x; ca = c; }
            const double mid = 0.5 * (a + b);
            lam = mid;
            double xn = mid, dxn = b - a;
            bool nwn = false;
            if (ca == j && cb == j + 1 && dp != 0.0) {
                const double dx = -p / dp, adx = fabs(dx);
                const double t = fmin(fmax(x + dx, a), b);
                if (adx <= 32.0 * tol || (nw && adx < 0.01 * dxp && adx * adx * adx <= 0.1 * tol * dxp * dxp)) { lam = t; break; }
                if (t > a && t < b && adx < 0.5 * dxp) { xn = t; dxn = adx; nwn = true; }
            }
            x = xn;
            dxp = dxn;
            nw = nwn;
        }
        lamall[(size_t)s * N + (N - 1 - j)] = lam;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// D4: Marcenko-Pastur fit.  Objective (covariance_transformer.py:134-167): on the 1000-point grid
// x_k = linspace(v (1-sqrt(1/q))^2, v (1+sqrt(1/q))^2), f(v) = sum_k (kde(x_k) - mp(x_k; v))^2 with
// kde = Gaussian KDE of the eigenvalues (bandwidth h), mp = MP density.  Minimised over v in [1e-5, 1-1e-5] from 0.5 by
// a faithful one-variable restatement of scipy's L-BFGS-B (projected secant direction + More'-Thuente dcsrch/dcstep
// line search, pgtol 1e-5, factr 1e7), with the analytic derivative df/dv in place of scipy's forward difference.
// tests/lbfgsb1d_ref.py is the same logic in numpy, checked against scipy iteration for iteration.
// ---------------------------------------------------------------------------------------------------------------
#define MP_THREADS 256
#define MP_PTS 1000

struct MPCtx {
    const double* lam;  // shared memory, N eigenvalues
    int N;
    double q, inv2h2, invh2, cnorm;
    double* red;
    // Taylor table of the kernel sum (mp_build_nodes); nodes = 0: evaluate the sum directly
    const double* coef = nullptr;   // shared memory, [nodes][MP_DEG + 1]
    int nodes = 0;
    double h = 0.0, inv_delta = 0.0, delta = 0.0;
};

// The kernel sum S(x) = sum_i exp(-(x - lam_i)^2 / 2h^2) does not depend on the variance being fitted -- only the grid on which
// the objective samples it does, and every grid lies in (0, (1 + sqrt(1/q))^2).  So S is tabulated ONCE per matrix as Taylor
// expansions about nodes x_g = g h / 4 covering that range:
//     S(x_g + t h) = sum_m C[g][m] t^m,   C[g][m] = sum_i c_m(u_i) exp(-u_i^2 / 2),   u_i = (x_g - lam_i) / h,
//     c_0 = 1, c_1 = -u, c_{m+1} = (-u c_m - c_{m-1}) / (m + 1)     ((-1)^m He_m(u) / m!, probabilists' Hermite)
// and every objective evaluation reads S and S' = -(1/h^2) sum_i E_i (x - lam_i) off the nearest node (|t| <= 1/8) with two
// Horner recurrences of degree 12 -- truncation below 1e-16 of max S (scratch/kde_taylor_proto.py) -- instead of summing
// N kernel terms per grid point.  Cost: one exp + ~40 FP64 instructions per (node, eigenvalue) once (48 nodes at
// h = 0.25, q = 2), then ~30 per grid point and evaluation; the direct form costs ~4 N per grid point and evaluation.
#define MP_DEG 12
#define MP_NODE_DIV 4        // nodes per bandwidth
#define MP_MAX_NODES 256     // more (a bandwidth below ~0.05 at q = 2): direct evaluation

__host__ __device__ inline int mp_node_count(double q, double bandwidth) {
    const double sq = sqrt(1.0 / q);
    const double top = (1.0 + sq) * (1.0 + sq);
    const double cnt = floor(top * MP_NODE_DIV / bandwidth) + 2.0;
    return (cnt >= 2.0 && cnt <= (double)MP_MAX_NODES) ? (int)cnt : 0;
}

// All MP_THREADS threads.  One warp per node and pass, lanes over the eigenvalues, fixed summation order.
__device__ void mp_build_nodes(const MPCtx& c, double* coef) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double invh = 1.0 / c.h;
    for (int g = wid; g < c.nodes; g += MP_THREADS / 32) {
        const double xg = (double)g * c.delta;
        double acc[MP_DEG + 1];
#pragma unroll
        for (int m = 0; m <= MP_DEG; ++m) acc[m] = 0.0;
        for (int i = lane; i < c.N; i += 32) {
            const double u = (xg - c.lam[i]) * invh;
            const double hu2 = 0.5 * u * u;
            if (hu2 > 745.0) continue;          // exp underflows to zero
            const double E = exp(-hu2);
            double cm1 = 0.0, cm = E;           // c_m(u) E
#pragma unroll
            for (int m = 0; m <= MP_DEG; ++m) {
                acc[m] += cm;
                const double nx = (-u * cm - cm1) * (1.0 / (double)(m + 1));
                cm1 = cm;
                cm = nx;
            }
        }
#pragma unroll
        for (int m = 0; m <= MP_DEG; ++m) {
            double v = acc[m];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) coef[g * (MP_DEG + 1) + m] = v;
        }
    }
    __syncthreads();
}

// Objective and analytic derivative at v.  The Gaussian kernel sums over a UNIFORM grid obey a two-multiply recurrence:
//   g_k = exp(-(x_k - lam)^2 / 2h^2),  g_{k+1} = g_k r_k,  r_{k+1} = r_k c,  c = exp(-dx^2 / h^2),
// so a thread that owns MP_KB consecutive grid points pays two exp per eigenvalue instead of MP_KB.  Thread layout:
// MP_IG threads share one block of MP_KB grid points and split the eigenvalues between them (reduced with shuffles).
// 1000 points = 32 blocks of 32 x 8 eigenvalue groups = 256 threads (blocks of 16 with 512 threads paid twice the exps).
#define MP_KB 32
#define MP_IG 8
__device__ void mp_eval(const MPCtx& c, double v, double& f, double& g) {
    const double sq = sqrt(1.0 / c.q);
    const double lo = v * (1.0 - sq) * (1.0 - sq), hi = v * (1.0 + sq) * (1.0 + sq);
    const double step = (hi - lo) / (MP_PTS - 1);
    const int kb = threadIdx.x / MP_IG, ig = threadIdx.x % MP_IG;
    const int k0 = kb * MP_KB;
    double fs = 0.0, gs = 0.0;
    if (c.nodes > 0) {
        for (int k = threadIdx.x; k < MP_PTS; k += MP_THREADS) {
            // numpy.linspace: arange * step + start (two roundings, no FMA), last point forced to the stop value
            const double x = (k == MP_PTS - 1) ? hi : __dadd_rn(__dmul_rn((double)k, step), lo);
            int g = (int)rint(x * c.inv_delta);
            g = min(max(g, 0), c.nodes - 1);
            const double t = (x - (double)g * c.delta) / c.h;
            const double* cg = c.coef + g * (MP_DEG + 1);
            double s0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int m = MP_DEG; m >= 0; --m) {
                d1 = fma(d1, t, s0);
                s0 = fma(s0, t, cg[m]);
            }
            const double kd = -c.h * d1;                                  // sum_i E_i (x - lam_i)
            const double rad = fmax((hi - x) * (x - lo), 0.0);
            const double pdf = c.q / (2.0 * M_PI * v * x) * sqrt(rad);
            const double kde = c.cnorm * s0;
            const double dkde = -c.cnorm * kd * c.invh2 * (x / v);
            const double r = kde - pdf;
            fs = fma(r, r, fs);
            gs = fma(2.0 * r, dkde + pdf / v, gs);
        }
    } else {   // all threads run the loop (the shuffles need full warps); the last group's points are past the grid
        double s0[MP_KB], sl[MP_KB];
#pragma unroll
        for (int j = 0; j < MP_KB; ++j) s0[j] = sl[j] = 0.0;
        const double x0 = __dadd_rn(__dmul_rn((double)k0, step), lo);
        const double cstep = exp(-step * step * c.invh2);
        for (int i = ig; i < c.N; i += MP_IG) {
            const double lam = c.lam[i];
            const double d0 = x0 - lam;
            double gk = exp(-d0 * d0 * c.inv2h2);
            if (gk == 0.0) continue;                       // the whole block is below 1e-323 for this eigenvalue
            double rk = exp(-(2.0 * d0 * step + step * step) * c.inv2h2);
            if (!(rk < 1e150)) {                           // pathological bandwidth: direct evaluation
#pragma unroll
                for (int j = 0; j < MP_KB; ++j) {
                    const double dj = d0 + j * step;
                    const double E = exp(-dj * dj * c.inv2h2);
                    s0[j] += E;
                    sl[j] = fma(E, lam, sl[j]);
                }
                continue;
            }
#pragma unroll
            for (int j = 0; j < MP_KB; ++j) {
                s0[j] += gk;
                sl[j] = fma(gk, lam, sl[j]);
                gk *= rk;
                rk *= cstep;
            }
        }
        // reduce over the MP_IG threads of the group (adjacent lanes)
#pragma unroll
        for (int j = 0; j < MP_KB; ++j) {
#pragma unroll
            for (int o = MP_IG / 2; o > 0; o >>= 1) {
                s0[j] += __shfl_xor_sync(0xffffffffu, s0[j], o);
                sl[j] += __shfl_xor_sync(0xffffffffu, sl[j], o);
            }
        }
        // each thread of the group finishes MP_KB / MP_IG of the block's grid points
#pragma unroll
        for (int j = 0; j < MP_KB; ++j) {
            if ((j % MP_IG) == ig) {
                const int k = k0 + j;
                if (k < MP_PTS) {
                    // numpy.linspace: arange * step + start (two roundings, no FMA), last point forced to the stop value
                    const double x = (k == MP_PTS - 1) ? hi : __dadd_rn(__dmul_rn((double)k, step), lo);
                    const double rad = fmax((hi - x) * (x - lo), 0.0);
                    const double pdf = c.q / (2.0 * M_PI * v * x) * sqrt(rad);
                    const double kde = c.cnorm * s0[j];
                    const double kd = x * s0[j] - sl[j];                 // sum_i E_i (x - lam_i)
                    const double dkde = -c.cnorm * kd * c.invh2 * (x / v);
                    const double r = kde - pdf;
                    fs = fma(r, r, fs);
                    gs = fma(2.0 * r, dkde + pdf / v, gs);
                }
            }
        }
    }
    f = block_sum(fs, c.red);
    g = block_sum(gs, c.red);
}

__device__ void dcstep(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp,
                       double fp, double dp, bool& brackt, double stpmin, double stpmax) {
    const double sgnd = dp * (dx / fabs(dx));
    double stpf, stpc, stpq, theta, s, gamma, p, q, r;
    if (fp > fx) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
        gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp < stx) gamma = -gamma;
        p = (gamma - dx) + theta;
        q = ((gamma - dx) + gamma) + dp;
        r = p / q;
        stpc = stx + r * (stp - stx);
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
        stpf = (fabs(stpc - stx) < fabs(stpq - stx)) ? stpc : stpc + (stpq - stpc) / 2.0;
        brackt = true;
    } else if (sgnd < 0.0) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
        gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = ((gamma - dp) + gamma) + dx;
        r = p / q;
        stpc = stp + r * (stx - stp);
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
        brackt = true;
    } else if (fabs(dp) < fabs(dx)) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
        gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = (gamma + (dx - dp)) + gamma;
        r = p / q;
        if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
        else if (stp > stx) stpc = stpmax;
        else stpc = stpmin;
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (brackt) {
            stpf = (fabs(stpc - stp) < fabs(stpq - stp)) ? stpc : stpq;
            if (stp > stx) stpf = fmin(stp + 0.66 * (sty - stp), stpf);
            else stpf = fmax(stp + 0.66 * (sty - stp), stpf);
        } else {
            stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
            stpf = fmin(stpmax, stpf);
            stpf = fmax(stpmin, stpf);
        }
    } else {
        if (brackt) {
            theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
            s = fmax(fabs(theta), fmax(fabs(dy), fabs(dp)));
            gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
            if (stp > sty) gamma = -gamma;
            p = (gamma - dp) + theta;
            q = ((gamma - dp) + gamma) + dy;
            r = p / q;
            stpc = stp + r * (sty - stp);
            stpf = stpc;
        } else if (stp > stx) stpf = stpmax;
        else stpf = stpmin;
    }
    if (fp > fx) {
        sty = stp; fy = fp; dy = dp;
    } else {
        if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
        stx = stp; fx = fp; dx = dp;
    }
    stp = stpf;
}

// All threads execute the scalar control flow redundantly on block-uniform values. Returns true on "success".
__device__ bool mp_lbfgsb(const MPCtx& c, double& xout) {
    const double l = 1e-5, u = 1.0 - 1e-5, pgtol = 1e-5, factr = 1e7, eps = 2.220446049250313e-16;
    const double ftol = 1e-3, gtol = 0.9, xtol = 0.1;
    double x = 0.5, f, g;
    mp_eval(c, x, f, g);
    xout = x;
    if (!isfinite(f) || !isfinite(g)) return false;
    auto projg = [&](double xx, double gg) { return gg < 0.0 ? fmax(xx - u, gg) : fmin(xx - l, gg); };
    if (fabs(projg(x, g)) <= pgtol) return true;
    double theta = 1.0;
    bool have_pair = false;
    int it = 0, nfev = 1;
    while (true) {
        const double z = fmin(fmax(x - g / theta, l), u);
        double d = z - x;
        double stpmx;
        if (it == 0) stpmx = 1.0;
        else {
            stpmx = 1e10;
            if (d < 0.0) stpmx = (l - x) / d;
            else if (d > 0.0) stpmx = (u - x) / d;
        }
        double stp = 1.0;
        const double xold = x, fold = f, gold = g;
        double gd = g * d;
        const double gdold = gd;
        if (gd >= 0.0) {  // not a descent direction: drop the curvature pair and retry, or give up
            if (!have_pair) { xout = x; return false; }
            have_pair = false; theta = 1.0;
            continue;
        }
        bool brackt = false;
        int stage = 1;
        const double finit = f, ginit = gd, gtest = ftol * ginit;
        double width = stpmx, width1 = 2.0 * width;
        double stx = 0.0, fx = finit, gx = ginit, sty = 0.0, fy = finit, gy = ginit, stmin = 0.0, stmax = stp + 4.0 * stp;
        int nls = 0;
        bool ok = true;
        while (true) {
            x = (stp == 1.0) ? z : xold + stp * d;
            mp_eval(c, x, f, g);
            ++nfev; ++nls;
#ifdef MCOSB_DEBUG_MPFIT
            if (threadIdx.x == 0) printf("ls it=%d nls=%d stp=%.17g x=%.17g f=%.17g g=%.17g finit=%.17g ginit=%.17g\n", it, nls, stp, x, f, g, finit, ginit);
#endif
            if (!isfinite(f) || !isfinite(g)) { ok = false; break; }
            gd = g * d;
            const double ftest = finit + stp * gtest;
            if (stage == 1 && f <= ftest && gd >= 0.0) stage = 2;
            bool done = false;
            if (brackt && (stp <= stmin || stp >= stmax)) done = true;
            if (brackt && stmax - stmin <= xtol * stmax) done = true;
            if (stp == stpmx && f <= ftest && gd <= gtest) done = true;
            if (stp == 0.0 && (f > ftest || gd >= gtest)) done = true;
            if (f <= ftest && fabs(gd) <= gtol * (-ginit)) done = true;
            if (done) break;
            if (nls >= 20) { ok = false; break; }
            if (stage == 1 && f <= fx && f > ftest) {
                double fm = f - stp * gtest, fxm = fx - stx * gtest, fym = fy - sty * gtest;
                double gm = gd - gtest, gxm = gx - gtest, gym = gy - gtest;
                dcstep(stx, fxm, gxm, sty, fym, gym, stp, fm, gm, brackt, stmin, stmax);
                fx = fxm + stx * gtest; fy = fym + sty * gtest; gx = gxm + gtest; gy = gym + gtest;
            } else {
                dcstep(stx, fx, gx, sty, fy, gy, stp, f, gd, brackt, stmin, stmax);
            }
            if (brackt) {
                if (fabs(sty - stx) >= 0.66 * width1) stp = stx + 0.5 * (sty - stx);
                width1 = width;
                width = fabs(sty - stx);
            }
            if (brackt) { stmin = fmin(stx, sty); stmax = fmax(stx, sty); }
            else { stmin = stp + 1.1 * (stp - stx); stmax = stp + 4.0 * (stp - stx); }
            stp = fmax(stp, 0.0);
            stp = fmin(stp, stpmx);
            if ((brackt && (stp <= stmin || stp >= stmax)) || (brackt && stmax - stmin <= xtol * stmax)) stp = stx;
        }
        if (!ok) {
            x = xold; f = fold; g = gold;
            if (!have_pair) { xout = x; return false; }
            have_pair = false; theta = 1.0;
            continue;
        }
        ++it;
        xout = x;
        if (fabs(projg(x, g)) <= pgtol) return true;
        const double ddum0 = fmax(fabs(fold), fmax(fabs(f), 1.0));
        if (fold - f <= eps * factr * ddum0) return true;
        if (it >= 15000 || nfev >= 15000) return false;
        const double r = g - gold;
        double dr, ddum;
        if (stp == 1.0) { dr = gd - gdold; ddum = -gdold; }
        else { dr = (gd - gdold) * stp; ddum = -gdold * stp; }
        if (dr > eps * ddum) { theta = r * r / dr; have_pair = true; }
    }
}

__global__ void __launch_bounds__(MP_THREADS) mpfit_kernel(const double* __restrict__ lamall, int N, double q,
                                                           double bandwidth, double* __restrict__ var_out,
                                                           int* __restrict__ nf_out, int* __restrict__ status,
                                                           int use_nodes) {
    extern __shared__ double sm[];
    __shared__ double red[32];
    __shared__ int s_cnt;
    const int s = blockIdx.x;
    const double* lam = lamall + (size_t)s * N;
    bool finite = true;
    for (int i = threadIdx.x; i < N; i += MP_THREADS) {
        sm[i] = lam[i];
        if (!isfinite(sm[i])) finite = false;
    }
    if (threadIdx.x == 0) s_cnt = 0;
    const int any_bad = __syncthreads_or(!finite);
    MPCtx c;
    c.lam = sm; c.N = N; c.q = q; c.red = red;
    c.inv2h2 = 1.0 / (2.0 * bandwidth * bandwidth);
    c.invh2 = 1.0 / (bandwidth * bandwidth);
    c.cnorm = 1.0 / (N * bandwidth * sqrt(2.0 * M_PI));
    c.h = bandwidth;
    c.delta = bandwidth / MP_NODE_DIV;
    c.inv_delta = MP_NODE_DIV / bandwidth;
    c.nodes = use_nodes ? mp_node_count(q, bandwidth) : 0;
    c.coef = sm + N;
    if (c.nodes > 0 && !any_bad) mp_build_nodes(c, sm + N);
    double v = 1.0;
    bool ok = false;
    if (!any_bad) ok = mp_lbfgsb(c, v);
    if (!ok) v = 1.0;  // covariance_transformer.py:127-130
    const double sq = sqrt(1.0 / q);
    const double lam_plus = v * (1.0 + sq) * (1.0 + sq);
    int cnt = 0;
    for (int i = threadIdx.x; i < N; i += MP_THREADS) cnt += (sm[i] >= lam_plus);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (var_out) var_out[s] = v;
        nf_out[s] = s_cnt;
        if (status) {
            int st = 0;
            if (!ok) st |= MCOSB_ST_MPFIT_FAIL;
            if (any_bad) st |= MCOSB_ST_NONFINITE;
            if (st) atomicOr(&status[s], st);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// D5: signal eigenvectors + clipped reconstruction, one CTA per matrix.
//   for chunks of EV_VCH eigenvalues (largest first):
//     - inverse iteration on (T - lam I) (tridiagonal LU with partial pivoting, one thread per vector, 3 sweeps);
//     - back-transformation y <- H_0 ... H_{N-3} y with the stored reflectors, one warp per vector;
//     - cov[i][j] (+)= sum_s (lam_s - lbar) y_s[i] y_s[j]            (first chunk also lays down lbar I)
//   then cov <- clip(R1 / sqrt(r_i r_j)) sigma_i sigma_j                 (cov_to_corr + corr_to_cov, :207, :96)
// ---------------------------------------------------------------------------------------------------------------
#define VC_THREADS 512
#define VC_VCH 16  // vectors per chunk (= warps per CTA)

__global__ void __launch_bounds__(VC_THREADS) eigvec_kernel(const double* __restrict__ Aall, int N,
                                                            const double* __restrict__ dall,
                                                            const double* __restrict__ eall,
                                                            const double* __restrict__ tauall,
                                                            const double* __restrict__ lamall,
                                                            const int* __restrict__ nfall,
                                                            const double* __restrict__ sigmaall,
                                                            double* __restrict__ workall,  // [S][5][N][VC_VCH]
                                                            unsigned char* __restrict__ pivall,  // [S][N][VC_VCH]
                                                            double* __restrict__ covall, int* __restrict__ status,
                                                            int skip_upto, const double* __restrict__ Q2all) {
    if (nfall[blockIdx.x] <= skip_upto) return;  // handled by the wide path (eigvec3.cu)
    const int off = Q2all ? SBR_B : 1;           // two-stage reflectors: Q2 (sb2st.cu) first, then the panels of sy2sb.cu
    extern __shared__ double sm[];
    double* sd = sm;                  // [N]
    double* se = sd + N;              // [N]
    double* ys = se + N;              // [VC_VCH][N] vectors
    double* coef = ys + (size_t)VC_VCH * N;  // [VC_VCH]
    __shared__ double red[32];
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double* A = Aall + (size_t)s * N * N;
    const double* lam = lamall + (size_t)s * N;
    const double* tau = tauall + (size_t)s * N;
    const double* sigma = sigmaall + (size_t)s * N;
    double* cov = covall + (size_t)s * N * N;
    double* work = workall + (size_t)s * 5 * N * VC_VCH;
    unsigned char* piv = pivall + (size_t)s * N * VC_VCH;
    const int nf = nfall[s];

    double tnorm = 0.0;
    for (int i = tid; i < N; i += VC_THREADS) {
        sd[i] = dall[(size_t)s * N + i];
        se[i] = eall[(size_t)s * N + i];
        tnorm = fmax(tnorm, fabs(sd[i]) + 2.0 * fabs(se[i]));
    }
    // noise mean: sum of the eigenvalues from n_facts on (covariance_transformer.py:204)
    double part = 0.0;
    for (int i = nf + tid; i < N; i += VC_THREADS) part += lam[i];
    const double lbar = (nf < N) ? block_sum(part, red) / (double)(N - nf) : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tnorm = fmax(tnorm, __shfl_xor_sync(0xffffffffu, tnorm, o));
    __syncthreads();
    if (lane == 0) red[wid] = tnorm;
    __syncthreads();
    tnorm = red[0];
    for (int i = 1; i < VC_THREADS / 32; ++i) tnorm = fmax(tnorm, red[i]);
    __syncthreads();
    const double tiny_piv = 2.2e-16 * fmax(tnorm, 1e-300);
    bool bad_vec = false;

    for (int c0 = 0; c0 < nf || c0 == 0; c0 += VC_VCH) {
        const int nv = max(0, min(VC_VCH, nf - c0));
        // ---- inverse iteration: thread j < nv of warp 0 owns vector c0 + j ------------------------------------------
        if (tid < nv) {
            const int j = tid;
            double lj = lam[c0 + j];
            // separate numerically repeated eigenvalues a little so the iterates differ (dstein does the same)
            if (c0 + j > 0 && lam[c0 + j - 1] - lj < 10.0 * tiny_piv) lj -= 10.0 * tiny_piv * (double)((c0 + j) % 7 + 1);
            double* dl = work + 0 * (size_t)N * VC_VCH;
            double* dd = work + 1 * (size_t)N * VC_VCH;
            double* du = work + 2 * (size_t)N * VC_VCH;
            double* du2 = work + 3 * (size_t)N * VC_VCH;
            double* y = work + 4 * (size_t)N * VC_VCH;
#define IX(i) ((size_t)(i) * VC_VCH + j)
            // LU with partial pivoting (dgttrf recurrences), streaming the unfactored part from shared memory
            double dcur = sd[0] - lj;                    // U diagonal candidate of row i
            double ucur = (N > 1) ? se[0] : 0.0;         // super-diagonal of row i
            for (int i = 0; i + 1 < N; ++i) {
                const double sub = se[i];                // sub-diagonal entry (i+1, i)
                double dnext = sd[i + 1] - lj;           // diagonal of row i+1
                double unext = (i + 2 < N) ? se[i + 1] : 0.0;  // super-diagonal of row i+1
                if (fabs(dcur) >= fabs(sub)) {
                    if (dcur == 0.0) dcur = tiny_piv;
                    const double fact = sub / dcur;
                    dl[IX(i)] = fact; dd[IX(i)] = dcur; du[IX(i)] = ucur; du2[IX(i)] = 0.0; piv[IX(i)] = 0;
                    dcur = dnext - fact * ucur;
                    ucur = unext;
                } else {
                    const double fact = dcur / sub;
                    dl[IX(i)] = fact; dd[IX(i)] = sub; du[IX(i)] = dnext; du2[IX(i)] = unext; piv[IX(i)] = 1;
                    dcur = ucur - fact * dnext;
                    ucur = -fact * unext;
                }
            }
            if (dcur == 0.0) dcur = tiny_piv;
            dd[IX(N - 1)] = dcur;
            // start vector: deterministic pseudo-random in (-1, 1)
            for (int i = 0; i < N; ++i) {
                uint32_t h = (uint32_t)(i * 2654435761u) ^ (uint32_t)((c0 + j + 1) * 40503u);
                h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
                y[IX(i)] = ((double)(h & 0xffffff) + 0.5) / 8388608.0 - 1.0;
            }
            double scale = 1.0;
            for (int sweep = 0; sweep < 3; ++sweep) {
                // forward substitution with the row interchanges
                double bi = y[IX(0)] * scale;
                for (int i = 0; i + 1 < N; ++i) {
                    double bn = y[IX(i + 1)] * scale;
                    if (piv[IX(i)] == 0) {
                        y[IX(i)] = bi;
                        bi = bn - dl[IX(i)] * bi;
                    } else {
                        y[IX(i)] = bn;
                        bi = bi - dl[IX(i)] * bn;
                    }
                }
                // back substitution
                double x1 = bi / dd[IX(N - 1)], x2 = 0.0, mx = fabs(x1);
                y[IX(N - 1)] = x1;
                for (int i = N - 2; i >= 0; --i) {
                    double xi = (y[IX(i)] - du[IX(i)] * x1 - du2[IX(i)] * x2) / dd[IX(i)];
                    y[IX(i)] = xi;
                    x2 = x1; x1 = xi;
                    mx = fmax(mx, fabs(xi));
                }
                scale = 1.0 / mx;
            }
            double nrm = 0.0;
            for (int i = 0; i < N; ++i) { double t = y[IX(i)] * scale; nrm = fma(t, t, nrm); }
            const double sc = scale / sqrt(nrm);
            if (!isfinite(sc)) bad_vec = true;
            for (int i = 0; i < N; ++i) ys[(size_t)j * N + i] = y[IX(i)] * sc;
            coef[j] = lam[c0 + j] - lbar;
#undef IX
        }
        __syncthreads();
        // ---- back-transformation, one warp per vector ------------------------------------------------------------------
        if (wid < nv) {
            double* yv = ys + (size_t)wid * N;
            if (Q2all) {
                // y <- Q2 y: the reflectors of a sweep act on disjoint windows of 8 rows, one lane each
                const double* Q2 = Q2all + (size_t)s * N * N;
                for (int sw = N - 3; sw >= 0; --sw) {
                    const double* q = Q2 + (size_t)sw * N;
                    for (int lo = sw + 1 + 8 * lane; N - lo >= 2; lo += 8 * 32) {
                        const double tq = q[lo - sw - 1];
                        double dot = yv[lo];
                        for (int m = 1; m < 8 && lo + m < N; ++m) dot = fma(q[lo - sw - 1 + m], yv[lo + m], dot);
                        dot *= tq;
                        yv[lo] -= dot;
                        for (int m = 1; m < 8 && lo + m < N; ++m) yv[lo + m] = fma(-dot, q[lo - sw - 1 + m], yv[lo + m]);
                    }
                    __syncwarp();
                }
            }
            for (int k = N - 2 - off; k >= 0; --k) {
                const double tk = tau[k];
                if (tk == 0.0) continue;
                const double* u = A + (size_t)k * N;  // u[k+off] = 1 implicit, u[j] = A[k][j] for j > k+off
                double dot = 0.0;
                for (int jj = k + off + lane; jj < N; jj += 32) {
                    double uj = (jj == k + off) ? 1.0 : u[jj];
                    dot = fma(uj, yv[jj], dot);
                }
                dot = warp_sum(dot) * tk;
                for (int jj = k + off + lane; jj < N; jj += 32) {
                    double uj = (jj == k + off) ? 1.0 : u[jj];
                    yv[jj] = fma(-dot, uj, yv[jj]);
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // ---- low-rank accumulation into cov ----------------------------------------------------------------------------
        for (size_t idx = tid; idx < (size_t)N * N; idx += VC_THREADS) {
            const int i = (int)(idx / N), jx = (int)(idx % N);
            double acc = (c0 == 0) ? ((i == jx) ? lbar : 0.0) : cov[idx];
            for (int v = 0; v < nv; ++v) acc = fma(coef[v] * ys[(size_t)v * N + i], ys[(size_t)v * N + jx], acc);
            cov[idx] = acc;
        }
        __syncthreads();
    }
    // ---- renormalise to unit diagonal, clip, rescale by the input standard deviations ------------------------------
    for (int i = tid; i < N; i += VC_THREADS) sd[i] = sqrt(cov[(size_t)i * N + i]);
    __syncthreads();
    for (size_t idx = tid; idx < (size_t)N * N; idx += VC_THREADS) {
        const int i = (int)(idx / N), jx = (int)(idx % N);
        double r = cov[idx] / (sd[i] * sd[jx]);
        r = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
        cov[idx] = r * (sigma[i] * sigma[jx]);
    }
    if (status && __syncthreads_or(bad_vec) && tid == 0) atomicOr(&status[s], MCOSB_ST_EIGVEC);
}

// Symmetric A[S][N][N] (lower triangle read) -> tridiagonal (d, e), reflectors left in A / tau (/ *q2).
int mcosb_dev_tridiagonalize(mcosb_ctx* ctx, int S, double* A, double* d, double* e, double* tau, const double** q2) {
    const int N = ctx->N;
    const size_t n = (size_t)N;
    *q2 = nullptr;
    const bool sbr = mcosb_sbr_supported(ctx) && (ctx->eig_method == 2 || (ctx->eig_method == 0 && N >= SBR_MIN_N));
    if (ctx->eig_method == 2 && !sbr)
        return mcosb_fail(ctx, MCOSB_ERR_ARG, "two-stage tridiagonalisation needs 10 <= n_assets <= 1024");
    if (sbr) {
        // two-stage: dense -> band (sy2sb.cu, DMMA) -> tridiagonal (sb2st.cu, bulge chasing in shared memory)
        double* Q2;
        MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_Q2, (size_t)S * n * n, &Q2));
        int r1 = mcosb_launch_stage1(ctx, S, A, tau);
        if (r1 < 0) return r1;
        int r2 = r1 == 1 ? mcosb_launch_sb2st(ctx, S, A, d, e, Q2) : 0;
        if (r2 < 0) return r2;
        if (r1 != 1 || r2 != 1) return mcosb_fail(ctx, MCOSB_ERR_STATE, "two-stage tridiagonalisation failed to launch");
        *q2 = Q2;
        return MCOSB_OK;
    }
    int r2 = mcosb_launch_tridiag2(ctx, S, A, d, e, tau);  // blocked, lower triangle only (tridiag2.cu)
    if (r2 < 0) return r2;
    if (r2 == 1) {
        MCOSB_LAUNCHED(ctx, MK_TRIDIAG2);
    } else {
        size_t smem_td = (4 * n + TD_WARPS * 32) * sizeof(double);
        MCOSB_CUDA(ctx, cudaFuncSetAttribute(tridiag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_td));
        tridiag_kernel<<<S, TD_THREADS, smem_td, ctx->stream>>>(A, N, d, e, tau);
        MCOSB_LAUNCHED(ctx, MK_TRIDIAG);
    }
    return MCOSB_OK;
}

// corr = cov_to_corr(cov) -> tridiagonalisation (reflectors left in A) -> all eigenvalues, descending.
// Shared by the DeNoiser and the Detone transformer.
int mcosb_dev_eig_prepare(mcosb_ctx* ctx, int S, const double* dcov, double* A, double* sigma, double* d, double* e,
                          double* tau, double* lam, const double** q2, const double* lwcoef) {
    const int N = ctx->N;
    const size_t n = (size_t)N;
    corr_kernel<<<dim3((N + CORR_ROWS - 1) / CORR_ROWS, S), 256, n * sizeof(double), ctx->stream>>>(dcov, N, A, sigma, lwcoef);
    MCOSB_LAUNCHED(ctx, MK_CORR);
    MCOSB_TRY(mcosb_dev_tridiagonalize(ctx, S, A, d, e, tau, q2));
    size_t smem_ev = 2 * n * sizeof(double) + (n + 1) * sizeof(int) + n * (sizeof(double) + sizeof(int));   // (d, e^2), grid counts, probes
    MCOSB_CUDA(ctx, cudaFuncSetAttribute(eigvals_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ev));
    eigvals_kernel<<<S, EV_THREADS, smem_ev, ctx->stream>>>(d, e, N, lam);
    MCOSB_LAUNCHED(ctx, MK_EIGVALS);
    return MCOSB_OK;
}

// Diagnostic entry point: the tridiagonal form of symmetric matrices (first half of np.linalg.eigh,
// covariance_transformer.py:105), by the selected method.  The work copy of A (reflectors / band) comes back in A_work.
extern "C" int mcosb_tridiagonalize(mcosb_ctx* ctx, int S, const double* sym, double* d_out, double* e_out,
                                    double* a_work) {
    MCOSB_CHECK_CTX(ctx);
    if (S < 0 || !sym || !d_out || !e_out) return mcosb_fail(ctx, MCOSB_ERR_ARG, "bad argument to mcosb_tridiagonalize");
    if (S == 0) return MCOSB_OK;
    const size_t n = ctx->N;
    Staging st(ctx);
    const double* dsym;
    double *dd, *de, *dwork, *A, *tau;
    MCOSB_TRY(st.in(sym, (size_t)S * n * n, SL_STAGE_IN0, &dsym));
    MCOSB_TRY(st.out(d_out, (size_t)S * n, SL_STAGE_OUT0, &dd));
    MCOSB_TRY(st.out(e_out, (size_t)S * n, SL_STAGE_OUT1, &de));
    MCOSB_TRY(st.out(a_work, (size_t)S * n * n, SL_STAGE_OUT2, &dwork));
    if (dwork) A = dwork;
    else MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_A, (size_t)S * n * n, &A));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_TAU, (size_t)S * n, &tau));
    MCOSB_CUDA(ctx, cudaMemcpyAsync(A, dsym, (size_t)S * n * n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    MCOSB_PROFILE_MARK(ctx);   // the copy is not part of the first kernel's time
    const double* q2 = nullptr;
    MCOSB_TRY(mcosb_dev_tridiagonalize(ctx, S, A, dd, de, tau, &q2));
    return st.finish();
}

extern "C" int mcosb_set_eig_method(mcosb_ctx* ctx, int method) {
    MCOSB_CHECK_CTX(ctx);
    if (method < 0 || method > 2) return mcosb_fail(ctx, MCOSB_ERR_ARG, "eig method must be 0 (auto), 1 or 2");
    ctx->eig_method = method;
    return MCOSB_OK;
}

int mcosb_dev_denoise(mcosb_ctx* ctx, int S, double* dcov, int n_obs, double bandwidth, double* deig, double* dvar,
                      int* dnf, int* dstatus, const double* lwcoef) {
    if (S == 0) return MCOSB_OK;
    if (n_obs < 1 || !(bandwidth > 0.0)) return mcosb_fail(ctx, MCOSB_ERR_ARG, "bad n_obs/bandwidth");
    const int N = ctx->N;
    const size_t n = (size_t)N;
    double *A, *sigma, *d, *e, *tau, *lam;
    int* nf;
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_A, (size_t)S * n * n, &A));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_SIGMA, (size_t)S * n, &sigma));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_D, (size_t)S * n, &d));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_E, (size_t)S * n, &e));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_TAU, (size_t)S * n, &tau));
    lam = deig;
    if (!lam) MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_LAM, (size_t)S * n, &lam));
    nf = dnf;
    if (!nf) MCOSB_TRY(mcosb_scratch_t(ctx, SL_DN_NF, (size_t)S, &nf));
    const double* Q2 = nullptr;
    MCOSB_TRY(mcosb_dev_eig_prepare(ctx, S, dcov, A, sigma, d, e, tau, lam, &Q2, lwcoef));

    const double q = (double)n_obs / (double)N;
    static const bool mp_direct = std::getenv("MCOSB_MPFIT_DIRECT") != nullptr;   // A/B runs: sum the kernel terms per grid point
    const int mp_nodes = mp_direct ? 0 : mp_node_count(q, bandwidth);
    size_t smem_mp = (n + (size_t)mp_nodes * (MP_DEG + 1)) * sizeof(double);
    MCOSB_CUDA(ctx, cudaFuncSetAttribute(mpfit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mp));
    mpfit_kernel<<<S, MP_THREADS, smem_mp, ctx->stream>>>(lam, N, q, bandwidth, dvar, nf, dstatus, mp_nodes > 0);
    MCOSB_LAUNCHED(ctx, MK_MPFIT);

    // wide path (eigvec3.cu) for simulations with at most 16 signal eigenvalues; eigvec_kernel takes the others
    int rv = mcosb_launch_eigvec3(ctx, S, A, d, e, tau, lam, nf, sigma, dcov, dstatus, Q2);
    if (rv < 0) return rv;
    {
        double* work2;
        unsigned char* piv2;
        MCOSB_TRY(mcosb_scratch_t(ctx, SL_MK_L, (size_t)S * 5 * n * VC_VCH, &work2));
        MCOSB_TRY(mcosb_scratch_t(ctx, SL_MK_W, (size_t)S * n * VC_VCH, &piv2));
        size_t smem_vc = (2 * n + (size_t)VC_VCH * n + VC_VCH) * sizeof(double);
        if (smem_vc > (size_t)ctx->smem_optin) return mcosb_fail(ctx, MCOSB_ERR_ARG, "n_assets too large for eigvec_kernel");
        MCOSB_CUDA(ctx, cudaFuncSetAttribute(eigvec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_vc));
        eigvec_kernel<<<S, VC_THREADS, smem_vc, ctx->stream>>>(A, N, d, e, tau, lam, nf, sigma, work2, piv2, dcov, dstatus,
                                                              rv == 1 ? 16 : -1, Q2);
        MCOSB_LAUNCHED(ctx, MK_EIGVEC);
    }
    return MCOSB_OK;
}

extern "C" int mcosb_denoise(mcosb_ctx* ctx, int S, double* cov_inout, int n_obs, double bandwidth, double* eigvals,
                             double* var, int* n_facts, int* status) {
    MCOSB_CHECK_CTX(ctx);
    if (S < 0 || !cov_inout) return mcosb_fail(ctx, MCOSB_ERR_ARG, "bad argument to mcosb_denoise");
    const size_t N = ctx->N;
    Staging st(ctx);
    double *dcov, *deig, *dvar;
    int *dnf, *dst;
    MCOSB_TRY(st.inout(cov_inout, (size_t)S * N * N, SL_STAGE_IN0, &dcov));
    MCOSB_TRY(st.out(eigvals, (size_t)S * N, SL_STAGE_OUT0, &deig));
    MCOSB_TRY(st.out(var, (size_t)S, SL_STAGE_OUT1, &dvar));
    MCOSB_TRY(st.out(n_facts, (size_t)S, SL_STAGE_OUT2, &dnf));
    MCOSB_TRY(st.out(status, (size_t)S, SL_STAGE_OUT3, &dst));
    if (dst) MCOSB_CUDA(ctx, cudaMemsetAsync(dst, 0, (size_t)S * sizeof(int), ctx->stream));
    MCOSB_TRY(mcosb_dev_denoise(ctx, S, dcov, n_obs, bandwidth, deig, dvar, dnf, dst));
    return st.finish();
}
