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
#include <cstdlib>
#include "common.cuh"
#include "dmma.cuh"
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
    // Every step of the two substitutions needs operands from global memory (the interleaved factor arrays) that do NOT depend
    // on the recurrence, but the compiler keeps each load behind the previous step's store: a thread paid a DRAM round trip per
    // step (~3300 cycles, all 16 N threads of a wave resident and waiting).  The operands of IV_PF steps are loaded together
    // ahead of the dependent chain; reads and writes of y keep their order per element (y[i + 1] is read before the step
    // that overwrites it).
    constexpr int IV_PF = 8;
    double scale = 1.0;
    for (int sweep = 0; sweep < 3; ++sweep) {
        double bi = y[0] * scale;
        int i = 0;
        for (; i + IV_PF < N; i += IV_PF) {
            double bnv[IV_PF], fv[IV_PF];
            unsigned char pvv[IV_PF];
#pragma unroll
            for (int q = 0; q < IV_PF; ++q) {
                bnv[q] = y[(size_t)(i + q + 1) * G];
                fv[q] = dl[(size_t)(i + q) * G];
                pvv[q] = pv[(size_t)(i + q) * G];
            }
#pragma unroll
            for (int q = 0; q < IV_PF; ++q) {
                const double bn = bnv[q] * scale;
                if (pvv[q] == 0) {
                    y[(size_t)(i + q) * G] = bi;
                    bi = bn - fv[q] * bi;
                } else {
                    y[(size_t)(i + q) * G] = bn;
                    bi = bi - fv[q] * bn;
                }
            }
        }
        for (; i + 1 < N; ++i) {
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
        i = N - 2;
        for (; i - (IV_PF - 1) >= 0; i -= IV_PF) {
            double yv[IV_PF], duv[IV_PF], du2v[IV_PF], ipv[IV_PF];
#pragma unroll
            for (int q = 0; q < IV_PF; ++q) {
                yv[q] = y[(size_t)(i - q) * G];
                duv[q] = du[(size_t)(i - q) * G];
                du2v[q] = du2[(size_t)(i - q) * G];
                ipv[q] = ip[(size_t)(i - q) * G];
            }
#pragma unroll
            for (int q = 0; q < IV_PF; ++q) {
                const double xi = (yv[q] - duv[q] * x1 - du2v[q] * x2) * ipv[q];
                y[(size_t)(i - q) * G] = xi;
                x2 = x1;
                x1 = xi;
                mx = fmax(mx, fabs(xi));
            }
        }
        for (; i >= 0; --i) {
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

// Compact-WY back-transformation for the panel reflectors of the band reduction (off = 8): x = Q_0 Q_1 ... Q_{P-1} y with
// Q_p = H_0 ... H_7 = I - V T V', T^-1 = striu(V'V) + diag(1 / tau).  One CTA per simulation; a warp owns 32 rows of all 16
// vectors as 4 x 2 FP64 tensor-core C fragments (16 doubles per thread).  A panel costs ONE reduction instead of eight and
// three barriers instead of eight, and every flop of it is a DMMA:
//   G  = V' [X | V]   8 x 24, k = rows: V straight from the reflector rows in A-fragment layout (the Gram part V'V uses the
//                     same registers as both operands), X through a shared-memory transpose of the warp's fragments;
//                     folded over the warps in warp order;
//   G' = T G          16 threads, back-substitution with T^-1 = striu(V'V) + diag(1 / tau);
//   X -= V G'         C fragments updated in place, V as A fragments, G' as 4 B fragments per thread.
// Warps whose rows all lie above the panel skip it.  backtransform_kernel spent ~1600 cycles per REFLECTOR (barrier + 5-level
// shuffle reduction of three interleaved CTAs).
#define WY_LDX 20      // row stride of the staged X rows (20 t + g: 16 distinct banks)
#define WY_SMEM(NW) ((size_t)((NW) * 32 * WY_LDX + (NW) * 192 + 192 + 128 + 2 * (NW) * 16) * sizeof(double))
template <int RPT, int NW>
__global__ void __launch_bounds__(NW * 32) backtransform_wy_kernel(const double* __restrict__ Aall,
                                                                   const double* __restrict__ tauall,
                                                                   const int* __restrict__ nfall, int N, int S,
                                                                   const double* __restrict__ W, double* __restrict__ V,
                                                                   int* __restrict__ status) {
    extern __shared__ __align__(16) double wsm[];
    double* stX = wsm;                                   // [NW][32][WY_LDX]
    double* part = stX + NW * 32 * WY_LDX;               // [NW][8][24]
    double* Gs = part + NW * 192;                        // [8][24]
    double* Gp = Gs + 192;                               // [8][16]
    double* red = Gp + 128;                              // [2][NW][16]
    const size_t G = (size_t)S * EV3_VMAX;
    const int sim = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int nf = nfall[sim];
    if (nf > EV3_VMAX || nf <= 0) return;
    const double* A = Aall + (size_t)sim * N * N;
    const double* tau = tauall + (size_t)sim * N;
    const double* y = W + 4 * (size_t)N * G + (size_t)sim * EV3_VMAX;      // element (i, vec) at y[i * G + vec]
    // x[r][rb][vb][e]: row 32 (w + NW r) + 8 rb + g, vector 8 vb + 2 t + e
    double x[RPT][4][2][2];
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int rb = 0; rb < 4; ++rb) {
            const int i = 32 * (w + NW * r) + 8 * rb + g;
#pragma unroll
            for (int vb = 0; vb < 2; ++vb) {
                const int v0 = 8 * vb + 2 * t;
                const double2 a = *reinterpret_cast<const double2*>(y + (size_t)min(i, N - 1) * G + v0);
                x[r][rb][vb][0] = (i < N && v0 < nf) ? a.x : 0.0;
                x[r][rb][vb][1] = (i < N && v0 + 1 < nf) ? a.y : 0.0;
            }
        }
    // ---- normalise every vector: scale by 1 / max |y|, then by 1 / ||y||; lane t of a warp's first quad ends up with the
    // ---- values of vectors 8 vb + 2 t + e
    {
        double m4[2][2];
#pragma unroll
        for (int vb = 0; vb < 2; ++vb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double m = 0.0;
#pragma unroll
                for (int r = 0; r < RPT; ++r)
#pragma unroll
                    for (int rb = 0; rb < 4; ++rb) m = fmax(m, fabs(x[r][rb][vb][e]));
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
                m4[vb][e] = m;
            }
        if (g == 0) {
#pragma unroll
            for (int vb = 0; vb < 2; ++vb)
#pragma unroll
                for (int e = 0; e < 2; ++e) red[w * 16 + 8 * vb + 2 * t + e] = m4[vb][e];
        }
        __syncthreads();
        double pre[2][2];
#pragma unroll
        for (int vb = 0; vb < 2; ++vb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int v = 8 * vb + 2 * t + e;
                double m = 0.0;
                for (int ww = 0; ww < NW; ++ww) m = fmax(m, red[ww * 16 + v]);
                pre[vb][e] = (v < nf) ? 1.0 / m : 0.0;
                double q = 0.0;
#pragma unroll
                for (int r = 0; r < RPT; ++r)
#pragma unroll
                    for (int rb = 0; rb < 4; ++rb) {
                        x[r][rb][vb][e] *= pre[vb][e];
                        q = fma(x[r][rb][vb][e], x[r][rb][vb][e], q);
                    }
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
                m4[vb][e] = q;
            }
        if (g == 0) {
#pragma unroll
            for (int vb = 0; vb < 2; ++vb)
#pragma unroll
                for (int e = 0; e < 2; ++e) red[(NW + w) * 16 + 8 * vb + 2 * t + e] = m4[vb][e];
        }
        __syncthreads();
        bool bad = false;
#pragma unroll
        for (int vb = 0; vb < 2; ++vb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int v = 8 * vb + 2 * t + e;
                double q = 0.0;
                for (int ww = 0; ww < NW; ++ww) q += red[(NW + ww) * 16 + v];
                const double sc = (v < nf) ? 1.0 / sqrt(q) : 0.0;
                bad |= (v < nf) && (!isfinite(sc) || !isfinite(pre[vb][e]));
#pragma unroll
                for (int r = 0; r < RPT; ++r)
#pragma unroll
                    for (int rb = 0; rb < 4; ++rb) x[r][rb][vb][e] *= sc;
            }
        if (bad && status) atomicOr(&status[sim], MCOSB_ST_EIGVEC);
    }
    // ---- panels, last to first ------------------------------------------------------------------------------------------
    const int npan = (N >= SBR_B + 2) ? (N - SBR_B - 2) / SBR_B + 1 : 0;
    double* sx = stX + w * 32 * WY_LDX;
    for (int p = npan - 1; p >= 0; --p) {
        const int j1 = SBR_B * p, R0 = j1 + SBR_B;
        double acc[3][2];
#pragma unroll
        for (int nb = 0; nb < 3; ++nb) acc[nb][0] = acc[nb][1] = 0.0;
        double va[RPT][4][2];                              // V as A fragments of the update: row 8 rb + g, reflector 4 h + t
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int row0 = 32 * (w + NW * r);
            if (row0 + 31 < R0 || row0 >= N) continue;     // warp-uniform: nothing of this panel in these rows
            // reflector values: unconditional loads from clamped rows, structural zeros and ones put in afterwards
            double vg[8];                                  // V as A = B fragments of G: row 4 q + t, reflector g
#pragma unroll
            for (int q = 0; q < 8; ++q) vg[q] = __ldg(A + (size_t)(j1 + g) * N + min(max(row0 + 4 * q + t, R0), N - 1));
#pragma unroll
            for (int rb = 0; rb < 4; ++rb)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    va[r][rb][h] = __ldg(A + (size_t)(j1 + 4 * h + t) * N + min(max(row0 + 8 * rb + g, R0), N - 1));
            // this warp's X fragments -> shared memory (row-major), read back with the rows on the k index
            __syncwarp();
#pragma unroll
            for (int rb = 0; rb < 4; ++rb)
#pragma unroll
                for (int vb = 0; vb < 2; ++vb)
                    *reinterpret_cast<double2*>(sx + (8 * rb + g) * WY_LDX + 8 * vb + 2 * t) =
                        make_double2(x[r][rb][vb][0], x[r][rb][vb][1]);
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int i = row0 + 4 * q + t, rho = i - R0;
                const double a = (i >= N || rho < g) ? 0.0 : (rho == g ? 1.0 : vg[q]);
                dmma_884(acc[0][0], acc[0][1], a, sx[(4 * q + t) * WY_LDX + g]);
                dmma_884(acc[1][0], acc[1][1], a, sx[(4 * q + t) * WY_LDX + 8 + g]);
                dmma_884(acc[2][0], acc[2][1], a, a);
            }
#pragma unroll
            for (int rb = 0; rb < 4; ++rb)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = row0 + 8 * rb + g, rho = i - R0, c = 4 * h + t;
                    va[r][rb][h] = (i >= N || rho < c) ? 0.0 : (rho == c ? -1.0 : -va[r][rb][h]);
                }
        }
#pragma unroll
        for (int nb = 0; nb < 3; ++nb)
            *reinterpret_cast<double2*>(part + w * 192 + g * 24 + 8 * nb + 2 * t) = make_double2(acc[nb][0], acc[nb][1]);
        __syncthreads();
        if (tid < 192) {
            double a = part[tid];
#pragma unroll
            for (int ww = 1; ww < NW; ++ww) a += part[ww * 192 + tid];
            Gs[tid] = a;
        }
        __syncthreads();
        if (tid < EV3_VMAX) {                                  // T^-1 G' = G for vector tid: back-substitution
            double gp[8];
#pragma unroll
            for (int c = 7; c >= 0; --c) {
                double a = Gs[c * 24 + tid];
#pragma unroll
                for (int xx = c + 1; xx < 8; ++xx) a = fma(-Gs[c * 24 + EV3_VMAX + xx], gp[xx], a);
                gp[c] = a * tau[j1 + c];
                Gp[c * 16 + tid] = gp[c];
            }
        }
        __syncthreads();
        double bg[2][2];                                       // G' as B fragments: reflector 4 h + t, vector 8 vb + g
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int vb = 0; vb < 2; ++vb) bg[h][vb] = Gp[(4 * h + t) * 16 + 8 * vb + g];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int row0 = 32 * (w + NW * r);
            if (row0 + 31 < R0 || row0 >= N) continue;
#pragma unroll
            for (int rb = 0; rb < 4; ++rb)
#pragma unroll
                for (int vb = 0; vb < 2; ++vb)
#pragma unroll
                    for (int h = 0; h < 2; ++h) dmma_884(x[r][rb][vb][0], x[r][rb][vb][1], va[r][rb][h], bg[h][vb]);
        }
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int rb = 0; rb < 4; ++rb) {
            const int i = 32 * (w + NW * r) + 8 * rb + g;
            if (i < N) {
#pragma unroll
                for (int vb = 0; vb < 2; ++vb)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int v = 8 * vb + 2 * t + e;
                        if (v < nf) V[((size_t)sim * EV3_VMAX + v) * N + i] = x[r][rb][vb][e];
                    }
            }
        }
}

// grid = (ceil(N / 128), S), 256 threads: the CTA stages the signal eigenvectors in shared memory once (they are re-read
// N / 32 times per simulation instead of N / 8 from L2), each warp owns 4 output rows and a lane walks the columns, so one
// shared-memory load of v_sj feeds 4 FMAs.  r_i = lbar + sum_s c_s v_si^2 is the diagonal of the clipped matrix.
#ifndef RC_ROWS
#define RC_ROWS 2
#endif
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
    static const bool wy_off = getenv("MCOSB_BACKTRANSFORM_WY") && atoi(getenv("MCOSB_BACKTRANSFORM_WY")) == 0;
    if (Q2 && !wy_off && N >= SBR_B + 2) {
        // panel reflectors of the band reduction: compact-WY, one CTA per simulation
#define WY_LAUNCH(RPT, NW)                                                                                            \
        do {                                                                                                          \
            MCOSB_CUDA(ctx, cudaFuncSetAttribute(backtransform_wy_kernel<RPT, NW>,                                    \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WY_SMEM(NW)));      \
            backtransform_wy_kernel<RPT, NW><<<S, NW * 32, WY_SMEM(NW), ctx->stream>>>(A, tau, nf, N, S, W, V, status); \
        } while (0)
        if (N <= 256) WY_LAUNCH(1, 8);
        else if (N <= 512) WY_LAUNCH(1, 16);
        else WY_LAUNCH(2, 16);
#undef WY_LAUNCH
    } else {
    const unsigned bt_grid = (unsigned)((G + 7) / 8);
    if (N <= 128) backtransform_kernel<4><<<bt_grid, 256, 2 * n * sizeof(double), ctx->stream>>>(A, tau, nf, N, S, W, V, status, off);
    else if (N <= 256) backtransform_kernel<8><<<bt_grid, 256, 2 * n * sizeof(double), ctx->stream>>>(A, tau, nf, N, S, W, V, status, off);
    else if (N <= 512) backtransform_kernel<16><<<bt_grid, 256, 2 * n * sizeof(double), ctx->stream>>>(A, tau, nf, N, S, W, V, status, off);
    else if (N <= 768) backtransform_kernel<24><<<bt_grid, 256, 2 * n * sizeof(double), ctx->stream>>>(A, tau, nf, N, S, W, V, status, off);
    else backtransform_kernel<32><<<bt_grid, 256, 2 * n * sizeof(double), ctx->stream>>>(A, tau, nf, N, S, W, V, status, off);
    }
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
