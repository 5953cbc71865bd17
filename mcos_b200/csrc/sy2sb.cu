// D2 (third generation), stage 1 of the two-stage tridiagonalisation: dense symmetric -> band (half bandwidth 8).
//
// tridiag2_kernel still pays one symmetric matrix-vector product per column (BLAS-2: 4 N^3 / 3 bytes of L2/HBM traffic
// per matrix, FP64 pipe at 7 % of peak).  Successive band reduction removes it: a panel of 8 columns is QR-factorised
// (Q = I - V T V'), and the whole trailing matrix is updated at once, A <- Q' A Q = A - V W' - W V' with
// W = Z T - 1/2 V (T' V' Z T), Z = A V.  Both the rank-16 update and the product Z = A V are FP64 tensor-core work
// (mma.sync m8n8k4 -> DMMA.8x8x4), and with one panel of look-ahead they share a single pass over the lower triangle:
//
//   for each panel p (columns j1 .. j1+7; V_p, W_p of the previous panel in shared memory):
//     a. bring the panel's own columns up to date with (V_p, W_p), one thread per row, rows in registers;
//        Householder-QR them below the band (one fused block reduction per column); R goes to the band, the reflectors
//        to row j1+c of the UPPER triangle (read coalesced by backtransform_kernel) and to shared memory (V_n), T_n.
//     b. ONE pass over the trailing lower triangle A[j1+8.., j1+8..] in 32 x 8 warp tiles:
//          tile -= V_p W_p' + W_p V_p'          (4 DMMA per 8 x 8 block; tile is the C fragment)
//          store tile
//          Z_n[rows] += tile  * V_n[cols]        (2 DMMA; the C fragment is the A operand, k index permuted)
//          Z_n[cols] += tile' * V_n[rows]        (2 DMMA; the block is transposed across the warp with 4 shuffles)
//        A warp owns the column blocks cb = warp (mod 8), so its Z_n[cols] accumulators never leave registers; the
//        Z_n[rows] partials of the 8 warps are summed per 32-row block row through shared memory in a fixed order.
//     c. W_n from Z_n (thread per row + one 8 x 8 DMMA reduction), rotate (V_p, W_p) <- (V_n, W_n).
//
// Traffic: one read + one write of the trailing triangle per panel = 8 N^3 / (3 * 8) bytes = 42 MB at N = 500 (tridiag2:
// 210 MB by design); flops 4 N^3 / 3 as before, now 32 FMA per element touched.  Everything is deterministic (no atomics).
// The band is finished by sb2st.cu (bulge chasing in shared memory).
#include "common.cuh"
#include "dmma.cuh"
#include "sbr.cuh"

#define S1_NT 256
#define S1_NW 8

// Sum 8 per-lane values over the 32 lanes (halving butterfly, 9 shuffles); lane 4 i ends up owning the total of value i.
__device__ __forceinline__ double s1_reduce8(double (&v)[8], int lane) {
    double a4[4], a2[2], a1;
    {
        const bool up = lane & 16;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double send = up ? v[q] : v[q + 4];
            const double keep = up ? v[q + 4] : v[q];
            a4[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const double send = up ? a4[q] : a4[q + 2];
            const double keep = up ? a4[q + 2] : a4[q];
            a2[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool up = lane & 4;
        const double send = up ? a2[0] : a2[1];
        const double keep = up ? a2[1] : a2[0];
        a1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
    a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
    return a1;   // value index = ((lane>>4)&1)*4 + ((lane>>3)&1)*2 + ((lane>>2)&1) = lane >> 2
}

__device__ __forceinline__ double s1_ldcg(const double* p) { return __ldcg(p); }
__device__ __forceinline__ void s1_stcg(double* p, double v) { __stcg(p, v); }

// RPT = rows per thread in the panel phases (ceil(N / 256)); QMAX = column blocks per warp (ceil(N / 64)).
template <int RPT, int QMAX>
__global__ void __launch_bounds__(S1_NT, 1)
sy2sb_kernel(double* Aall, int N, int LD, double* __restrict__ tauall) {
    extern __shared__ double sm[];
    double* Vp = sm;                          // [8][LD]   k-major: element (row i, k) at [k * LD + i]
    double* Wp = Vp + 8 * (size_t)LD;
    double* Vn = Wp + 8 * (size_t)LD;
    double* Zz = Vn + 8 * (size_t)LD;
    double* part = Zz + 8 * (size_t)LD;       // [2][8 warps][32 rows][8]
    double* red = part + 2 * S1_NW * 32 * 8;  // [2][8 warps][8]
    double* rowc = red + 2 * S1_NW * 8;       // [2][8]
    double* Tm = rowc + 16;                   // [8][8] T (upper triangular), row-major
    double* M0 = Tm + 64;                     // [8][8]
    double* Mm = M0 + 64;                     // [8][8]
    double* partM = Mm + 64;                  // [8 warps][64]

    const int s = blockIdx.x;
    double* A = Aall + (size_t)s * N * N;
    double* tau = tauall + (size_t)s * N;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const bool even = (N & 1) == 0;

    for (int i = tid; i < 32 * LD; i += S1_NT) sm[i] = 0.0;
    for (int i = tid; i < N; i += S1_NT) tau[i] = 0.0;
    __syncthreads();

    bool first = true;
    for (int j1 = 0;; j1 += SBR_B) {
        const int R0 = j1 + SBR_B;           // first row / column of the new trailing matrix
        const int ncols = min(SBR_B, N - j1);
        // ---- a. panel columns j1 .. j1+7, rows j1 .. N-1: update with (Vp, Wp) ------------------------------------
        const bool last = (N - R0 < 2);
        double p[RPT][8];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int i = j1 + tid + S1_NT * r;
#pragma unroll
            for (int c = 0; c < 8; ++c) p[r][c] = 0.0;
            if (i < N) {
                const double* row = A + (size_t)i * N + j1;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c < ncols && j1 + c <= i) p[r][c] = s1_ldcg(row + c);
                if (!first) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const double vi = Vp[k * LD + i], wi = Wp[k * LD + i];
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            p[r][c] = fma(-vi, Wp[k * LD + j1 + c], fma(-wi, Vp[k * LD + j1 + c], p[r][c]));
                    }
                }
                if (i < R0 || last) {          // diagonal block of the band (or nothing left to factorise): final
                    double* wrow = A + (size_t)i * N + j1;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (c < ncols && j1 + c <= i) s1_stcg(wrow + c, p[r][c]);
                }
            }
        }
        if (last) {
            // nothing left to eliminate; at most one more column (the last diagonal element) needs the update
            if (N - R0 == 1 && tid == 0 && !first) {
                const int i = N - 1;
                double a = s1_ldcg(A + (size_t)i * N + i);
                for (int k = 0; k < 8; ++k) a = fma(-2.0 * Vp[k * LD + i], Wp[k * LD + i], a);
                s1_stcg(A + (size_t)i * N + i, a);
            }
            break;
        }
        // ---- Householder QR of P = rows >= R0 of the panel; rho = i - R0 = tid - 8 + 256 r -----------------------------
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int par = c & 1;
            double pr[8];
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) pr[cc] = 0.0;
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const int rho = tid - 8 + S1_NT * r;
                if (rho > c) {                 // rows beyond N hold zeros
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) pr[cc] = fma(p[r][c], p[r][cc], pr[cc]);
                }
            }
            const double tot_lane = s1_reduce8(pr, lane);
            if (t == 0) red[(par * S1_NW + w) * 8 + g] = tot_lane;
            if (tid == c + 8) {
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) rowc[par * 8 + cc] = p[0][cc];
            }
            __syncthreads();
            double tot[8];
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
                double a = 0.0;
#pragma unroll
                for (int ww = 0; ww < S1_NW; ++ww) a += red[(par * S1_NW + ww) * 8 + cc];
                tot[cc] = a;
            }
            const double alpha = rowc[par * 8 + c];
            const double xn2 = tot[c];
            double tc = 0.0, scale = 0.0, beta = alpha;
            if (xn2 > 0.0) {
                beta = -copysign(sqrt(fma(alpha, alpha, xn2)), alpha);
                tc = (beta - alpha) / beta;
                scale = 1.0 / (alpha - beta);
            }
            // apply H_c to the columns on its right, then scale column c into v (beta on the diagonal)
#pragma unroll
            for (int cc = c + 1; cc < 8; ++cc) {
                const double wv = tc * fma(scale, tot[cc], rowc[par * 8 + cc]);
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    const int rho = tid - 8 + S1_NT * r;
                    if (rho > c) p[r][cc] = fma(-(p[r][c] * scale), wv, p[r][cc]);
                    else if (rho == c) p[r][cc] -= wv;
                }
            }
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const int rho = tid - 8 + S1_NT * r;
                if (rho > c) p[r][c] *= scale;
                else if (rho == c) p[r][c] = beta;
            }
            // column c of T: T[a][c] = -tau_c sum_{x<c} T[a][x] G[x][c], G[x][c] = v_x[c] + scale * (v_x . x_c below c)
            if (tid == 0) {
                double G[8];
#pragma unroll
                for (int x = 0; x < 8; ++x) G[x] = fma(scale, tot[x], rowc[par * 8 + x]);
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    double acc = 0.0;
                    if (a < c) {
#pragma unroll
                        for (int x = 0; x < 8; ++x)
                            if (x >= a && x < c) acc = fma(Tm[a * 8 + x], G[x], acc);
                        acc = -tc * acc;
                    } else if (a == c) {
                        acc = tc;
                    }
                    Tm[a * 8 + c] = acc;
                }
                tau[j1 + c] = tc;
            }
        }
        // ---- R to the band, reflectors to the upper triangle and to Vn -------------------------------------------------
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int rho = tid - 8 + S1_NT * r;
            const int i = R0 + rho;
            if (rho >= 0 && i < N) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (rho <= c) {
                        s1_stcg(A + (size_t)i * N + j1 + c, p[r][c]);
                        Vn[c * LD + i] = (rho == c) ? 1.0 : 0.0;
                    } else {
                        Vn[c * LD + i] = p[r][c];
                        A[(size_t)(j1 + c) * N + i] = p[r][c];
                    }
                }
            }
        }
        // block-row count and first tile of this warp
        const int nbr = (N - R0 + 31) >> 5;
        const int cb0 = R0 >> 3;
        int qlo = (cb0 - w + 7) >> 3;
        if (qlo < 0) qlo = 0;
        auto qhi_of = [&](int n) -> int {
            int last = R0 + 32 * n + 31;
            if (last > N - 1) last = N - 1;
            const int cbl = last >> 3;
            return (cbl - w) >= 0 ? ((cbl - w) >> 3) : -1;
        };
        double zc[QMAX][2];
#pragma unroll
        for (int q = 0; q < QMAX; ++q) zc[q][0] = zc[q][1] = 0.0;
        __syncthreads();   // Vn, Tm complete; panel stores visible

        // ---- b. fused pass --------------------------------------------------------------------------------------------
        double cur[4][2], nxt[4][2];
        auto load_tile = [&](int n, int q, double (&dst)[4][2]) {
            const int i0 = R0 + 32 * n, j0 = 8 * (8 * q + w);
            const bool full = (j0 + 7 <= i0) && (i0 + 31 < N) && even;
            if (full) {
#pragma unroll
                for (int rb = 0; rb < 4; ++rb) {
                    const double2 v2 = __ldcg(reinterpret_cast<const double2*>(A + (size_t)(i0 + 8 * rb + g) * N + j0 + 2 * t));
                    dst[rb][0] = v2.x;
                    dst[rb][1] = v2.y;
                }
            } else {
#pragma unroll
                for (int rb = 0; rb < 4; ++rb) {
                    const int i = i0 + 8 * rb + g, j = j0 + 2 * t;
                    dst[rb][0] = (i < N && j <= i) ? s1_ldcg(A + (size_t)i * N + j) : 0.0;
                    dst[rb][1] = (i < N && j + 1 <= i) ? s1_ldcg(A + (size_t)i * N + j + 1) : 0.0;
                }
            }
        };
        {
            // prologue: first tile of this warp (if any)
            int n = 0;
            while (n < nbr && qlo > qhi_of(n)) ++n;
            if (n < nbr) load_tile(n, qlo, cur);
        }
        for (int n = 0; n < nbr; ++n) {
            const int i0 = R0 + 32 * n;
            const int qhi = qhi_of(n);
            const int par = n & 1;
            double zi[4][2];
#pragma unroll
            for (int rb = 0; rb < 4; ++rb) zi[rb][0] = zi[rb][1] = 0.0;
            if (qlo <= qhi) {
                // row operands of this block row
                double av[4][4], vni[4][2];
#pragma unroll
                for (int rb = 0; rb < 4; ++rb) {
                    const int i = i0 + 8 * rb + g;
                    av[rb][0] = -Vp[t * LD + i];
                    av[rb][1] = -Vp[(4 + t) * LD + i];
                    av[rb][2] = -Wp[t * LD + i];
                    av[rb][3] = -Wp[(4 + t) * LD + i];
                    vni[rb][0] = Vn[g * LD + i0 + 8 * rb + t];
                    vni[rb][1] = Vn[g * LD + i0 + 8 * rb + 4 + t];
                }
#pragma unroll
                for (int q = 0; q < QMAX; ++q) {
                    if (q >= qlo && q <= qhi) {
                        const int j0 = 8 * (8 * q + w);
                        // prefetch the next tile of this warp
                        {
                            int nn = n, qq = q + 1;
                            if (qq > qhi) {
                                nn = n + 1;
                                qq = qlo;
                                while (nn < nbr && qlo > qhi_of(nn)) ++nn;
                            }
                            if (nn < nbr) load_tile(nn, qq, nxt);
                        }
                        const bool full = (j0 + 7 <= i0) && (i0 + 31 < N);
                        // column operands
                        const double bw0 = Wp[t * LD + j0 + g], bw1 = Wp[(4 + t) * LD + j0 + g];
                        const double bv0 = Vp[t * LD + j0 + g], bv1 = Vp[(4 + t) * LD + j0 + g];
                        const double2 vnj = *reinterpret_cast<const double2*>(Vn + g * LD + j0 + 2 * t);
                        if (!first) {
#pragma unroll
                            for (int rb = 0; rb < 4; ++rb) dmma_884(cur[rb][0], cur[rb][1], av[rb][0], bw0);
#pragma unroll
                            for (int rb = 0; rb < 4; ++rb) dmma_884(cur[rb][0], cur[rb][1], av[rb][1], bw1);
#pragma unroll
                            for (int rb = 0; rb < 4; ++rb) dmma_884(cur[rb][0], cur[rb][1], av[rb][2], bv0);
#pragma unroll
                            for (int rb = 0; rb < 4; ++rb) dmma_884(cur[rb][0], cur[rb][1], av[rb][3], bv1);
                            // store
                            if (full && even) {
#pragma unroll
                                for (int rb = 0; rb < 4; ++rb)
                                    __stcg(reinterpret_cast<double2*>(A + (size_t)(i0 + 8 * rb + g) * N + j0 + 2 * t),
                                           make_double2(cur[rb][0], cur[rb][1]));
                            } else {
#pragma unroll
                                for (int rb = 0; rb < 4; ++rb) {
                                    const int i = i0 + 8 * rb + g, j = j0 + 2 * t;
                                    if (i < N && j <= i) s1_stcg(A + (size_t)i * N + j, cur[rb][0]);
                                    if (i < N && j + 1 <= i) s1_stcg(A + (size_t)i * N + j + 1, cur[rb][1]);
                                }
                            }
                        }
                        // strictly-lower copy for the transposed product; mask what lies above the diagonal
                        double lt[4][2];
                        if (full) {
#pragma unroll
                            for (int rb = 0; rb < 4; ++rb) { lt[rb][0] = cur[rb][0]; lt[rb][1] = cur[rb][1]; }
                        } else {
#pragma unroll
                            for (int rb = 0; rb < 4; ++rb) {
                                const int i = i0 + 8 * rb + g, j = j0 + 2 * t;
                                const bool in = i < N;
                                cur[rb][0] = (in && j <= i) ? cur[rb][0] : 0.0;
                                cur[rb][1] = (in && j + 1 <= i) ? cur[rb][1] : 0.0;
                                lt[rb][0] = (in && j < i) ? cur[rb][0] : 0.0;
                                lt[rb][1] = (in && j + 1 < i) ? cur[rb][1] : 0.0;
                            }
                        }
                        // Z[rows] += tile * Vn[cols]  (k index permuted: step 0 = even columns, step 1 = odd columns)
#pragma unroll
                        for (int rb = 0; rb < 4; ++rb) dmma_884(zi[rb][0], zi[rb][1], cur[rb][0], vnj.x);
#pragma unroll
                        for (int rb = 0; rb < 4; ++rb) dmma_884(zi[rb][0], zi[rb][1], cur[rb][1], vnj.y);
                        // Z[cols] += tile' * Vn[rows]: B operand element (k = t -> row 4h + t, n = g -> column g)
#pragma unroll
                        for (int rb = 0; rb < 4; ++rb) {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int src = 4 * (4 * h + t) + (g >> 1);
                                const double e0 = __shfl_sync(0xffffffffu, lt[rb][0], src);
                                const double e1 = __shfl_sync(0xffffffffu, lt[rb][1], src);
                                dmma_884(zc[q][0], zc[q][1], vni[rb][h], (g & 1) ? e1 : e0);
                            }
                        }
#pragma unroll
                        for (int rb = 0; rb < 4; ++rb) { cur[rb][0] = nxt[rb][0]; cur[rb][1] = nxt[rb][1]; }
                    }
                }
            }
            // block-row reduction of the Z[rows] partials over the 8 warps
            double* pp = part + (size_t)par * S1_NW * 256;
#pragma unroll
            for (int rb = 0; rb < 4; ++rb)
                *reinterpret_cast<double2*>(pp + (w * 32 + rb * 8 + g) * 8 + 2 * t) = make_double2(zi[rb][0], zi[rb][1]);
            __syncthreads();
            {
                const int r = tid >> 3, kk = tid & 7;
                double a = 0.0;
#pragma unroll
                for (int ww = 0; ww < S1_NW; ++ww) a += pp[(ww * 32 + r) * 8 + kk];
                Zz[kk * LD + i0 + r] = a;
            }
        }
        __syncthreads();
        // add the column-side accumulators (each (k, column) has exactly one owner)
#pragma unroll
        for (int q = 0; q < QMAX; ++q) {
            const int j0 = 8 * (8 * q + w);
            if (j0 + 7 >= R0 && j0 < N) {
                Zz[g * LD + j0 + 2 * t] += zc[q][0];
                Zz[g * LD + j0 + 2 * t + 1] += zc[q][1];
            }
        }
        __syncthreads();

        // ---- c. W = Z T - 1/2 V (T' V' Z T) -------------------------------------------------------------------------------
        double y[RPT][8];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int i = R0 + tid + S1_NT * r;
#pragma unroll
            for (int c = 0; c < 8; ++c) y[r][c] = 0.0;
            if (i < N) {
                double z[8];
#pragma unroll
                for (int x = 0; x < 8; ++x) z[x] = Zz[x * LD + i];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    double a = 0.0;
#pragma unroll
                    for (int x = 0; x <= c; ++x) a = fma(z[x], Tm[x * 8 + c], a);
                    y[r][c] = a;
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) Zz[c * LD + i] = y[r][c];
            }
        }
        __syncthreads();
        {
            double m0 = 0.0, m1 = 0.0;
            for (int i4 = R0 + 4 * w; i4 < N; i4 += 4 * S1_NW) dmma_884(m0, m1, Vn[g * LD + i4 + t], Zz[g * LD + i4 + t]);
            partM[w * 64 + g * 8 + 2 * t] = m0;
            partM[w * 64 + g * 8 + 2 * t + 1] = m1;
        }
        __syncthreads();
        if (tid < 64) {
            double a = 0.0;
#pragma unroll
            for (int ww = 0; ww < S1_NW; ++ww) a += partM[ww * 64 + tid];
            M0[tid] = a;
        }
        __syncthreads();
        if (tid < 64) {
            const int c2 = tid >> 3, c = tid & 7;
            double a = 0.0;
            for (int x = 0; x <= c2; ++x) a = fma(Tm[x * 8 + c2], M0[x * 8 + c], a);
            Mm[tid] = a;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int i = R0 + tid + S1_NT * r;
            if (i < N) {
                double v[8];
#pragma unroll
                for (int x = 0; x < 8; ++x) v[x] = Vn[x * LD + i];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    double a = 0.0;
#pragma unroll
                    for (int x = 0; x < 8; ++x) a = fma(v[x], Mm[x * 8 + c], a);
                    Zz[c * LD + i] = fma(-0.5, a, y[r][c]);
                }
            }
        }
        __syncthreads();
        // rotate: (Vp, Wp) <- (Vn, Wn)
        double* tp = Vp; Vp = Vn; Vn = tp;
        tp = Wp; Wp = Zz; Zz = tp;
        first = false;
    }
}

// LD = 8 (mod 16) keeps the 8-lane x 4-k fragment loads of the k-major arrays on distinct banks; >= N + 32 so block
// rows may overhang N (those rows hold zeros).
int sbr_stage1_ld(int N) {
    int ld = N + 32;
    while ((ld & 15) != 8) ++ld;
    return ld;
}
size_t sbr_stage1_smem(int N) {
    return ((size_t)32 * sbr_stage1_ld(N) + 2 * S1_NW * 256 + 2 * S1_NW * 8 + 16 + 3 * 64 + S1_NW * 64) * sizeof(double);
}

// Returns 1 if launched, 0 if N is outside the range of this kernel.
int mcosb_launch_sy2sb(mcosb_ctx* ctx, int S, double* A, double* tau) {
    const int N = ctx->N;
    if (N < SBR_B + 2) return 0;
    const size_t smem = sbr_stage1_smem(N);
    if (smem > (size_t)ctx->smem_optin) return 0;
    const int LD = sbr_stage1_ld(N);
#define S1_LAUNCH(RPT, QMAX)                                                                                          \
    do {                                                                                                              \
        MCOSB_CUDA(ctx, cudaFuncSetAttribute(sy2sb_kernel<RPT, QMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                             (int)smem));                                                             \
        sy2sb_kernel<RPT, QMAX><<<S, S1_NT, smem, ctx->stream>>>(A, N, LD, tau);                                      \
    } while (0)
    if (N <= 256) S1_LAUNCH(1, 4);
    else if (N <= 512) S1_LAUNCH(2, 8);
    else if (N <= 768) S1_LAUNCH(3, 12);
    else return 0;
#undef S1_LAUNCH
    MCOSB_LAUNCHED(ctx, MK_SY2SB);
    return 1;
}
