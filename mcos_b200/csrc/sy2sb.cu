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


#define S1_D 2    // tiles in flight per warp (register ring), a power of two
#ifndef S1_L2_PREFETCH
#define S1_L2_PREFETCH 0   // block rows ahead of an L2 prefetch of the tiles; measured 0: 18.36, 1: 18.37, 4: 19.09 us -- off
#endif

#ifdef S1_TIMING
#define S1_TICK(idx) do { if (threadIdx.x == 0) { long long now__ = clock64(); s1_acc[idx] += now__ - s1_last; s1_last = now__; } } while (0)
#else
#define S1_TICK(idx) do { } while (0)
#endif

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

// 1/x and 1/sqrt(x) to ~1 ulp: hardware approximation (2^-23) + two Newton steps; FP64 division / sqrt are ~40-instruction
// sequences that sit on the critical path of every Householder column
__device__ __forceinline__ double s1_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}
__device__ __forceinline__ double s1_rsqrt(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double h = 0.5 * x;
    r = fma(r, fma(-h * r, r, 0.5), r);
    r = fma(r, fma(-h * r, r, 0.5), r);
    return r;
}

__device__ __forceinline__ double s1_ldcg(const double* p) { return __ldcg(p); }
__device__ __forceinline__ void s1_stcg(double* p, double v) { __stcg(p, v); }

// RPT = rows per thread in the panel phases (ceil(N / threads)); QMAX = column blocks per class (ceil(N / 64)); NW = 16
// warps: warp w works on column blocks cb = w % 8 (mod 8) and on the 16-row half w / 8 of a 32-row block row.
template <int RPT, int QMAX, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
sy2sb_kernel(double* Aall, int N, int LD, double* __restrict__ tauall) {
    constexpr int NT = NW * 32;
    extern __shared__ double sm[];
    // four k-major panel arrays [8][LD] (element (row i, k) at [k * LD + i]) in two pairs that swap roles every panel:
    // (V_p, W_p) = pair `flip`, (V_n, Z / W_n) = the other.  The addresses are derived from `flip` where they are used
    // (four rotating pointer variables cost two spilled registers that every warp reloaded from L2 -- the L1 is 28 KB next
    // to 182 KB of shared memory -- right before the barrier of every 32-row block row).
    double* part = sm + 32 * (size_t)LD;      // [2][8 column classes][32 rows][8]
    int flip = 0;
#define Vp (sm + (flip ? 16 * LD : 0))
#define Wp (sm + (flip ? 24 * LD : 8 * LD))
#define Vn (sm + (flip ? 0 : 16 * LD))
#define Zz (sm + (flip ? 8 * LD : 24 * LD))
    double* red = part + 2 * 8 * 32 * 8;      // [2][NW][8]
    double* rowc = red + 2 * NW * 8;          // [2][8]
    double* Gm = rowc + 16;                   // [8][8] strictly upper part of V'V (T^-1 = striu(V'V) + diag(1/tau))
    double* tauS = Gm + 64;                   // [8]
    double* M0 = tauS + 8;                    // [8][8]
    double* Mm = M0 + 64;                     // [8][8]
    double* partM = Mm + 64;                  // [NW][64]

    const int s = blockIdx.x;
    double* A = Aall + (size_t)s * N * N;
    double* tau = tauall + (size_t)s * N;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wc = w & 7, wh = w >> 3;
    const bool even = (N & 1) == 0;

    for (int i = tid; i < 32 * LD; i += NT) sm[i] = 0.0;
    for (int i = tid; i < N; i += NT) tau[i] = 0.0;
    __syncthreads();

#ifdef S1_TIMING
    long long s1_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long s1_last = clock64();
#endif
    bool first = true;
    for (int j1 = 0;; j1 += SBR_B) {
        const int R0 = j1 + SBR_B;           // first row / column of the new trailing matrix
        const int ncols = min(SBR_B, N - j1);
        // ---- a. panel columns j1 .. j1+7, rows j1 .. N-1: update with (Vp, Wp) ------------------------------------
        const bool last = (N - R0 < 2);
        double p[RPT][8];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int i = j1 + tid + NT * r;
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
        S1_TICK(0);
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
        // ---- Householder QR of P = rows >= R0 of the panel; rho = i - R0 = tid - 8 + NT r ------------------------------
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int par = c & 1;
            double pr[8];
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) pr[cc] = 0.0;
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const int rho = tid - 8 + NT * r;
                if (rho > c) {                 // rows beyond N hold zeros
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) pr[cc] = fma(p[r][c], p[r][cc], pr[cc]);
                }
            }
            const double tot_lane = s1_reduce8(pr, lane);
            if (t == 0) red[(par * NW + w) * 8 + g] = tot_lane;
            if (tid == c + 8) {
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) rowc[par * 8 + cc] = p[0][cc];
            }
            __syncthreads();
            // second level: lane (j, cc) sums warps j, j+4, ...; two shuffles fold j; then every lane collects all 8
            double tot[8];
            {
                const int cc = lane & 7, jj = lane >> 3;
                double a = 0.0;
#pragma unroll
                for (int ww = 0; ww < NW; ww += 4) a += red[(par * NW + ww + jj) * 8 + cc];
                a += __shfl_xor_sync(0xffffffffu, a, 8);
                a += __shfl_xor_sync(0xffffffffu, a, 16);
#pragma unroll
                for (int c2 = 0; c2 < 8; ++c2) tot[c2] = __shfl_sync(0xffffffffu, a, c2);
            }
            const double alpha = rowc[par * 8 + c];
            const double xn2 = tot[c];
            double tc = 0.0, scale = 0.0, beta = alpha;
            if (xn2 > 1e-280) {
                const double n2 = fma(alpha, alpha, xn2);
                const double rs = s1_rsqrt(n2), nrm = n2 * rs, aa = fabs(alpha);
                beta = -copysign(nrm, alpha);
                tc = fma(aa, rs, 1.0);                               // (beta - alpha) / beta
                scale = copysign(s1_rcp(aa + nrm), alpha);          // 1 / (alpha - beta)
            }
            // apply H_c to the columns on its right, then scale column c into v (beta on the diagonal)
#pragma unroll
            for (int cc = c + 1; cc < 8; ++cc) {
                const double wv = tc * fma(scale, tot[cc], rowc[par * 8 + cc]);
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    const int rho = tid - 8 + NT * r;
                    if (rho > c) p[r][cc] = fma(-(p[r][c] * scale), wv, p[r][cc]);
                    else if (rho == c) p[r][cc] -= wv;
                }
            }
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const int rho = tid - 8 + NT * r;
                if (rho > c) p[r][c] *= scale;
                else if (rho == c) p[r][c] = beta;
            }
            // G[x][c] = v_x . v_c = v_x[c] + scale * (v_x . x_c below row c), x < c
            if (tid == 64) {
#pragma unroll
                for (int x = 0; x < 8; ++x)
                    if (x < c) Gm[x * 8 + c] = fma(scale, tot[x], rowc[par * 8 + x]);
                tauS[c] = tc;
                tau[j1 + c] = tc;
            }
        }
        S1_TICK(1);
        // ---- R to the band, reflectors to the upper triangle and to Vn -------------------------------------------------
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int rho = tid - 8 + NT * r;
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
        // tile (n, q) of this warp: rows R0 + 32 n + 16 wh + [0, 16), column block 8 q + wc.  Column blocks are owned by
        // class (wc fixed), so the column-side sums zc[q] stay in registers for the whole pass.
        const int nbr = (N - R0 + 31) >> 5;
        int qlo = ((R0 >> 3) - wc + 7) >> 3;
        if (qlo < 0) qlo = 0;
        auto qhi_of = [&](int n) -> int {        // last column block that reaches the lower triangle in block row n
            int lastrow = R0 + 32 * n + 16 * wh + 15;
            if (lastrow > N - 1) lastrow = N - 1;
            const int cbl = lastrow >> 3;
            return (cbl - wc) >= 0 ? ((cbl - wc) >> 3) : -1;
        };
        auto qfast_of = [&](int n) -> int {      // last column block entirely below the diagonal, all 16 rows inside
            const int i0 = R0 + 32 * n + 16 * wh;
            if (!even || n >= nbr || i0 + 15 >= N) return -1;
            const int num = i0 - 7 - 8 * wc;
            return num >= 0 ? (num >> 6) : -1;
        };
        double zc[QMAX][2];
#pragma unroll
        for (int q = 0; q < QMAX; ++q) zc[q][0] = zc[q][1] = 0.0;
        __syncthreads();   // Vn, Gm complete; panel stores visible
        S1_TICK(2);

        // ---- b. ONE pass over the trailing lower triangle -------------------------------------------------------------
        const int cK0 = t * LD + 8 * wc + g, cK4 = (4 + t) * LD + 8 * wc + g;   // fragment offsets for column block q = 0
        const int cVn = g * LD + 8 * wc + 2 * t;
        const int rowstep = 8 * N;
        // ring of S1_D tiles in registers; slot q % S1_D holds the next unprocessed fast tile with that residue.  Every
        // load is unconditional (clamped address) and lands directly in the ring: the ring registers are only read (C
        // input of an out-of-place DMMA) before they are refilled, so ptxas has no reason to stage loads in temporaries
        // and MOVE them -- a move would wait for the load and lose the prefetch.
        double ring[S1_D][2][2];
        auto ring_load = [&](double (&dst)[2][2], const double* ptr) {
#pragma unroll
            for (int rb = 0; rb < 2; ++rb) {
                const double2 v2 = __ldcg(reinterpret_cast<const double2*>(ptr + rb * rowstep));
                dst[rb][0] = v2.x;
                dst[rb][1] = v2.y;
            }
        };
        const double* tile00 = A + (size_t)(R0 + 16 * wh + g) * N + 8 * wc + 2 * t;   // tile (0, 0); (n, q): + 32 n N + 64 q
        {
            const int qf0 = min(qfast_of(0), qhi_of(0));
#pragma unroll
            for (int d = 0; d < S1_D; ++d) {
                const int qd = qlo + ((d - qlo) & (S1_D - 1));
                if (qd <= qf0) ring_load(ring[d], tile00 + 64 * qd);
            }
        }
        for (int n = 0; n < nbr; ++n) {
            const int i0 = R0 + 32 * n + 16 * wh;
            const int qhi = qhi_of(n);
            const int qf = min(qfast_of(n), qhi);
            const int qfn = (n + 1 < nbr) ? min(qfast_of(n + 1), qhi_of(n + 1)) : -1;
            const int par = n & 1;
            double* rowbase = A + (size_t)(i0 + g) * N + 8 * wc + 2 * t;
            const double* nextbase = rowbase + (size_t)32 * N;
#if S1_L2_PREFETCH
            // pull this warp's tiles of block row n + S1_L2_PREFETCH towards L2 (no registers, no scoreboard): 28 % of the tile
            // loads miss L2 (one matrix per SM = 150 MB of triangles against 126 MB), and the register ring only covers an L2
            // hit.  One instruction per tile: lanes 0-15 touch the first byte of the tile's 16 row segments (64 B each),
            // lanes 16-31 the last, so segments that straddle a line are covered.
            {
                const int np = n + S1_L2_PREFETCH;
                if (np < nbr) {
                    const int prow = R0 + 32 * np + 16 * wh + (lane & 15);
                    if (prow < N) {
                        const double* pbase = A + (size_t)prow * N + 8 * wc + ((lane & 16) ? 7 : 0);
                        const int qh = qhi_of(np);
                        for (int q = qlo; q <= qh; ++q)
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(pbase + 64 * q));
                    }
                }
            }
#endif
            // slots this block row does not use take their tile of the next block row now
#pragma unroll
            for (int d = 0; d < S1_D; ++d) {
                const int qd = qlo + ((d - qlo) & (S1_D - 1));
                if (qd > qf && qd <= qfn) ring_load(ring[d], nextbase + 64 * qd);
            }
            double zi[2][2];
#pragma unroll
            for (int rb = 0; rb < 2; ++rb) zi[rb][0] = zi[rb][1] = 0.0;
            if (qlo <= qhi) {
                double av[2][4], vni[2][2];
#pragma unroll
                for (int rb = 0; rb < 2; ++rb) {
                    const int i = i0 + 8 * rb + g;
                    av[rb][0] = -Vp[t * LD + i];
                    av[rb][1] = -Vp[(4 + t) * LD + i];
                    av[rb][2] = -Wp[t * LD + i];
                    av[rb][3] = -Wp[(4 + t) * LD + i];
                    vni[rb][0] = Vn[g * LD + i0 + 8 * rb + t];
                    vni[rb][1] = Vn[g * LD + i0 + 8 * rb + 4 + t];
                }
                // ---- tiles entirely below the diagonal ---------------------------------------------------------------------
#pragma unroll
                for (int q = 0; q < QMAX; ++q) {
                    if (q >= qlo && q <= qf) {
                        double (&cur)[2][2] = ring[q % S1_D];
                        // refill address: same block row if it has a tile q + D, else this slot's tile of the next block row,
                        // else (nothing left for the slot) the current tile again -- always valid memory
                        const int q2 = qlo + ((q - qlo) & (S1_D - 1));
                        const double* refill = (q + S1_D <= qf) ? rowbase + 64 * (q + S1_D)
                                                                 : ((q2 <= qfn) ? nextbase + 64 * q2 : rowbase + 64 * q);
                        const double bw0 = Wp[cK0 + 64 * q], bw1 = Wp[cK4 + 64 * q];
                        const double bv0 = Vp[cK0 + 64 * q], bv1 = Vp[cK4 + 64 * q];
                        const double2 vnj = *reinterpret_cast<const double2*>(Vn + cVn + 64 * q);
                        double u[2][2];
                        if (!first) {
#pragma unroll
                            for (int rb = 0; rb < 2; ++rb) dmma_884_oop(u[rb][0], u[rb][1], av[rb][0], bw0, cur[rb][0], cur[rb][1]);
                            ring_load(cur, refill);
#pragma unroll
                            for (int rb = 0; rb < 2; ++rb) dmma_884(u[rb][0], u[rb][1], av[rb][1], bw1);
#pragma unroll
                            for (int rb = 0; rb < 2; ++rb) dmma_884(u[rb][0], u[rb][1], av[rb][2], bv0);
#pragma unroll
                            for (int rb = 0; rb < 2; ++rb) dmma_884(u[rb][0], u[rb][1], av[rb][3], bv1);
#pragma unroll
                            for (int rb = 0; rb < 2; ++rb)
                                __stcg(reinterpret_cast<double2*>(rowbase + 64 * q + rb * rowstep),
                                       make_double2(u[rb][0], u[rb][1]));
                        } else {
#pragma unroll
                            for (int rb = 0; rb < 2; ++rb) { u[rb][0] = cur[rb][0] + 0.0; u[rb][1] = cur[rb][1] + 0.0; }
                            ring_load(cur, refill);
                        }
#pragma unroll
                        for (int rb = 0; rb < 2; ++rb) dmma_884(zi[rb][0], zi[rb][1], u[rb][0], vnj.x);
#pragma unroll
                        for (int rb = 0; rb < 2; ++rb) dmma_884(zi[rb][0], zi[rb][1], u[rb][1], vnj.y);
                        // Z[cols] += tile' * Vn[rows]: B operand element (k = t -> row 4h + t, n = g -> column g)
#pragma unroll
                        for (int rb = 0; rb < 2; ++rb) {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int src = 4 * (4 * h + t) + (g >> 1);
                                const double e0 = __shfl_sync(0xffffffffu, u[rb][0], src);
                                const double e1 = __shfl_sync(0xffffffffu, u[rb][1], src);
                                dmma_884(zc[q][0], zc[q][1], vni[rb][h], (g & 1) ? e1 : e0);
                            }
                        }
                    }
                }
                // ---- tiles that touch the diagonal or overhang N: element-wise predicates -----------------------------------
#pragma unroll
                for (int q = 0; q < QMAX; ++q) {
                    if (q >= qlo && q > qf && q <= qhi) {
                        const int j0 = 8 * (8 * q + wc);
                        double cur[2][2], lt[2][2];
#pragma unroll
                        for (int rb = 0; rb < 2; ++rb) {
                            const int i = i0 + 8 * rb + g, j = j0 + 2 * t;
                            cur[rb][0] = (i < N && j <= i) ? s1_ldcg(A + (size_t)i * N + j) : 0.0;
                            cur[rb][1] = (i < N && j + 1 <= i) ? s1_ldcg(A + (size_t)i * N + j + 1) : 0.0;
                        }
                        const double bw0 = Wp[cK0 + 64 * q], bw1 = Wp[cK4 + 64 * q];
                        const double bv0 = Vp[cK0 + 64 * q], bv1 = Vp[cK4 + 64 * q];
                        const double2 vnj = *reinterpret_cast<const double2*>(Vn + cVn + 64 * q);
                        if (!first) {
#pragma unroll
                            for (int rb = 0; rb < 2; ++rb) dmma_884(cur[rb][0], cur[rb][1], av[rb][0], bw0);
#pragma unroll
                            for (int rb = 0; rb < 2; ++rb) dmma_884(cur[rb][0], cur[rb][1], av[rb][1], bw1);
#pragma unroll
                            for (int rb = 0; rb < 2; ++rb) dmma_884(cur[rb][0], cur[rb][1], av[rb][2], bv0);
#pragma unroll
                            for (int rb = 0; rb < 2; ++rb) dmma_884(cur[rb][0], cur[rb][1], av[rb][3], bv1);
#pragma unroll
                            for (int rb = 0; rb < 2; ++rb) {
                                const int i = i0 + 8 * rb + g, j = j0 + 2 * t;
                                if (i < N && j <= i) s1_stcg(A + (size_t)i * N + j, cur[rb][0]);
                                if (i < N && j + 1 <= i) s1_stcg(A + (size_t)i * N + j + 1, cur[rb][1]);
                            }
                        }
#pragma unroll
                        for (int rb = 0; rb < 2; ++rb) {
                            const int i = i0 + 8 * rb + g, j = j0 + 2 * t;
                            const bool in = i < N;
                            cur[rb][0] = (in && j <= i) ? cur[rb][0] : 0.0;
                            cur[rb][1] = (in && j + 1 <= i) ? cur[rb][1] : 0.0;
                            lt[rb][0] = (in && j < i) ? cur[rb][0] : 0.0;        // strictly lower: the diagonal counts once
                            lt[rb][1] = (in && j + 1 < i) ? cur[rb][1] : 0.0;
                        }
#pragma unroll
                        for (int rb = 0; rb < 2; ++rb) dmma_884(zi[rb][0], zi[rb][1], cur[rb][0], vnj.x);
#pragma unroll
                        for (int rb = 0; rb < 2; ++rb) dmma_884(zi[rb][0], zi[rb][1], cur[rb][1], vnj.y);
#pragma unroll
                        for (int rb = 0; rb < 2; ++rb) {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int src = 4 * (4 * h + t) + (g >> 1);
                                const double e0 = __shfl_sync(0xffffffffu, lt[rb][0], src);
                                const double e1 = __shfl_sync(0xffffffffu, lt[rb][1], src);
                                dmma_884(zc[q][0], zc[q][1], vni[rb][h], (g & 1) ? e1 : e0);
                            }
                        }
                    }
                }
            }
            // block-row reduction of the Z[rows] partials over the 8 column classes
            double* pp = part + (size_t)par * 8 * 256;
#pragma unroll
            for (int rb = 0; rb < 2; ++rb)
                *reinterpret_cast<double2*>(pp + (wc * 32 + 16 * wh + rb * 8 + g) * 8 + 2 * t) =
                    make_double2(zi[rb][0], zi[rb][1]);
            __syncthreads();
            if (tid < 256) {
                const int r = tid >> 3, kk = tid & 7;
                double a = 0.0;
#pragma unroll
                for (int ww = 0; ww < 8; ++ww) a += pp[(ww * 32 + r) * 8 + kk];
                Zz[kk * LD + R0 + 32 * n + r] = a;
            }
        }
        __syncthreads();
        S1_TICK(3);
        // add the column-side accumulators (each (k, column) has one owner per row half; halves in a fixed order)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            if (wh == hh) {
#pragma unroll
                for (int q = 0; q < QMAX; ++q) {
                    const int j0 = 8 * (8 * q + wc);
                    if (j0 + 7 >= R0 && j0 < N) {
                        Zz[g * LD + j0 + 2 * t] += zc[q][0];
                        Zz[g * LD + j0 + 2 * t + 1] += zc[q][1];
                    }
                }
            }
            __syncthreads();
        }

        // ---- c. W = Z T - 1/2 V (T' V' Z T), T^-1 = striu(G) + diag(1/tau): triangular solves instead of forming T -------
        double y[RPT][8];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int i = R0 + tid + NT * r;
#pragma unroll
            for (int c = 0; c < 8; ++c) y[r][c] = 0.0;
            if (i < N) {
                double z[8];
#pragma unroll
                for (int x = 0; x < 8; ++x) z[x] = Zz[x * LD + i];
#pragma unroll
                for (int c = 0; c < 8; ++c) {       // Y U = Z
                    double a = z[c];
#pragma unroll
                    for (int x = 0; x < c; ++x) a = fma(-y[r][x], Gm[x * 8 + c], a);
                    y[r][c] = a * tauS[c];
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) Zz[c * LD + i] = y[r][c];
            }
        }
        __syncthreads();
        {
            double m0 = 0.0, m1 = 0.0;
            for (int i4 = R0 + 4 * w; i4 < N; i4 += 4 * NW) dmma_884(m0, m1, Vn[g * LD + i4 + t], Zz[g * LD + i4 + t]);
            partM[w * 64 + g * 8 + 2 * t] = m0;
            partM[w * 64 + g * 8 + 2 * t + 1] = m1;
        }
        __syncthreads();
        if (tid < 64) {
            double a = 0.0;
#pragma unroll
            for (int ww = 0; ww < NW; ++ww) a += partM[ww * 64 + tid];
            M0[tid] = a;
        }
        __syncthreads();
        if (tid < 8) {                               // U' M = M0, column tid of M
            const int c = tid;
            double mcol[8];
#pragma unroll
            for (int c2 = 0; c2 < 8; ++c2) {
                double a = M0[c2 * 8 + c];
#pragma unroll
                for (int x = 0; x < c2; ++x) a = fma(-Gm[x * 8 + c2], mcol[x], a);
                mcol[c2] = a * tauS[c2];
                Mm[c2 * 8 + c] = mcol[c2];
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int i = R0 + tid + NT * r;
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
        S1_TICK(4);
        // rotate: (Vp, Wp) <- (Vn, Wn)
        flip ^= 1;
        first = false;
    }
#ifdef S1_TIMING
    if (tid == 0 && blockIdx.x == 0)
        printf("sy2sb cycles: panel-update %lld qr %lld store %lld fused %lld wphase %lld\n", s1_acc[0], s1_acc[1],
               s1_acc[2], s1_acc[3], s1_acc[4]);
#endif
}

#undef Vp
#undef Wp
#undef Vn
#undef Zz

// LD = 8 (mod 16) keeps the 8-lane x 4-k fragment loads of the k-major arrays on distinct banks; >= N + 32 so block
// rows may overhang N (those rows hold zeros).
int sbr_stage1_ld(int N) {
    int ld = N + 32;
    while ((ld & 15) != 8) ++ld;
    return ld;
}
#define S1_NW 16
size_t sbr_stage1_smem(int N) {
    return ((size_t)32 * sbr_stage1_ld(N) + 2 * 8 * 256 + 2 * S1_NW * 8 + 16 + 64 + 8 + 64 + 64 + S1_NW * 64) * sizeof(double);
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
        MCOSB_CUDA(ctx, cudaFuncSetAttribute(sy2sb_kernel<RPT, QMAX, S1_NW>,                                         \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                \
        sy2sb_kernel<RPT, QMAX, S1_NW><<<S, S1_NW * 32, smem, ctx->stream>>>(A, N, LD, tau);                          \
    } while (0)
    if (N <= 256) S1_LAUNCH((256 + S1_NW * 32 - 1) / (S1_NW * 32), 4);
    else if (N <= 512) S1_LAUNCH((512 + S1_NW * 32 - 1) / (S1_NW * 32), 8);
    else if (N <= 768) S1_LAUNCH((768 + S1_NW * 32 - 1) / (S1_NW * 32), 12);
    else return 0;
#undef S1_LAUNCH
    MCOSB_LAUNCHED(ctx, MK_SY2SB);
    return 1;
}
