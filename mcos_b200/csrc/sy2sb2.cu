// D2 (fourth generation), stage 1 of the two-stage tridiagonalisation: dense symmetric -> band (half bandwidth 8).
//
// Same mathematics as sy2sb.cu (panel QR, ONE fused pass over the trailing lower triangle that applies the previous
// panel's rank-16 update and accumulates Z = A V for the new one, W from Z by triangular solves), restructured around
// what the round-1 profile showed (profiles/sbr_kernels_r01c.md: DMMA pipe 26 %, one 16-warp CTA per SM, 18 % of the
// stalls on the per-block-row barrier, 20 % on exposed L2 latency, and 30 % of the kernel in serial panel phases with
// nothing else resident on the SM):
//
//   * TWO matrices per SM (N <= 512): 8 warps per matrix, three k-major panel arrays (V_p, W_p, V_n) in shared memory
//     instead of four -- Z never lives in shared memory any more -- so the panel QR / W phase of one matrix hides under
//     the trailing pass of the other;
//   * NO barrier inside the pass: a warp owns the column blocks cb = warp (mod 8) as before, but writes its partial row
//     sums of Z for every 16-row block row to a per-matrix scratch in global memory (L2-resident: 8 x 8 x N doubles,
//     rewritten every panel) instead of meeting the other seven warps at a __syncthreads per block row; the W phase folds
//     the eight partials (fixed order: deterministic) while it reads its row of Z;
//   * tiles prefetched with cp.async into thread-private shared-memory slots (4 ahead) instead of a register ring of 2:
//     no registers, no barrier, and no scoreboard shared with the operand loads of the tile in work;
//   * N up to 1024 with one 16-warp CTA per SM (two row groups x 8 column classes, the column blocks of a class taken in
//     two sub-passes of 8 so the column-side accumulators stay at 16 registers): config 5 (N = 1000) no longer falls
//     back to the one-stage kernel.
//
// Traffic by design: one read + one write of the trailing triangle per panel = 8 N^3 / (3 * 8) bytes (42 MB at N = 500)
// + 2 x 64 N x 8 B of Z partials per panel (L2).  Flops 4 N^3 / 3 on DMMA.8x8x4.  Deterministic (no atomics).
#include <cstdlib>

#include "common.cuh"
#include "dmma.cuh"
#include "sbr.cuh"

#ifndef S2_D
#define S2_D 2    // register-ring variant (SD == 0): tiles in flight per warp, a power of two
#endif
#ifndef S2_VNI_REG
#define S2_VNI_REG 0   // 1: keep the V_n[rows] fragments of a block row in registers (8 more)
#endif
#ifndef S2_MINB
#define S2_MINB(NW) ((NW) == 8 ? 2 : 1)
#endif

#ifdef S2_TIMING
#define S2_TICK(idx) do { if (threadIdx.x == 0) { long long now__ = clock64(); s2_acc[idx] += now__ - s2_last; s2_last = now__; } } while (0)
#else
#define S2_TICK(idx) do { } while (0)
#endif

// Sum 8 per-lane values over the 32 lanes (halving butterfly, 9 shuffles); lane 4 i ends up owning the total of value i.
__device__ __forceinline__ double s2_reduce8(double (&v)[8], int lane) {
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
    return a1;   // value index = lane >> 2
}

// 1/x and 1/sqrt(x) to ~1 ulp: hardware approximation (2^-23) + two Newton steps
__device__ __forceinline__ double s2_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}
__device__ __forceinline__ double s2_rsqrt(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double h = 0.5 * x;
    r = fma(r, fma(-h * r, r, 0.5), r);
    r = fma(r, fma(-h * r, r, 0.5), r);
    return r;
}

// NW = 8 RG warps: warp w works on column blocks cb = 8 q + (w & 7) and on the 16-row block rows n = (w >> 3) (mod RG).
// RPT = rows per thread in the panel phases (ceil(N / threads)); QH = sub-passes of 8 column blocks per class (ceil(N / 512)).
// SD > 0: tiles are prefetched with cp.async into thread-private shared-memory slots, SD tiles ahead of their use (each
// lane copies exactly the 2 x 16 bytes it will read itself: no barrier, no registers held, and -- unlike loads into a
// register ring -- no scoreboard shared with the LDS / DMMA traffic of the tile being worked on).  SD == 0: register ring of
// S2_D tiles (the arrangement for N > 640, where shared memory has no room for the slots).
template <int NW, int RPT, int QH, int SD, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
sy2sb2_kernel(double* Aall, int N, int LD, int LDZ, double* __restrict__ tauall, double* __restrict__ Zp_all,
              double* __restrict__ Zc_all) {
    constexpr int NT = NW * 32;
    constexpr int RG = NW / 8;
    constexpr int NCLS = 8 * QH;
    extern __shared__ double sm[];
    // three k-major panel arrays [8][LD] (element (row i, k) at [k * LD + i]): W_p in the middle, V_p and V_n swap ends every
    // panel; their addresses are derived from `flip` at the point of use (rotating pointer variables cost spilled registers)
    int flip = 0;
#define Vp (sm + (flip ? 16 * LD : 0))
#define Wp (sm + 8 * LD)
#define Vn (sm + (flip ? 0 : 16 * LD))
    double* red = sm + 24 * (size_t)LD;       // [2][NW][8]
    double* rowc = red + 2 * NW * 8;          // [2][8]
    double* Gm = rowc + 16;                   // [8][8] strictly upper part of V'V (T^-1 = striu(V'V) + diag(1/tau))
    double* tauS = Gm + 64;                   // [8]
    double* M0 = tauS + 8;                    // [8][8]
    double* Mm = M0 + 64;                     // [8][8]
    double* partM = Mm + 64;                  // [NW][64]
    double* ringS = partM + NW * 64;          // [NW][SD][128]  (SD > 0): lane's pieces at [2 lane] (rows g) and [64 + 2 lane] (rows g + 8)

    const int s = blockIdx.x;
    double* A = Aall + (size_t)s * N * N;
    double* tau = tauall + (size_t)s * N;
    double* Zp = Zp_all + (size_t)s * NCLS * LDZ * 8;   // [NCLS][LDZ][8]  partial row sums of Z per column class
    double* Zc = Zc_all + (size_t)s * RG * LDZ * 8;     // [RG][LDZ][8]    column sums of Z (one owner per column block and row group)
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wc = w & 7, wh = w >> 3;
    const bool even = (N & 1) == 0;

    for (int i = tid; i < 24 * LD; i += NT) sm[i] = 0.0;
    for (int i = tid; i < N; i += NT) tau[i] = 0.0;
    __syncthreads();

#ifdef S2_TIMING
    long long s2_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long s2_last = clock64();
#endif
    for (int j1 = 0;; j1 += SBR_B) {
        const int R0 = j1 + SBR_B;           // first row / column of the new trailing matrix
        const int ncols = min(SBR_B, N - j1);
        // ---- a. panel columns j1 .. j1+7, rows j1 .. N-1: update with (Vp, Wp) (zero before the first panel) -----------
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
                    if (c < ncols && j1 + c <= i) p[r][c] = __ldcg(row + c);
                if (j1 > 0) {
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
                        if (c < ncols && j1 + c <= i) __stcg(wrow + c, p[r][c]);
                }
            }
        }
        S2_TICK(0);
        if (last) {
            // nothing left to eliminate; at most one more column (the last diagonal element) needs the update
            if (N - R0 == 1 && tid == 0 && j1 > 0) {
                const int i = N - 1;
                double a = __ldcg(A + (size_t)i * N + i);
                for (int k = 0; k < 8; ++k) a = fma(-2.0 * Vp[k * LD + i], Wp[k * LD + i], a);
                __stcg(A + (size_t)i * N + i, a);
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
            const double tot_lane = s2_reduce8(pr, lane);
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
                const double rs = s2_rsqrt(n2), nrm = n2 * rs, aa = fabs(alpha);
                beta = -copysign(nrm, alpha);
                tc = fma(aa, rs, 1.0);                               // (beta - alpha) / beta
                scale = copysign(s2_rcp(aa + nrm), alpha);          // 1 / (alpha - beta)
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
        S2_TICK(1);
        // ---- R to the band, reflectors to the upper triangle and to Vn -------------------------------------------------
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int rho = tid - 8 + NT * r;
            const int i = R0 + rho;
            if (rho >= 0 && i < N) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (rho <= c) {
                        __stcg(A + (size_t)i * N + j1 + c, p[r][c]);
                        Vn[c * LD + i] = (rho == c) ? 1.0 : 0.0;
                    } else {
                        Vn[c * LD + i] = p[r][c];
                        A[(size_t)(j1 + c) * N + i] = p[r][c];
                    }
                }
            }
        }
        __syncthreads();   // Vn, Gm complete; panel stores visible
        S2_TICK(2);

        // ---- b. ONE pass over the trailing lower triangle, no barrier inside ------------------------------------------
        // tile (n, q) of this warp: rows R0 + 16 n + [0, 16), n = wh (mod RG); column block cb = 8 q + wc
        const int nbr = (N - R0 + 15) >> 4;
        const int rowstep = 8 * N;
        const int cK0 = t * LD + 8 * wc + g, cK4 = (4 + t) * LD + 8 * wc + g;   // fragment offsets for column block q = 0
        const int cVn = g * LD + 8 * wc + 2 * t;
        auto ring_load = [&](double (&dst)[2][2], const double* ptr) {
#pragma unroll
            for (int rb = 0; rb < 2; ++rb) {
                const double2 v2 = __ldcg(reinterpret_cast<const double2*>(ptr + rb * rowstep));
                dst[rb][0] = v2.x;
                dst[rb][1] = v2.y;
            }
        };
#pragma unroll 1
        for (int h = 0; h < QH; ++h) {
            const int qbase = 8 * h;
            int qlo = ((R0 >> 3) - wc + 7) >> 3;                 // first q whose column block lies at or beyond R0
            if (qlo < qbase) qlo = qbase;
            auto qhi_of = [&](int n) -> int {        // last q of this sub-pass that reaches the lower triangle in block row n
                int lastrow = R0 + 16 * n + 15;
                if (lastrow > N - 1) lastrow = N - 1;
                const int cbl = lastrow >> 3;
                int qh = (cbl - wc) >= 0 ? ((cbl - wc) >> 3) : -1;
                return qh > qbase + 7 ? qbase + 7 : qh;
            };
            auto qfast_of = [&](int n) -> int {      // last q entirely below the diagonal blocks, all 16 rows inside N
                const int i0 = R0 + 16 * n;
                if (!even || n >= nbr || i0 + 15 >= N) return -1;
                const int num = (i0 >> 3) - 1 - wc;
                return num >= 0 ? (num >> 3) : -1;
            };
            double zc[8][2];
#pragma unroll
            for (int q = 0; q < 8; ++q) zc[q][0] = zc[q][1] = 0.0;
            auto qf_of = [&](int n) -> int { return min(qfast_of(n), qhi_of(n)); };
            // ---- tile prefetch -----------------------------------------------------------------------------------------------
            // SD > 0: the warp's fast tiles form one sequence (owned block rows in order, q ascending inside a row); the
            // producer cursor (pn, pq) runs SD tiles ahead of the consumer, slot = sequence number mod SD, one commit group per
            // sequence number (empty once the producer has run out), so cp.async.wait_group SD-1 before tile k guarantees group k.
            double* myslot = ringS + (size_t)w * (SD > 0 ? SD : 1) * 128 + 2 * lane;
            int pn = wh, pq = qlo;
            unsigned pseq = 0, cseq = 0;
            auto issue = [&]() {
                if (SD > 0) {
                    if (pn < nbr) {
                        const double* src = A + (size_t)(R0 + 16 * pn + g) * N + 8 * wc + 2 * t + 64 * pq;
                        double* dst = myslot + (pseq % (SD > 0 ? SD : 1)) * 128;
                        cp_async16(dst, src, true);
                        cp_async16(dst + 64, src + rowstep, true);
                        if (pq < qf_of(pn)) ++pq;
                        else {
                            pn += RG;
                            pq = qlo;
                            if (pn < nbr && qf_of(pn) < qlo) pn = nbr;     // only the overhanging last block row: no fast tiles
                        }
                    }
                    cp_async_commit();
                    ++pseq;
                }
            };
            if (SD > 0) {
                while (pn < nbr && qf_of(pn) < qlo) pn += RG;              // first owned block row with a fast tile
#pragma unroll
                for (int d = 0; d < (SD > 0 ? SD : 1); ++d) issue();
            }
            // SD == 0: ring of S2_D tiles in registers; slot (q - qbase) % S2_D holds the next unprocessed fast tile with that
            // residue.  Every load is unconditional (clamped address) and lands directly in the ring: the ring registers are only
            // read (C input of an out-of-place DMMA) before they are refilled, so ptxas keeps the loads in flight.
            double ring[S2_D][2][2];
            const double* tile00 = A + (size_t)(R0 + g) * N + 8 * wc + 2 * t;   // tile (0, 0); (n, q): + 16 n N + 64 q
            if (SD == 0) {
                const int qf0 = qf_of(wh);
#pragma unroll
                for (int d = 0; d < S2_D; ++d) {
                    const int qd = qlo + ((d - qlo) & (S2_D - 1));
                    if (qd <= qf0) ring_load(ring[d], tile00 + (size_t)16 * wh * N + 64 * qd);
                }
            }
            for (int n = wh; n < nbr; n += RG) {
                const int i0 = R0 + 16 * n;
                const int qhi = qhi_of(n);
                const int qf = min(qfast_of(n), qhi);
                const int qfn = (n + RG < nbr) ? min(qfast_of(n + RG), qhi_of(n + RG)) : -1;
                double* rowbase = A + (size_t)(i0 + g) * N + 8 * wc + 2 * t;
                const double* nextbase = rowbase + (size_t)16 * RG * N;
                // (register ring) slots this block row does not use take their tile of the next block row now
                if (SD == 0) {
#pragma unroll
                    for (int d = 0; d < S2_D; ++d) {
                        const int qd = qlo + ((d - qlo) & (S2_D - 1));
                        if (qd > qf && qd <= qfn) ring_load(ring[d], nextbase + 64 * qd);
                    }
                }
                double zi[2][2];
#pragma unroll
                for (int rb = 0; rb < 2; ++rb) zi[rb][0] = zi[rb][1] = 0.0;
                if (qlo <= qhi) {
                    double av[2][4];
#if S2_VNI_REG
                    double vni[2][2];
#define S2_VNI(rb, hh) vni[rb][hh]
#else
                    const double* vnrow = Vn + g * LD + i0 + t;      // V_n[rows] fragments re-read per tile: 8 registers less
#define S2_VNI(rb, hh) vnrow[8 * (rb) + 4 * (hh)]
#endif
#pragma unroll
                    for (int rb = 0; rb < 2; ++rb) {
                        const int i = i0 + 8 * rb + g;
                        av[rb][0] = -Vp[t * LD + i];
                        av[rb][1] = -Vp[(4 + t) * LD + i];
                        av[rb][2] = -Wp[t * LD + i];
                        av[rb][3] = -Wp[(4 + t) * LD + i];
#if S2_VNI_REG
                        vni[rb][0] = Vn[g * LD + i0 + 8 * rb + t];
                        vni[rb][1] = Vn[g * LD + i0 + 8 * rb + 4 + t];
#endif
                    }
                    // ---- tiles entirely below the diagonal -----------------------------------------------------------------
#pragma unroll
                    for (int ql = 0; ql < 8; ++ql) {
                        const int q = qbase + ql;
                        if (q >= qlo && q <= qf) {
                            const double bw0 = Wp[cK0 + 64 * q], bw1 = Wp[cK4 + 64 * q];
                            const double bv0 = Vp[cK0 + 64 * q], bv1 = Vp[cK4 + 64 * q];
                            const double2 vnj = *reinterpret_cast<const double2*>(Vn + cVn + 64 * q);
                            double u[2][2];
                            if (SD > 0) {
                                cp_async_wait<(SD > 0 ? SD - 1 : 0)>();
                                const double* slot = myslot + (cseq % (SD > 0 ? SD : 1)) * 128;
                                const double2 c0 = *reinterpret_cast<const double2*>(slot);
                                const double2 c1 = *reinterpret_cast<const double2*>(slot + 64);
                                ++cseq;
                                dmma_884_oop(u[0][0], u[0][1], av[0][0], bw0, c0.x, c0.y);
                                dmma_884_oop(u[1][0], u[1][1], av[1][0], bw0, c1.x, c1.y);
                                issue();
                            } else {
                                double (&cur)[2][2] = ring[ql % S2_D];
                                // refill address: same block row if it has a tile q + D, else this slot's tile of the next block row,
                                // else (nothing left for the slot) the current tile again -- always valid memory
                                const int q2 = qlo + ((q - qlo) & (S2_D - 1));
                                const double* refill = (q + S2_D <= qf) ? rowbase + 64 * (q + S2_D)
                                                                         : ((q2 <= qfn) ? nextbase + 64 * q2 : rowbase + 64 * q);
#pragma unroll
                                for (int rb = 0; rb < 2; ++rb) dmma_884_oop(u[rb][0], u[rb][1], av[rb][0], bw0, cur[rb][0], cur[rb][1]);
                                ring_load(cur, refill);
                            }
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
#pragma unroll
                            for (int rb = 0; rb < 2; ++rb) dmma_884(zi[rb][0], zi[rb][1], u[rb][0], vnj.x);
#pragma unroll
                            for (int rb = 0; rb < 2; ++rb) dmma_884(zi[rb][0], zi[rb][1], u[rb][1], vnj.y);
                            // Z[cols] += tile' * Vn[rows]: B operand element (k = t -> row 4h + t, n = g -> column g)
#pragma unroll
                            for (int rb = 0; rb < 2; ++rb) {
#pragma unroll
                                for (int hh = 0; hh < 2; ++hh) {
                                    const int src = 4 * (4 * hh + t) + (g >> 1);
                                    const double e0 = __shfl_sync(0xffffffffu, u[rb][0], src);
                                    const double e1 = __shfl_sync(0xffffffffu, u[rb][1], src);
                                    dmma_884(zc[ql][0], zc[ql][1], S2_VNI(rb, hh), (g & 1) ? e1 : e0);
                                }
                            }
                        }
                    }
                    // ---- tiles that touch the diagonal or overhang N: element-wise predicates -------------------------------
#pragma unroll
                    for (int ql = 0; ql < 8; ++ql) {
                        const int q = qbase + ql;
                        if (q >= qlo && q > qf && q <= qhi) {
                            const int j0 = 8 * (8 * q + wc);
                            double cur[2][2], lt[2][2];
#pragma unroll
                            for (int rb = 0; rb < 2; ++rb) {
                                const int i = i0 + 8 * rb + g, j = j0 + 2 * t;
                                cur[rb][0] = (i < N && j <= i) ? __ldcg(A + (size_t)i * N + j) : 0.0;
                                cur[rb][1] = (i < N && j + 1 <= i) ? __ldcg(A + (size_t)i * N + j + 1) : 0.0;
                            }
                            const double bw0 = Wp[cK0 + 64 * q], bw1 = Wp[cK4 + 64 * q];
                            const double bv0 = Vp[cK0 + 64 * q], bv1 = Vp[cK4 + 64 * q];
                            const double2 vnj = *reinterpret_cast<const double2*>(Vn + cVn + 64 * q);
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
                                if (i < N && j <= i) __stcg(A + (size_t)i * N + j, cur[rb][0]);
                                if (i < N && j + 1 <= i) __stcg(A + (size_t)i * N + j + 1, cur[rb][1]);
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
                                for (int hh = 0; hh < 2; ++hh) {
                                    const int src = 4 * (4 * hh + t) + (g >> 1);
                                    const double e0 = __shfl_sync(0xffffffffu, lt[rb][0], src);
                                    const double e1 = __shfl_sync(0xffffffffu, lt[rb][1], src);
                                    dmma_884(zc[ql][0], zc[ql][1], S2_VNI(rb, hh), (g & 1) ? e1 : e0);
                                }
                            }
                        }
                    }
                }
                // partial row sums of this class for the block row (zeros when it had no tile): [cls][row][k], k = 2t, 2t+1
                {
                    double* zp = Zp + ((size_t)(qbase + wc) * LDZ + i0 + g) * 8 + 2 * t;
#pragma unroll
                    for (int rb = 0; rb < 2; ++rb)
                        __stcg(reinterpret_cast<double2*>(zp + rb * 64), make_double2(zi[rb][0], zi[rb][1]));
                }
            }
            // column sums of this sub-pass: one owner per (column block, row group), plane wh of Zc
            {
                double* zcw = Zc + (size_t)wh * LDZ * 8;
#pragma unroll
                for (int ql = 0; ql < 8; ++ql) {
                    const int j0 = 8 * (8 * (qbase + ql) + wc);
                    if (j0 >= R0 && j0 < N) {
                        __stcg(zcw + (size_t)(j0 + 2 * t) * 8 + g, zc[ql][0]);
                        __stcg(zcw + (size_t)(j0 + 2 * t + 1) * 8 + g, zc[ql][1]);
                    }
                }
            }
        }
        __syncthreads();
        S2_TICK(3);

        // ---- c. W = Z T - 1/2 V (T' V' Z T), T^-1 = striu(G) + diag(1/tau): triangular solves instead of forming T -------
        // Z[i][:] = column sums + the NCLS class partials, in a fixed order.  Y goes to the dead W_p array.
        double y[RPT][8];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int i = R0 + tid + NT * r;
#pragma unroll
            for (int c = 0; c < 8; ++c) y[r][c] = 0.0;
            if (i < N) {
                double z[8];
                {
                    const double2* zr = reinterpret_cast<const double2*>(Zc + (size_t)i * 8);
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const double2 v2 = __ldcg(zr + x);
                        z[2 * x] = v2.x;
                        z[2 * x + 1] = v2.y;
                    }
                }
                if (RG > 1) {
#pragma unroll
                    for (int hh = 1; hh < RG; ++hh) {
                        const double2* zr = reinterpret_cast<const double2*>(Zc + ((size_t)hh * LDZ + i) * 8);
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            const double2 v2 = __ldcg(zr + x);
                            z[2 * x] += v2.x;
                            z[2 * x + 1] += v2.y;
                        }
                    }
                }
#pragma unroll 4
                for (int cls = 0; cls < NCLS; ++cls) {
                    const double2* zr = reinterpret_cast<const double2*>(Zp + ((size_t)cls * LDZ + i) * 8);
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const double2 v2 = __ldcg(zr + x);
                        z[2 * x] += v2.x;
                        z[2 * x + 1] += v2.y;
                    }
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) {       // Y U = Z
                    double a = z[c];
#pragma unroll
                    for (int x = 0; x < c; ++x) a = fma(-y[r][x], Gm[x * 8 + c], a);
                    y[r][c] = a * tauS[c];
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) Wp[c * LD + i] = y[r][c];
            }
        }
        __syncthreads();
        {
            double m0 = 0.0, m1 = 0.0;
            for (int i4 = R0 + 4 * w; i4 < N; i4 += 4 * NW) dmma_884(m0, m1, Vn[g * LD + i4 + t], Wp[g * LD + i4 + t]);
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
                    Wp[c * LD + i] = fma(-0.5, a, y[r][c]);
                }
            }
        }
        __syncthreads();
        S2_TICK(4);
        // rotate: V_p <- V_n, W_p stays where W_n was just written, the old V_p array becomes the next V_n
        flip ^= 1;
    }
#ifdef S2_TIMING
    if (tid == 0 && blockIdx.x == 0)
        printf("sy2sb2 cycles: panel-update %lld qr %lld store %lld fused %lld wphase %lld\n", s2_acc[0], s2_acc[1],
               s2_acc[2], s2_acc[3], s2_acc[4]);
#endif
}

#undef Vp
#undef Wp
#undef Vn

// LD = 8 (mod 16) keeps the 16-byte fragment loads of the k-major arrays on distinct bank groups (MCOSB_S2_LDMOD for
// experiments); >= N + 16 so the last 16-row block row may overhang N (those rows hold zeros).
static int s2_env(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
int sbr2_ld(int N) {
    static const int mod = s2_env("MCOSB_S2_LDMOD", 8) & 15;
    int ld = N + 16;
    while ((ld & 15) != mod) ++ld;
    return ld;
}
static int sbr2_ldz(int N) { return (N + 16 + 7) & ~7; }

// Kernel configuration by problem size: warps per matrix, rows per thread in the panel phases, column sub-passes, depth of
// the cp.async tile ring (0 = register ring) and matrices per SM.
struct S2Config { int nw, rpt, qh, sd, minb; };
static S2Config sbr2_config(int N) {
    static const int force16 = s2_env("MCOSB_S2_FORCE16", 0);      // A/B: one 16-warp CTA per SM also for small N
    static const int sd16 = s2_env("MCOSB_S2_SD16", 4);           // ring depth of the 16-warp variant (4 or 6)
    if (N <= 256 && !force16) return {8, 1, 1, 4, 2};
    if (N <= 440 && !force16) return {8, 2, 1, 2, 2};
    if (N <= 512) return {16, 1, 1, sd16 == 6 ? 6 : 4, 1};
    if (N <= 640) return {16, 2, 2, 4, 1};
    return {16, 2, 2, 0, 1};
}
size_t sbr2_smem(int N) {
    const S2Config c = sbr2_config(N);
    return ((size_t)24 * sbr2_ld(N) + 2 * c.nw * 8 + 16 + 64 + 8 + 64 + 64 + c.nw * 64 + (size_t)c.nw * c.sd * 128) * sizeof(double);
}
bool sbr2_supported(mcosb_ctx* ctx) {
    const int N = ctx->N;
    return N >= SBR_B + 2 && N <= 1024 && sbr2_smem(N) <= (size_t)ctx->smem_optin;
}

template <int NW, int RPT, int QH, int SD, int MINB>
static int s2_launch(mcosb_ctx* ctx, int S, double* A, double* tau, size_t smem, int LD, int LDZ, double* Zp, double* Zc) {
    static const bool trace = getenv("MCOSB_TRACE") != nullptr;
    auto kern = sy2sb2_kernel<NW, RPT, QH, SD, MINB>;
    MCOSB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // two CTAs per SM need the whole 228 KB as shared memory: ask for the maximum carve-out explicitly
    MCOSB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    if (trace) {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, NW * 32, smem);
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, kern);
        fprintf(stderr, "sy2sb2<%d,%d,%d,%d,%d>: N %d LD %d smem %zu regs %d local %zu -> %d CTAs per SM\n", NW, RPT, QH, SD,
                MINB, ctx->N, LD, smem, fa.numRegs, (size_t)fa.localSizeBytes, nb);
    }
    kern<<<S, NW * 32, smem, ctx->stream>>>(A, ctx->N, LD, LDZ, tau, Zp, Zc);
    return MCOSB_OK;
}

// Returns 1 if launched, 0 if N is outside the range of this kernel.
int mcosb_launch_sy2sb2(mcosb_ctx* ctx, int S, double* A, double* tau) {
    const int N = ctx->N;
    if (!sbr2_supported(ctx)) return 0;
    const S2Config c = sbr2_config(N);
    const size_t smem = sbr2_smem(N);
    const int LD = sbr2_ld(N), LDZ = sbr2_ldz(N);
    double *Zp, *Zc;
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_S2_ZP, (size_t)S * 8 * c.qh * LDZ * 8, &Zp));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_S2_ZC, (size_t)S * (c.nw / 8) * LDZ * 8, &Zc));
    if (c.nw == 8 && c.rpt == 1) MCOSB_TRY((s2_launch<8, 1, 1, 4, 2>(ctx, S, A, tau, smem, LD, LDZ, Zp, Zc)));
    else if (c.nw == 8) MCOSB_TRY((s2_launch<8, 2, 1, 2, 2>(ctx, S, A, tau, smem, LD, LDZ, Zp, Zc)));
    else if (c.rpt == 1 && c.sd == 6) MCOSB_TRY((s2_launch<16, 1, 1, 6, 1>(ctx, S, A, tau, smem, LD, LDZ, Zp, Zc)));
    else if (c.rpt == 1) MCOSB_TRY((s2_launch<16, 1, 1, 4, 1>(ctx, S, A, tau, smem, LD, LDZ, Zp, Zc)));
    else if (c.sd == 4) MCOSB_TRY((s2_launch<16, 2, 2, 4, 1>(ctx, S, A, tau, smem, LD, LDZ, Zp, Zc)));
    else MCOSB_TRY((s2_launch<16, 2, 2, 0, 1>(ctx, S, A, tau, smem, LD, LDZ, Zp, Zc)));
    MCOSB_LAUNCHED(ctx, MK_SY2SB);
    return 1;
}
