// D2 (fourth generation), stage 1 of the two-stage tridiagonalisation as a BATCHED, multi-kernel band reduction.
//
// sy2sb_kernel keeps one matrix per SM for its whole reduction: the per-panel chain (panel update -> Householder QR ->
// trailing pass -> W) is serial inside a matrix, the CTA owns all of the SM's registers and shared memory, so nothing else is
// resident to hide the chain and the DMMA pipe idles 74 % of the time (profiles/kernels_r02.md).  Here the same mathematics
// (np.linalg.eigh's front end, covariance_transformer.py:105) is cut at the panel boundaries into two kernels that each run
// over ALL matrices of the wave, so the latency of one matrix's chain hides behind the other thousands:
//
//   for each panel p (columns j1 .. j1+7, R0 = j1 + 8):
//     s3_panel_kernel   one CTA per matrix, one thread per row (RPT rows each):
//         W_{p-1} = Y - 1/2 V (T' V' Y),  Y = Z T,  Z = sum of the partial products the previous pass left      (p > 0)
//         panel columns brought up to date with (V_{p-1}, W_{p-1}); rows j1 .. j1+7 are final (the band)
//         Householder QR of the rows below the band: R -> band, reflectors -> row j1+c of the upper triangle, V_p, tau
//     s3_pass_kernel    one CTA of 4 warps per 64 x 64 tile of the trailing LOWER triangle (grid: tiles x matrices):
//         tile -= V_{p-1} W_{p-1}' + W_{p-1} V_{p-1}'      (4 DMMA.8x8x4 per 8 x 8 block, the tile is the C fragment)
//         store tile
//         Z[rows] += tile  * V_p[cols]                      (2 DMMA, C fragment reused as A operand, k permuted)
//         Z[cols] += tile' * V_p[rows]  (strictly lower)    (2 DMMA, block transposed across the warp by 4 shuffles)
//       A warp holds 16 rows x 64 columns = 16 blocks, all loaded up front (16 independent 16-byte loads per thread, then 16
//       independent DMMA chains), so neither the L2 / HBM latency nor the DMMA dependency latency is exposed the way it is
//       in the one-CTA kernel's ring of two tiles.  Row sums stay in registers across the tile; the column sums of the 4
//       warps are folded through shared memory.  Both go to a slot of the per-matrix partial array Zpart[slot][N][8]:
//       row block R receives slot J from tile (R, J), J <= R, and slot I + 1 from tile (I, R), I >= R -- every (slot <= nT,
//       row) pair is written exactly once per pass, and the panel kernel adds the nT + 1 slots in index order: no atomics,
//       bit-reproducible.
//
// Traffic per matrix: the trailing triangle read + written once per panel, 8 N^3 / 24 B = 41.7 MB at N = 500, plus the
// partial sums and panel arrays (~20 %); at 6.5 TB/s that is ~7.5 us per 500 x 500 matrix against 4.5 us of DMMA work: this
// form is HBM-bound where the one-CTA form was latency-bound at 17.9 us.  Output layout (band in the lower 9 diagonals,
// reflectors in the rows of the upper triangle, tau) is the one sb2st_kernel / backtransform_kernel read.
#include "common.cuh"
#include "dmma.cuh"
#include "sbr.cuh"
#include <algorithm>

#define S3_PANEL_WARPS 8
#ifndef S3_PANEL_PREFETCH
#define S3_PANEL_PREFETCH 1
#endif

__device__ __forceinline__ double s3_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}
__device__ __forceinline__ double s3_rsqrt(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double h = 0.5 * x;
    r = fma(r, fma(-h * r, r, 0.5), r);
    r = fma(r, fma(-h * r, r, 0.5), r);
    return r;
}

// Sum 8 per-lane values over the 32 lanes (halving butterfly, 9 shuffles); lane 4 i ends up owning the total of value i.
__device__ __forceinline__ double s3_reduce8(double (&v)[8], int lane) {
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
    return a1;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Panel kernel.  Row i = j1 + tid + NT r of matrix s; rho = i - R0 = tid - 8 + NT r is its index below the band.
// Vc = V_{p-1} and Zp = partial products of the previous pass (read), Wc = W_{p-1} (written), Vn = V_p (written),
// GT = [64 strictly-upper entries of V'V | 8 tau] of the panel whose W is still to be formed (read, then replaced).
#define S3_PANEL_SMEM(NW) ((size_t)(2 * (NW) * 256 + (NW) * 64 + 3 * 64 + 8 + 2 * (NW) * 8 + 16 + 2 * 64) * sizeof(double))
template <int RPT, int NW>
__global__ void __launch_bounds__(NW * 32, (NW * RPT <= 16 ? 2 : 1))
s3_panel_kernel(double* __restrict__ Aall, int N, int j1, double* __restrict__ tauall, const double* __restrict__ Vc_all,
                double* __restrict__ Wc_all, double* __restrict__ Vn_all, double* __restrict__ GT_all,
                const double* __restrict__ Zp_all, int nslot, int nused, int tr) {
    constexpr int NT = NW * 32;
    extern __shared__ __align__(16) double psm[];
    double (*stV)[32 * 8] = reinterpret_cast<double (*)[32 * 8]>(psm);
    double (*stY)[32 * 8] = reinterpret_cast<double (*)[32 * 8]>(psm + NW * 256);
    double* partM = psm + 2 * NW * 256;
    double* M0 = partM + NW * 64;
    double* Mm = M0 + 64;
    double* Gm = Mm + 64;
    double* tauS = Gm + 64;
    double* red = tauS + 8;
    double* rowc = red + 2 * NW * 8;
    double* pv = rowc + 16;
    double* pw = pv + 64;

    const int s = blockIdx.x;
    double* A = Aall + (size_t)s * N * N;
    double* tau = tauall + (size_t)s * N;
    const double* Vc = Vc_all + (size_t)s * N * 8;
    double* Wc = Wc_all + (size_t)s * N * 8;
    double* Vn = Vn_all + (size_t)s * N * 8;
    double* GT = GT_all + (size_t)s * 72;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const bool first = (j1 == 0);
    const int R0 = j1 + SBR_B;
    const int ncols = min(SBR_B, N - j1);
    const bool last = (N - R0 < 2);

    double v[RPT][8], wv[RPT][8];
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) v[r][c] = wv[r][c] = 0.0;
    // the panel's own columns are needed only after the W phase: their loads go out first
    double p[RPT][8];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int i = j1 + tid + NT * r;
        const double* row = A + (size_t)min(i, N - 1) * N + j1;
#pragma unroll
        for (int c = 0; c < 8; ++c) p[r][c] = __ldcg(row + min(c, ncols - 1));
    }

    if (first) {
        for (int i = tid; i < N; i += NT) tau[i] = 0.0;
    } else {
        // ---- W of the previous panel: Y U = Z, M0 = V' Y, U' M = M0, W = Y - 1/2 V M; U = T^-1 = striu(V'V) + diag(1/tau) ----
        if (tid < 64) Gm[tid] = GT[tid];
        if (tid < 8) tauS[tid] = GT[64 + tid];
        __syncthreads();
        double m0 = 0.0, m1 = 0.0;
#if S3_PANEL_PREFETCH
        // the slots of the row are summed four at a time (register budget), i.e. in up to three dependent DRAM round trips:
        // ask L2 for all of them now, so that the later batches are L2 hits
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int i = j1 + tid + NT * r;
            if (i < N) {
                const double* zp = Zp_all + ((size_t)s * nslot * N + i) * 8;
                const int sl0 = (i - j1) / tr;
                for (int sl = sl0 + 4; sl < nused; ++sl) {
                    const double* q = zp + (size_t)sl * N * 8;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(q + 4));
                }
            }
        }
#endif
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int i = j1 + tid + NT * r;
            if (i < N) {
                double z[8];
                {   // V_{p-1} row: independent of the slot sums, issued ahead of them
                    const double2* vq = reinterpret_cast<const double2*>(Vc + (size_t)i * 8);
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const double2 a = __ldcg(vq + x);
                        v[r][2 * x] = a.x;
                        v[r][2 * x + 1] = a.y;
                    }
                }
#pragma unroll
                for (int x = 0; x < 8; ++x) z[x] = 0.0;
                // the previous pass left the row-side sum of this row in slot 0 and the column-side sums of the tile rows
                // I >= R = (i - j1) / tr in slots I + 1 (tr = rows per tile row of the pass)
                const double* zp = Zp_all + ((size_t)s * nslot * N + i) * 8;
                // four slots in flight at a time; the additions stay in slot order
                const int sl0 = (i - j1) / tr;
                for (int sb = sl0; sb < nused; sb += 4) {
                    double2 buf[4][4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int sl = sb + u;
                        const double2* q2 = reinterpret_cast<const double2*>(zp + (size_t)(sl == sl0 ? 0 : min(sl, nused - 1)) * N * 8);
#pragma unroll
                        for (int x = 0; x < 4; ++x) buf[u][x] = __ldcg(q2 + x);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (sb + u < nused) {
#pragma unroll
                            for (int x = 0; x < 4; ++x) {
                                z[2 * x] += buf[u][x].x;
                                z[2 * x + 1] += buf[u][x].y;
                            }
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) {       // Y U = Z
                    double a = z[c];
#pragma unroll
                    for (int x = 0; x < c; ++x) a = fma(-wv[r][x], Gm[x * 8 + c], a);
                    wv[r][c] = a * tauS[c];
                }
            }
            // this warp's 32 rows of V and Y -> shared memory, then M0 += V' Y as 8 DMMAs (k = rows, 4 at a time)
            __syncwarp();
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                *reinterpret_cast<double2*>(&stV[w][lane * 8 + 2 * x]) = make_double2(v[r][2 * x], v[r][2 * x + 1]);
                *reinterpret_cast<double2*>(&stY[w][lane * 8 + 2 * x]) = make_double2(wv[r][2 * x], wv[r][2 * x + 1]);
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; ++q) dmma_884(m0, m1, stV[w][(4 * q + t) * 8 + g], stY[w][(4 * q + t) * 8 + g]);
        }
        partM[w * 64 + g * 8 + 2 * t] = m0;
        partM[w * 64 + g * 8 + 2 * t + 1] = m1;
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
            const int i = j1 + tid + NT * r;
            if (i < N) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    double a = 0.0;
#pragma unroll
                    for (int x = 0; x < 8; ++x) a = fma(v[r][x], Mm[x * 8 + c], a);
                    wv[r][c] = fma(-0.5, a, wv[r][c]);
                }
                double2* wq = reinterpret_cast<double2*>(Wc + (size_t)i * 8);
#pragma unroll
                for (int x = 0; x < 4; ++x) __stcg(wq + x, make_double2(wv[r][2 * x], wv[r][2 * x + 1]));
            }
        }
        if (tid < 8) {                               // rows j1 .. j1+7: the column side of the panel's own update
#pragma unroll
            for (int x = 0; x < 8; ++x) {
                pv[tid * 8 + x] = v[0][x];
                pw[tid * 8 + x] = wv[0][x];
            }
        }
        __syncthreads();
    }

    // ---- panel columns j1 .. j1+7, rows j1 .. N-1: load, bring up to date with (V_{p-1}, W_{p-1}) ------------------------
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int i = j1 + tid + NT * r;
#pragma unroll
        for (int c = 0; c < 8; ++c) p[r][c] = (i < N && c < ncols && j1 + c <= i) ? p[r][c] : 0.0;
        if (i < N) {
            if (!first) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const double vi = v[r][k], wi = wv[r][k];
#pragma unroll
                    for (int c = 0; c < 8; ++c) p[r][c] = fma(-vi, pw[c * 8 + k], fma(-wi, pv[c * 8 + k], p[r][c]));
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
    if (last) {
        // nothing left to eliminate; at most one more column (the last diagonal element) needs the update
        if (N - R0 == 1 && tid == 8 && !first) {     // tid 8, r = 0 holds row R0 = N - 1
            const int i = N - 1;
            double a = __ldcg(A + (size_t)i * N + i);
#pragma unroll
            for (int k = 0; k < 8; ++k) a = fma(-2.0 * v[0][k], wv[0][k], a);
            __stcg(A + (size_t)i * N + i, a);
        }
        return;
    }

    // ---- Householder QR of the rows >= R0 of the panel --------------------------------------------------------------------
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
        const double tot_lane = s3_reduce8(pr, lane);
        if (t == 0) red[(par * NW + w) * 8 + g] = tot_lane;
        if (tid == c + 8) {
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) rowc[par * 8 + cc] = p[0][cc];
        }
        __syncthreads();
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
            const double rs = s3_rsqrt(n2), nrm = n2 * rs, aa = fabs(alpha);
            beta = -copysign(nrm, alpha);
            tc = fma(aa, rs, 1.0);                               // (beta - alpha) / beta
            scale = copysign(s3_rcp(aa + nrm), alpha);          // 1 / (alpha - beta)
        }
#pragma unroll
        for (int cc = c + 1; cc < 8; ++cc) {
            const double wvv = tc * fma(scale, tot[cc], rowc[par * 8 + cc]);
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const int rho = tid - 8 + NT * r;
                if (rho > c) p[r][cc] = fma(-(p[r][c] * scale), wvv, p[r][cc]);
                else if (rho == c) p[r][cc] -= wvv;
            }
        }
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int rho = tid - 8 + NT * r;
            if (rho > c) p[r][c] *= scale;
            else if (rho == c) p[r][c] = beta;
        }
        if (tid == 64) {                   // G[x][c] = v_x . v_c = v_x[c] + scale * (v_x . x_c below row c), x < c
#pragma unroll
            for (int x = 0; x < 8; ++x)
                if (x < c) GT[x * 8 + c] = fma(scale, tot[x], rowc[par * 8 + x]);
            GT[64 + c] = tc;
            tau[j1 + c] = tc;
        }
    }
    // ---- R to the band, reflectors to the upper triangle and to V_p ------------------------------------------------------
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int rho = tid - 8 + NT * r;
        const int i = R0 + rho;
        if (rho >= 0 && i < N) {
            double vrow[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (rho <= c) {
                    __stcg(A + (size_t)i * N + j1 + c, p[r][c]);
                    vrow[c] = (rho == c) ? 1.0 : 0.0;
                } else {
                    vrow[c] = p[r][c];
                    A[(size_t)(j1 + c) * N + i] = p[r][c];
                }
            }
            double2* vq = reinterpret_cast<double2*>(Vn + (size_t)i * 8);
#pragma unroll
            for (int x = 0; x < 4; ++x) __stcg(vq + x, make_double2(vrow[2 * x], vrow[2 * x + 1]));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Pass kernel.  One CTA of S3_PW warps per (matrix, tile row I of S3_TR = 16 S3_PW rows): it walks the trailing matrix that
// starts at (R0, R0) from column R0 to the diagonal in steps of 32 columns; warp w owns rows i00 + 16 w + [0, 16) = 2 x 4 blocks
// of 8 x 8 per step.  Global -> shared with cp.async through a ring of two stages, issued one full step ahead: a thread copies
// exactly the eight 16-byte pieces it will read back itself (its C fragments), so the tile data needs no barrier; the column
// operands of the step (V_{p-1}, W_{p-1}, V_p rows j0 .. j0+31) are shared by the warps.  Two barriers per step: "stage k
// has landed", and "column sums of the warps are in shared memory / stage k may be refilled".  Steps entirely below the
// diagonal and inside the matrix run a copy of the code without any predicate.
#ifndef S3_PW
#define S3_PW 4
#endif
#ifndef S3_G
#define S3_G 1            // tile rows per CTA of the pass (their column sums share a slot); measured at N = 500: 2 -> panel 5.1
                        // instead of 5.4 us but pass 12.8 instead of 11.6 us (imbalance, pipeline restarts), 4 -> 5.0 and 13.6
#endif
#ifndef S3_L2_AHEAD
#define S3_L2_AHEAD 2      // steps of L2 prefetch ahead of the cp.async ring (0 = off)
#endif
#define S3_TR (16 * S3_PW)
#define S3_PT (32 * S3_PW)
#define P2_COLS 32
#define P2_LDV 12      // row stride of the V / W operand stages ([column][8] + 4): 12 g + t are 16 distinct banks
#define P2_LDN 10      // row stride of the V_p operand stage: 20 t + g are 16 distinct banks
#define P2_LDC 36      // k-major column sums per warp [8][P2_LDC]
#define P2_OPS (2 * P2_COLS * P2_LDV + P2_COLS * P2_LDN)     // doubles per operand stage
#define P2_TILE (S3_PT * 8 * 2)                              // doubles per tile stage
#define P2_SMEM_DOUBLES (2 * P2_OPS + 2 * P2_TILE + S3_PW * 8 * P2_LDC)

template <bool FIRST, bool EVEN, bool MASKED>
__device__ __forceinline__ void s3_issue(int k, int j0, int N, int ib0, int g, int t, int tid, const double* __restrict__ A,
                                         const double* __restrict__ Vc, const double* __restrict__ Wc,
                                         const double* __restrict__ Vn, double* sOps, double* sTile) {
    double* ops = sOps + (k & 1) * P2_OPS;
    for (int q = tid; q < 4 * P2_COLS; q += S3_PT) {
        const int col = q >> 2, pc = q & 3;
        const bool in = !MASKED || j0 + col < N;
        const size_t off = in ? ((size_t)(j0 + col) * 8 + 2 * pc) : 0;
        cp_async16(ops + 2 * P2_COLS * P2_LDV + col * P2_LDN + 2 * pc, Vn + off, in);
        if (!FIRST) {
            cp_async16(ops + col * P2_LDV + 2 * pc, Vc + off, in);
            cp_async16(ops + P2_COLS * P2_LDV + col * P2_LDV + 2 * pc, Wc + off, in);
        }
    }
    double* tile = sTile + (k & 1) * P2_TILE + 2 * tid;
#pragma unroll
    for (int rb = 0; rb < 2; ++rb) {
        const int ib = ib0 + 8 * rb, i = ib + g;
        const double* row = A + (size_t)((!MASKED || i < N) ? i : 0) * N + j0 + 2 * t;
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
            double* dst = tile + (rb * 4 + cb) * (2 * S3_PT);
            if (!MASKED) {
                if (EVEN) cp_async16(dst, row + 8 * cb, true);
                else { cp_async8(dst, row + 8 * cb, true); cp_async8(dst + 1, row + 8 * cb + 1, true); }
            } else {
                const int jb = j0 + 8 * cb, j = jb + 2 * t;
                const bool blk = i < N && jb <= ib;
                if (EVEN) {
                    const bool in = blk && j < N;
                    cp_async16(dst, in ? row + 8 * cb : A, in);
                } else {
                    const bool in0 = blk && j < N, in1 = blk && j + 1 < N;
                    cp_async8(dst, in0 ? row + 8 * cb : A, in0);
                    cp_async8(dst + 1, in1 ? row + 8 * cb + 1 : A, in1);
                }
            }
        }
    }
}

template <bool FIRST, bool EVEN, bool MASKED>
__device__ __forceinline__ void s3_step(int k, int j0, int N, int ib0, int g, int t, int tid, int w, double* __restrict__ A,
                                        const double* sOps, const double* sTile, double* sZc, const double (&av)[2][4],
                                        const double (&vni)[2][2], double (&zi)[2][2]) {
    const double* ops = sOps + (k & 1) * P2_OPS;
    const double* sV = ops, *sW = ops + P2_COLS * P2_LDV, *sVn = ops + 2 * P2_COLS * P2_LDV;
    const double* tile = sTile + (k & 1) * P2_TILE + 2 * tid;
    double c[2][4][2];
#pragma unroll
    for (int rb = 0; rb < 2; ++rb)
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
            const double2 v2 = *reinterpret_cast<const double2*>(tile + (rb * 4 + cb) * (2 * S3_PT));
            c[rb][cb][0] = v2.x;
            c[rb][cb][1] = v2.y;
        }
    // transposing shuffles: in round A the lanes of rows 0-3 publish their first element and rows 4-7 their second, in
    // round B the other way round; a lane with even g takes (h = 0 from A, h = 1 from B), with odd g (h = 0 from B, h = 1 from A)
    const bool lowrow = g < 4, godd = g & 1;
    const int srcA = 4 * ((godd ? 4 : 0) + t) + (g >> 1), srcB = 4 * ((godd ? 0 : 4) + t) + (g >> 1);
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
        const int jb = j0 + 8 * cb;
        double zc0 = 0.0, zc1 = 0.0;
        const double vnj0 = sVn[(8 * cb + 2 * t) * P2_LDN + g], vnj1 = sVn[(8 * cb + 2 * t + 1) * P2_LDN + g];
        double bw0 = 0.0, bw1 = 0.0, bv0 = 0.0, bv1 = 0.0;
        if (!FIRST) {
            bw0 = sW[(8 * cb + g) * P2_LDV + t];
            bw1 = sW[(8 * cb + g) * P2_LDV + 4 + t];
            bv0 = sV[(8 * cb + g) * P2_LDV + t];
            bv1 = sV[(8 * cb + g) * P2_LDV + 4 + t];
        }
#pragma unroll
        for (int rb = 0; rb < 2; ++rb) {
            const int ib = ib0 + 8 * rb;
            // blocks entirely above the diagonal or below row N contribute nothing (warp-uniform)
            if (!MASKED || (jb <= ib && ib < N)) {
                const int i = ib + g, j = jb + 2 * t;
                double u0 = c[rb][cb][0], u1 = c[rb][cb][1];
                const bool p0 = !MASKED || (i < N && j <= i), p1 = !MASKED || (i < N && j + 1 <= i);
                if (MASKED) {                            // pieces come whole: drop what lies above the diagonal
                    u0 = p0 ? u0 : 0.0;
                    u1 = p1 ? u1 : 0.0;
                }
                if (!FIRST) {
                    dmma_884(u0, u1, av[rb][0], bw0);
                    dmma_884(u0, u1, av[rb][1], bw1);
                    dmma_884(u0, u1, av[rb][2], bv0);
                    dmma_884(u0, u1, av[rb][3], bv1);
                    double* dst = A + (size_t)i * N + j;
                    if (MASKED) {
                        u0 = p0 ? u0 : 0.0;
                        u1 = p1 ? u1 : 0.0;
                        if (p0) __stcg(dst, u0);
                        if (p1) __stcg(dst + 1, u1);
                    } else if (EVEN) {
                        __stcg(reinterpret_cast<double2*>(dst), make_double2(u0, u1));
                    } else {
                        __stcg(dst, u0);
                        __stcg(dst + 1, u1);
                    }
                }
                // Z[rows] += block * V_p[cols]: k = block columns, permuted (2 t | 2 t + 1) to match the C fragment
                dmma_884(zi[rb][0], zi[rb][1], u0, vnj0);
                dmma_884(zi[rb][0], zi[rb][1], u1, vnj1);
                // Z[cols] += block' * V_p[rows], strictly lower part: B operand element (k = t -> row 4 h + t, n = g -> column g)
                double l0 = u0, l1 = u1;
                if (MASKED) {
                    l0 = (j < i) ? u0 : 0.0;
                    l1 = (j + 1 < i) ? u1 : 0.0;
                }
                const double ra = __shfl_sync(0xffffffffu, lowrow ? l0 : l1, srcA);
                const double rb2 = __shfl_sync(0xffffffffu, lowrow ? l1 : l0, srcB);
                dmma_884(zc0, zc1, vni[rb][0], godd ? rb2 : ra);
                dmma_884(zc0, zc1, vni[rb][1], godd ? ra : rb2);
            }
        }
        *reinterpret_cast<double2*>(sZc + (w * 8 + g) * P2_LDC + 8 * cb + 2 * t) = make_double2(zc0, zc1);
    }
}

template <bool FIRST, bool EVEN>
__global__ void __launch_bounds__(S3_PT, (S3_PW == 4 ? 3 : (S3_PW == 2 ? 5 : 2)))
s3_pass_kernel(double* __restrict__ Aall, int N, int R0, const double* __restrict__ Vc_all, const double* __restrict__ Wc_all,
               const double* __restrict__ Vn_all, double* __restrict__ Zp_all, int nslot, int nT) {
    extern __shared__ __align__(16) double sm[];
    double* sOps = sm;                          // [2][P2_OPS]: V | W | V_p
    double* sTile = sm + 2 * P2_OPS;            // [2][8 pieces][threads][2]
    double* sZc = sTile + 2 * P2_TILE;          // [warps][8][P2_LDC]
    const int s = blockIdx.y;
    const int Ig = (nT + S3_G - 1) / S3_G - 1 - blockIdx.x;     // group of S3_G consecutive tile rows, long ones first
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    double* A = Aall + (size_t)s * N * N;
    const double* Vc = Vc_all + (size_t)s * N * 8;
    const double* Wc = Wc_all + (size_t)s * N * 8;
    const double* Vn = Vn_all + (size_t)s * N * 8;
    double* zrow = Zp_all + ((size_t)s * nslot + 0) * N * 8;
    double* zcol = Zp_all + ((size_t)s * nslot + Ig + 1) * N * 8;
    // The tile rows of a group are walked one after the other by the same CTA, and their column sums go to ONE slot: from
    // the second tile row on, a step adds its sums to what the same thread stored for the same columns one walk earlier
    // (fixed order, so still reproducible) -- the panel kernel has S3_G times fewer slots to read.
    for (int sub = 0; sub < S3_G; ++sub) {
    const int I = S3_G * Ig + sub;
    if (I >= nT) break;
    const int nprev = (sub == 0) ? 0 : (S3_TR / P2_COLS) * I;   // steps whose columns the previous walk has already summed
    const int i00 = R0 + S3_TR * I;
    const int ib0 = i00 + 16 * w;
    const int nsteps = (S3_TR / P2_COLS) * (I + 1);
    // steps whose 32 columns all lie below row i00, in a tile row that does not overhang N, need no predicate
    const int nfree = (i00 + S3_TR > N) ? 0 : (S3_TR / P2_COLS) * I;

    auto issue = [&](int k) {
        if (k < nsteps) {
            const int j0 = R0 + P2_COLS * k;
            if (k < nfree) s3_issue<FIRST, EVEN, false>(k, j0, N, ib0, g, t, tid, A, Vc, Wc, Vn, sOps, sTile);
            else s3_issue<FIRST, EVEN, true>(k, j0, N, ib0, g, t, tid, A, Vc, Wc, Vn, sOps, sTile);
        }
        cp_async_commit();
    };
    issue(0);
    issue(1);

    // row-side operands straight from the panel arrays (each used by this warp only, for the whole tile row)
    double av[2][4], vni[2][2];
#pragma unroll
    for (int rb = 0; rb < 2; ++rb) {
        const int i = ib0 + 8 * rb + g;
        const bool in = i < N;
        av[rb][0] = av[rb][1] = av[rb][2] = av[rb][3] = 0.0;
        if (!FIRST && in) {
            av[rb][0] = -__ldg(Vc + (size_t)i * 8 + t);
            av[rb][1] = -__ldg(Vc + (size_t)i * 8 + 4 + t);
            av[rb][2] = -__ldg(Wc + (size_t)i * 8 + t);
            av[rb][3] = -__ldg(Wc + (size_t)i * 8 + 4 + t);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ir = ib0 + 8 * rb + 4 * h + t;
            vni[rb][h] = ir < N ? __ldg(Vn + (size_t)ir * 8 + g) : 0.0;
        }
    }
    double zi[2][2];
#pragma unroll
    for (int rb = 0; rb < 2; ++rb) zi[rb][0] = zi[rb][1] = 0.0;

    // L2 prefetch of this warp's 16 x 32 piece of step k: one 128-byte half row per lane
    auto l2_prefetch = [&](int k) {
        const int row = ib0 + (lane >> 1), col = R0 + P2_COLS * k + 16 * (lane & 1);
        if (k < nsteps && row < N && col + 15 <= row && col + 15 < N)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(A + (size_t)row * N + col));
    };
#if S3_L2_AHEAD > 0
#pragma unroll
    for (int k = 2; k < 2 + S3_L2_AHEAD; ++k) l2_prefetch(k);
#endif

    for (int k = 0; k < nsteps; ++k) {
        const int j0 = R0 + P2_COLS * k;
#if S3_L2_AHEAD > 0
        l2_prefetch(k + 2 + S3_L2_AHEAD);                     // the cp.async of step k + 2 is issued at the end of this step
#endif
        cp_async_wait<1>();
        __syncthreads();                                      // stage k & 1 has landed (own pieces and the shared operands)
        if (k < nfree) s3_step<FIRST, EVEN, false>(k, j0, N, ib0, g, t, tid, w, A, sOps, sTile, sZc, av, vni, zi);
        else s3_step<FIRST, EVEN, true>(k, j0, N, ib0, g, t, tid, w, A, sOps, sTile, sZc, av, vni, zi);
        __syncthreads();                                      // column sums written; every warp is done with stage k & 1
        issue(k + 2);
        // column sums of the warps, in warp order -> slot I + 1, rows j0 .. j0 + 31
        for (int idx = tid; idx < P2_COLS * 8; idx += S3_PT) {
            const int jl = idx >> 3, kk = idx & 7;
            if (j0 + jl < N) {
                double a = sZc[(0 * 8 + kk) * P2_LDC + jl];
#pragma unroll
                for (int ww = 1; ww < S3_PW; ++ww) a += sZc[(ww * 8 + kk) * P2_LDC + jl];
                if (k < nprev) a += __ldcg(zcol + (size_t)j0 * 8 + idx);
                __stcg(zcol + (size_t)j0 * 8 + idx, a);
            }
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int rb = 0; rb < 2; ++rb) {
        const int i = ib0 + 8 * rb + g;
        if (i < N) __stcg(reinterpret_cast<double2*>(zrow + (size_t)i * 8 + 2 * t), make_double2(zi[rb][0], zi[rb][1]));
    }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
static int s3_nslot(int N) { return (N - SBR_B + S3_G * S3_TR - 1) / (S3_G * S3_TR) + 1; }

// extra device bytes per matrix (wave sizing in run.cu)
size_t sbr3_scratch_bytes(int N) { return ((size_t)(3 + s3_nslot(N)) * N * 8 + 72) * sizeof(double); }

bool sbr3_supported(mcosb_ctx* ctx) {
    const int N = ctx->N;
    return N >= SBR_B + 2 && N <= 1024;
}

// Returns 1 if launched, 0 if N is outside the range of these kernels.
int mcosb_launch_sy2sb3(mcosb_ctx* ctx, int S, double* A, double* tau) {
    const int N = ctx->N;
    if (!sbr3_supported(ctx)) return 0;
    if (S > 32768) {                     // the matrix index is blockIdx.y of the pass
        for (int s0 = 0; s0 < S; s0 += 32768) {
            const int rc = mcosb_launch_sy2sb3(ctx, std::min(32768, S - s0), A + (size_t)s0 * N * N, tau + (size_t)s0 * N);
            if (rc != 1) return rc;
        }
        return 1;
    }
    const size_t pass_smem = (size_t)P2_SMEM_DOUBLES * sizeof(double);
    if (!ctx->s3_attr_set) {   // per context (function attributes belong to its device): three 60 KB pass CTAs per SM
#define S3_ATTR(F, E)                                                                                                  \
        MCOSB_CUDA(ctx, cudaFuncSetAttribute(s3_pass_kernel<F, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pass_smem)); \
        MCOSB_CUDA(ctx, cudaFuncSetAttribute(s3_pass_kernel<F, E>, cudaFuncAttributePreferredSharedMemoryCarveout, 100))
        S3_ATTR(true, true); S3_ATTR(true, false); S3_ATTR(false, true); S3_ATTR(false, false);
#undef S3_ATTR
        MCOSB_CUDA(ctx, cudaFuncSetAttribute(s3_panel_kernel<2, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)S3_PANEL_SMEM(16)));
        ctx->s3_attr_set = true;
    }
    const int nslot = s3_nslot(N);
    double *Vb, *Wb, *GT, *Zp;
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_S3_V, (size_t)2 * S * N * 8, &Vb));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_S3_W, (size_t)S * N * 8, &Wb));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_S3_GT, (size_t)S * 72, &GT));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_S3_ZP, (size_t)S * nslot * N * 8, &Zp));
    double* Vbuf[2] = {Vb, Vb + (size_t)S * N * 8};
    static const int stop_after = getenv("MCOSB_S3_STOP") ? atoi(getenv("MCOSB_S3_STOP")) : -1;   // debugging: launches to run
    int launched = 0;
    int nused = 0;
    for (int p = 0, j1 = 0;; ++p, j1 += SBR_B) {
        const int R0 = j1 + SBR_B;
        const bool last = (N - R0 < 2);
        double* Vc = Vbuf[(p + 1) & 1];
        double* Vn = Vbuf[p & 1];
        if (stop_after >= 0 && launched >= stop_after) break;
#define S3_PANEL(RPT, NW)                                                                                              \
        s3_panel_kernel<RPT, NW><<<S, NW * 32, S3_PANEL_SMEM(NW), ctx->stream>>>(A, N, j1, tau, Vc, Wb, Vn, GT, Zp, nslot, nused, S3_G * S3_TR)
        if (N <= 256) S3_PANEL(1, 8);
        else if (N <= 512) S3_PANEL(2, 8);
        else S3_PANEL(2, 16);
#undef S3_PANEL
        MCOSB_LAUNCHED(ctx, MK_SY2SB_PANEL);
        ++launched;
        if (last) break;
        if (stop_after >= 0 && launched >= stop_after) break;
        const int m = N - R0;
        const int nT = (m + S3_TR - 1) / S3_TR;
        const dim3 grid((nT + S3_G - 1) / S3_G, S);
        const bool even = (N & 1) == 0;
#define S3_PASS(F, E) s3_pass_kernel<F, E><<<grid, S3_PT, pass_smem, ctx->stream>>>(A, N, R0, Vc, Wb, Vn, Zp, nslot, nT)
        if (p == 0) { if (even) S3_PASS(true, true); else S3_PASS(true, false); }
        else { if (even) S3_PASS(false, true); else S3_PASS(false, false); }
#undef S3_PASS
        MCOSB_LAUNCHED(ctx, MK_SY2SB_PASS);
        ++launched;
        nused = (nT + S3_G - 1) / S3_G + 1;
    }
    return 1;
}
