// K4b-d: HRPOptimizer.allocate (optimizer.py:198-287), batched.
//
//   hrp_dist_kernel    corr = clip(cov / sigma sigma'),  D = sqrt((1 - corr)/2), D_ii = 0      (optimizer.py:212-214, 281-287)
//   hrp_euclid_kernel  E_ij = sqrt(sum_k (D_ik - D_jk)^2)  -- what scipy's linkage() computes when it is handed the
//                      SQUARE matrix D (optimizer.py:216): pdist between its rows.  Accumulated sequentially over k with
//                      separate multiply and add (no FMA) so E is bit-identical to scipy and the merge order is too.
//   hrp_tree_kernel    one CTA per simulation: Prim's MST from node 0 (lowest index wins ties), stable sort of the
//                      edges, union-find relabelling (scipy mst_single_linkage + label), depth-first leaf order with the
//                      column-0 child first (optimizer.py:235-249), level-synchronous recursive bisection
//                      (optimizer.py:256-279).  Weights are returned in POSITION order, as the reference does.
//
// Fast path (default): only the MERGE ORDER and the N - 1 merge heights have to be scipy's bits, not all N^2 / 2
// distances.  hrp_gram_kernel forms approximate squared distances n_i + n_j - 2 D_i . D_j on the FP64 tensor cores
// (N^3 flop instead of 1.5 N^3 non-fused instructions); Prim runs on them and refuses to decide whenever two candidates
// are closer than twice a rigorous bound of the approximation error (then the simulation is flagged and redone with the
// exact distance matrix, bit for bit the old path); the heights of the accepted edges are recomputed with scipy's
// arithmetic.  Unflagged simulations therefore produce exactly the exact path's linkage.
#include "common.cuh"
#include "dmma.cuh"

// ---------------------------------------------------------------------------------------------------------------
#ifndef HD_ROWS
#define HD_ROWS 32   // rows per CTA (a multiple of 8)
#endif
#ifndef HD_UNROLL
#define HD_UNROLL 8  // independent row loads per lane
#endif
// grid = (ceil(N / HD_ROWS), S), HD_ROWS / 8 rows per warp; the square roots of the diagonal are taken once per CTA.
// (1 - r) / 2 is an exact halving, so the multiplication by 0.5 rounds like numpy's division.
__global__ void __launch_bounds__(256) hrp_dist_kernel(const double* __restrict__ cov, int N, double* __restrict__ D,
                                                       double* __restrict__ rn) {
    extern __shared__ double sg[];   // [N]
    const int s = blockIdx.y;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double* c = cov + (size_t)s * N * N;
    double* d = D + (size_t)s * N * N;
    for (int j = threadIdx.x; j < N; j += 256) sg[j] = sqrt(c[(size_t)j * N + j]);
    __syncthreads();
    for (int q = 0; q < HD_ROWS / 8; ++q) {
        const int row = blockIdx.x * HD_ROWS + wid * (HD_ROWS / 8) + q;
        if (row >= N) return;
        const double si = sg[row];
        double nrm = 0.0;
        // 8 independent loads per lane ahead of the divisions / square roots (memory-level parallelism); the additions into
        // nrm keep their column order
        for (int j0 = lane; j0 < N; j0 += 32 * HD_UNROLL) {
            double x[HD_UNROLL];
#pragma unroll
            for (int u = 0; u < HD_UNROLL; ++u) {
                const int j = j0 + 32 * u;
                x[u] = (j < N) ? c[(size_t)row * N + j] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < HD_UNROLL; ++u) {
                const int j = j0 + 32 * u;
                if (j < N) {
                    double r = __ddiv_rn(x[u], __dmul_rn(si, sg[j]));
                    r = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
                    double v = sqrt(__dmul_rn(__dsub_rn(1.0, r), 0.5));
                    v = (j == row) ? 0.0 : v;
                    d[(size_t)row * N + j] = v;
                    nrm = fma(v, v, nrm);
                }
            }
        }
        if (rn) {                       // squared row norms for the Gram form of the distances (fast path)
            nrm = warp_sum(nrm);
            if (lane == 0) rn[(size_t)s * N + row] = nrm;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 64 x 64 output tile per CTA, 16 x 16 threads, 4 x 4 outputs per thread; k streamed through shared memory in slabs of
// 32, stored k-major so a thread's four j (or i) values are contiguous.  grid = (lower tiles, S).
// 3 FP64 instructions (DADD, DMUL, DADD) per (i, j, k): the HRP floor is 1.5 N^3 non-fused FP64 ops per simulation.
// ---------------------------------------------------------------------------------------------------------------
#define EU_T 64
#define EU_K 32
#define EU_LD (EU_T + 2)

__global__ void __launch_bounds__(256) hrp_euclid_kernel(const double* __restrict__ Dall, int N,
                                                         double* __restrict__ Eall, const int* __restrict__ only) {
    __shared__ double Si[EU_K][EU_LD];
    __shared__ double Sj[EU_K][EU_LD];
    const int s = blockIdx.y;
    if (only && !only[s]) return;      // fallback pass: flagged simulations only
    int I = (int)((sqrt(8.0 * blockIdx.x + 1.0) - 1.0) * 0.5);
    while ((I + 1) * (I + 2) / 2 <= (int)blockIdx.x) ++I;
    while (I * (I + 1) / 2 > (int)blockIdx.x) --I;
    const int J = blockIdx.x - I * (I + 1) / 2;
    const double* D = Dall + (size_t)s * N * N;
    double* E = Eall + (size_t)s * N * N;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int lk = tid & 31, lr = tid >> 5;  // loader: k = lk, rows lr + 8 r
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    for (int k0 = 0; k0 < N; k0 += EU_K) {
        const int k = k0 + lk;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int ri = I * EU_T + lr + 8 * r, rj = J * EU_T + lr + 8 * r;
            Si[lk][lr + 8 * r] = (k < N && ri < N) ? D[(size_t)ri * N + k] : 0.0;
            Sj[lk][lr + 8 * r] = (k < N && rj < N) ? D[(size_t)rj * N + k] : 0.0;
        }
        __syncthreads();
        const int kmax = min(EU_K, N - k0);
        if (I != J) {
            for (int kk = 0; kk < kmax; ++kk) {
                double a[4], b[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) { a[q] = Si[kk][ty + 16 * q]; b[q] = Sj[kk][tx + 16 * q]; }  // lane-contiguous: no bank conflicts
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        double df = __dsub_rn(a[p], b[q]);
                        acc[p][q] = __dadd_rn(acc[p][q], __dmul_rn(df, df));
                    }
            }
        } else {
            // diagonal tile: outputs (ty + 16 p, tx + 16 q) with q > p lie above the diagonal for every thread
            for (int kk = 0; kk < kmax; ++kk) {
                double a[4], b[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) { a[q] = Si[kk][ty + 16 * q]; b[q] = Sj[kk][tx + 16 * q]; }
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int q = 0; q <= p; ++q) {
                        double df = __dsub_rn(a[p], b[q]);
                        acc[p][q] = __dadd_rn(acc[p][q], __dmul_rn(df, df));
                    }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int i = I * EU_T + ty + 16 * p, j = J * EU_T + tx + 16 * q;
            if (i < N && j < N && (I != J || j <= i)) {
                double v = sqrt(acc[p][q]);
                E[(size_t)i * N + j] = v;
                E[(size_t)j * N + i] = v;
            }
        }
}

// ---------------------------------------------------------------------------------------------------------------
// Fast path: E2[i][j] ~ |D_i - D_j|^2 = n_i + n_j - 2 D_i . D_j with the Gram matrix on the FP64 tensor cores.
// grid = (2 x lower 128 x 128 tiles, S); both operands are row tiles of D ([row][k] in shared memory, as the TRMM's).
// |E2 - exact| <= 2 g (n_i + n_j), g = N * 2^-53 (rounding of the N-term dot products and norms); hrp_tree_kernel<true>
// works with four times that.  Mirrored write: Prim reads rows.
// ---------------------------------------------------------------------------------------------------------------
#define GR_THREADS 128
#define GR_STAGE ((DM_BM + DM_BN / 2) * DM_LD_MMAJOR)   // doubles per pipeline stage: A (128 rows) + B (64 rows)
template <bool VEC, bool TAKE_SQRT>
__global__ void __launch_bounds__(GR_THREADS, 2)
hrp_gram_kernel(const double* __restrict__ Dall, const double* __restrict__ rnall, int N, double* __restrict__ E2all) {
    // two CTAs of 4 warps per 128 x 128 tile, as in syrk_kernel: the column halves of an off-diagonal tile, the two warp
    // sets of the sub-block dealing on a diagonal one
    extern __shared__ double sm[];
    const int s = blockIdx.y;
    int I, J;
    dm_tile_ij(blockIdx.x >> 1, I, J);
    const int half = blockIdx.x & 1;
    const bool diag = (I == J);
    const double* D = Dall + (size_t)s * N * N;
    const double* rn = rnall + (size_t)s * N;
    double* E2 = E2all + (size_t)s * N * N;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int wm0 = (wid >> 1) * DM_WM, wn0 = (wid & 1) * DM_WN;
    const int jrow0 = J * DM_BN + half * (DM_BN / 2);    // first row of D in this CTA's B operand
    const int dwid = half * 4 + wid;
    constexpr int LW = VEC ? 2 : 1, LRS = GR_THREADS * LW / DM_BK;
    const int lk = (tid % (DM_BK / LW)) * LW, lr = tid / (DM_BK / LW);

    double acc[DM_MB][DM_NB][2];
#pragma unroll
    for (int a = 0; a < DM_MB; ++a)
#pragma unroll
        for (int b = 0; b < DM_NB; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    auto issue = [&](int kt, int buf) {
        double* At = sm + (size_t)buf * GR_STAGE;
        double* Bt = At + DM_TILE_MMAJOR;
        const int k = kt * DM_BK + lk;
        const bool kok = k < N;
#pragma unroll
        for (int r = 0; r < DM_BM / LRS; ++r) {
            const int row = lr + LRS * r;
            const int ia = I * DM_BM + row;
            const bool aok = kok && ia < N;
            if (VEC) cp_async16(&At[row * DM_LD_MMAJOR + lk], D + (aok ? (size_t)ia * N + k : 0), aok);
            else cp_async8(&At[row * DM_LD_MMAJOR + lk], D + (aok ? (size_t)ia * N + k : 0), aok);
        }
        if (!diag) {
#pragma unroll
            for (int r = 0; r < DM_BN / 2 / LRS; ++r) {
                const int row = lr + LRS * r;
                const int ib = jrow0 + row;
                const bool bok = kok && ib < N;
                if (VEC) cp_async16(&Bt[row * DM_LD_MMAJOR + lk], D + (bok ? (size_t)ib * N + k : 0), bok);
                else cp_async8(&Bt[row * DM_LD_MMAJOR + lk], D + (bok ? (size_t)ib * N + k : 0), bok);
            }
        }
    };
    const int nkt = (N + DM_BK - 1) / DM_BK;
#pragma unroll
    for (int st = 0; st < DM_STAGES - 1; ++st) {
        if (st < nkt) issue(st, st);
        cp_async_commit();
    }
    int r0a = 0, c0a = 0, r0b = 0, c0b = 0;     // diagonal tiles: the lower triangle as sub-blocks (dmma_slab_diag)
    if (diag) dmma_diag_subtiles(dwid, r0a, c0a, r0b, c0b);
    for (int kt = 0; kt < nkt; ++kt) {
        cp_async_wait<DM_STAGES - 2>();
        __syncthreads();
        const double* At = sm + (size_t)(kt % DM_STAGES) * GR_STAGE;
        auto next = [&] {
            if (kt + DM_STAGES - 1 < nkt) issue(kt + DM_STAGES - 1, (kt + DM_STAGES - 1) % DM_STAGES);
            cp_async_commit();
        };
        if (diag) dmma_slab_diag<false, false>(At, dwid, lane, acc, next);
        else dmma_slab<false, false, false>(At, At + DM_TILE_MMAJOR, wm0, wn0, lane, acc, next);
    }
    const int g = lane >> 2, t4 = lane & 3;
    const bool two = diag && dwid >= 6;
#pragma unroll
    for (int mb = 0; mb < DM_MB; ++mb) {
        // diagonal tiles: accumulator rows 0..3 are sub-tile a, rows 4..7 sub-tile b (warps 6, 7 only)
        const int row = I * DM_BM + (diag ? (mb < 4 ? r0a + mb * 8 : r0b + (mb - 4) * 8) : wm0 + mb * 8) + g;
        const double nr = (row < N) ? rn[row] : 0.0;
#pragma unroll
        for (int nb = 0; nb < DM_NB; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = (diag ? J * DM_BN + (mb < 4 ? c0a : c0b) : jrow0 + wn0) + nb * 8 + 2 * t4 + e;
                if (row < N && col < N && col <= row && (!diag || mb < 4 || two)) {
                    double v = fma(-2.0, acc[mb][nb][e], nr + rn[col]);
                    v = (col == row || v < 0.0) ? 0.0 : v;     // NaN passes through and sends the simulation to the fallback
                    if (TAKE_SQRT) v = sqrt(v);                 // NCO: distances, not squares (mcosb_launch_gram_dist)
                    E2[(size_t)row * N + col] = v;
                    E2[(size_t)col * N + row] = v;
                }
            }
    }
}

// ---------------------------------------------------------------------------------------------------------------
#ifndef TR_THREADS
#define TR_THREADS 256
#endif

// ---------------------------------------------------------------------------------------------------------------
// Fast path, N <= 1024: scipy's Prim on the approximate squares with ONE WARP per simulation.  best[] / bestx[] / the
// merged set live in registers (lane l owns nodes l, l + 32, ...), the argmin is a shuffle butterfly, there is no shared
// memory and no barrier, and many simulations are in flight per SM to hide the latency of the row load each step depends
// on (the block-per-simulation version spent 2.6-3.5 M cycles per simulation here: three barriers and a serial
// section per step).  Same guarded decisions as hrp_tree_kernel<1>; the first ambiguous one flags the simulation and
// ends the warp.  edges[s][0][k], [1][k], [2][k] = x, y and the tree node that realises best[y] of step k.
// ---------------------------------------------------------------------------------------------------------------
#ifndef PR_WARPS
#define PR_WARPS 2   // measured (us per simulation at N = 500): 2 -> 0.58, 4 -> 0.66, 8 -> 0.85
#endif
template <int MAXM>
__global__ void __launch_bounds__(PR_WARPS * 32) hrp_prim_kernel(const double* __restrict__ E2all,
                                                                 const double* __restrict__ rnall, int N, int S,
                                                                 unsigned short* __restrict__ edges,
                                                                 int* __restrict__ slow, int force_slow,
                                                                 int* __restrict__ slow_count) {
    const int s = blockIdx.x * PR_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (s >= S) return;
    if (N < 2) {
        if (lane == 0) slow[s] = 0;
        return;
    }
    const double* E2 = E2all + (size_t)s * N * N;
    unsigned short* eg = edges + (size_t)s * 3 * N;
    double rmax = 0.0;
    for (int i = lane; i < N; i += 32) rmax = fmax(rmax, rnall[(size_t)s * N + i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
    const double delta2 = 3.0 * (16.0 * (double)N * 1.1102230246251565e-16 * rmax);
    int amb = (!(rmax == rmax) || !(rmax < INFINITY)) ? 1 : force_slow;
    double best[MAXM];
    int bx[MAXM];
    unsigned mergedmask = 0;            // bit m: node lane + 32 m is in the tree (or does not exist)
#pragma unroll
    for (int m = 0; m < MAXM; ++m) {
        best[m] = INFINITY;
        bx[m] = 0;
        if (lane + 32 * m >= N) mergedmask |= 1u << m;
    }
    int x = 0;
    for (int k = 0; k < N - 1 && !amb; ++k) {
        if ((x & 31) == lane) mergedmask |= 1u << (x >> 5);
        const double* row = E2 + (size_t)x * N;
        double mv = INFINITY, m2 = INFINITY;
        int mi = N;
        // the whole row first, unconditionally (clamped index): behind the per-node predicate every load would sit in its
        // own branch and expose its latency, 16 times per step
        double dv[MAXM];
#pragma unroll
        for (int m = 0; m < MAXM; ++m) dv[m] = row[min(lane + 32 * m, N - 1)];
#pragma unroll
        for (int m = 0; m < MAXM; ++m) {
            if (!((mergedmask >> m) & 1u)) {
                const double dd = dv[m];
                double b = best[m];
                if (!(dd == dd)) amb = 1;
                if (b > dd) {
                    if (b - dd <= delta2) amb = 1;     // two tree nodes (nearly) equally close to this node
                    b = dd; best[m] = b; bx[m] = x;
                } else if (dd - b <= delta2) {
                    amb = 1;
                }
                if (b < mv) { m2 = mv; mv = b; mi = lane + 32 * m; }
                else if (b < m2) m2 = b;
            }
        }
        int bxv = 0;                     // bx of this lane's best candidate
#pragma unroll
        for (int m = 0; m < MAXM; ++m)
            if (mi == lane + 32 * m) bxv = bx[m];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, mv, o), o2 = __shfl_xor_sync(0xffffffffu, m2, o);
            const int oi = __shfl_xor_sync(0xffffffffu, mi, o), ob = __shfl_xor_sync(0xffffffffu, bxv, o);
            const bool take = ov < mv || (ov == mv && oi < mi);
            m2 = fmin(fmin(m2, o2), take ? mv : ov);
            if (take) { mv = ov; mi = oi; bxv = ob; }
        }
        amb = __any_sync(0xffffffffu, amb) ? 1 : 0;
        if (!(m2 - mv > delta2) || mi >= N) amb = 1;   // runner-up within the error bound, or nothing finite left
        if (lane == 0) {
            eg[k] = (unsigned short)x;
            eg[N + k] = (unsigned short)mi;
            eg[2 * N + k] = (unsigned short)bxv;
        }
        x = mi;
    }
    if (lane == 0) {
        slow[s] = amb;
        if (amb) atomicAdd(slow_count, 1);
    }
}

// MODE 1 (FAST): Eall holds hrp_gram_kernel's approximate SQUARED distances; decisions closer than the error bound flag
// the simulation in slow[] (and end the CTA); merge heights are recomputed from D with scipy's arithmetic.
// MODE 2 (FAST, N <= 1024): Prim has already run in hrp_prim_kernel (one warp per simulation); its edges come from
// `edges`, unflagged simulations continue with the merge heights.
// MODE 0: Eall holds exact distances (hrp_euclid_kernel); with `only`, just the simulations flagged by the fast pass.
template <int MODE>
__global__ void __launch_bounds__(TR_THREADS, 8) hrp_tree_kernel(const double* __restrict__ Eall,
                                                              const double* __restrict__ cov_all, int N,
                                                              double* __restrict__ w_all, int* __restrict__ order_all,
                                                              double* __restrict__ link_all, int* __restrict__ status,
                                                              const double* __restrict__ Dall,
                                                              const double* __restrict__ rnall, int* __restrict__ slow,
                                                              const int* __restrict__ only, int force_slow,
                                                              int* __restrict__ slow_count,
                                                              const unsigned short* __restrict__ edges) {
    constexpr bool FAST = MODE != 0;
    extern __shared__ double sm[];
    // doubles
    double* best = sm;              // [N]   Prim: distance to the tree;  later: position weights
    double* edist = best + N;       // [N]   edge lengths in discovery order
    double* sdist = edist + N;      // [N]   sorted edge lengths
    double* red = sdist + N;        // [32]
    int* ired = reinterpret_cast<int*>(red + 32);   // [32]
    // node and cluster ids (< 2N <= 65535) as 16-bit: 46 N bytes per simulation instead of 84 N, so 8 CTAs fit an SM -- the
    // kernel is a chain of block-wide steps and lives on CTA-level parallelism
    typedef unsigned short u16;
    u16* ea = reinterpret_cast<u16*>(ired + 32);    // [N] edge endpoint x (discovery order)
    u16* eb = ea + N;               // [N] edge endpoint y
    u16* sa = eb + N;               // [N] sorted endpoints / linkage column 0
    u16* sb = sa + N;               // [N]                   linkage column 1
    u16* parent = sb + N;           // [2N] union-find
    u16* csize = parent + 2 * N;    // [2N] cluster sizes
    u16* order = csize + 2 * N;     // [N]
    u16* merged = order + N;        // [N]
    u16* bestx = merged + N;        // [N]   FAST: the tree node that realises best[i]
    __shared__ int s_x;
    __shared__ double red2[32];

    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int NW = TR_THREADS / 32;
    const double* E = Eall + (size_t)s * N * N;
    if (!FAST && only && !only[s]) return;

    if (N == 1) {
        if (tid == 0) {
            order_all[s] = 0;
            if (FAST && slow) slow[s] = 0;
        }
        return;
    }
    if (MODE == 2) {
        // edges (x, y, tree node realising best[y]) of hrp_prim_kernel; flagged simulations never get here
        if (slow[s]) return;
        const unsigned short* eg = edges + (size_t)s * 3 * N;
        for (int k = tid; k < N - 1; k += TR_THREADS) {
            ea[k] = eg[k];
            eb[k] = eg[N + k];
            bestx[eg[N + k]] = eg[2 * N + k];
        }
        __syncthreads();
    }
    // FAST: decisions need a gap of 3 delta, delta = 16 N u rmax being twice the worst-case bound on |approximate -
    // scipy's| squared distance (hrp_gram_kernel); the third delta keeps the square roots scipy compares distinct
    double delta2 = 0.0;
    int amb = 0;
    if (MODE != 2) {
    if (FAST) {
        double rmax = 0.0;
        for (int i = tid; i < N; i += TR_THREADS) rmax = fmax(rmax, rnall[(size_t)s * N + i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
        if (lane == 0) red2[wid] = rmax;
        __syncthreads();
        rmax = red2[0];
        for (int q = 1; q < NW; ++q) rmax = fmax(rmax, red2[q]);
        delta2 = 3.0 * (16.0 * (double)N * 1.1102230246251565e-16 * rmax);
        if (!(rmax == rmax) || !(rmax < INFINITY)) amb = 1;
        __syncthreads();
    }
#ifdef HRP_TIMING
    long long tq0 = clock64(), tq1 = 0, tq2 = 0, tq3 = 0;
#endif
    // ---- Prim's MST from node 0 (scipy _hierarchy.mst_single_linkage) -------------------------------------------------
    for (int i = tid; i < N; i += TR_THREADS) { best[i] = INFINITY; merged[i] = 0; }
    __syncthreads();
    int x = 0;
    for (int k = 0; k < N - 1; ++k) {
        if (tid == 0) merged[x] = 1;
        __syncthreads();
        const double* row = E + (size_t)x * N;
        double mv = INFINITY, m2 = INFINITY;     // smallest and (FAST) second smallest candidate
        int mi = N;
        for (int i = tid; i < N; i += TR_THREADS) {
            if (!merged[i]) {
                double dd = row[i];
                double b = best[i];
                if (FAST) {
                    if (!(dd == dd)) amb = 1;
                    if (b > dd) {
                        if (b - dd <= delta2) amb = 1;     // two tree nodes (nearly) equally close to i
                        b = dd; best[i] = b; bestx[i] = x;
                    } else if (dd - b <= delta2) {
                        amb = 1;
                    }
                    if (b < mv) { m2 = mv; mv = b; mi = i; }
                    else if (b < m2) m2 = b;
                } else {
                    if (b > dd) { b = dd; best[i] = b; }
                    if (b < mv) { mv = b; mi = i; }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, mv, o);
            int oi = __shfl_xor_sync(0xffffffffu, mi, o);
            if (FAST) {
                const double o2 = __shfl_xor_sync(0xffffffffu, m2, o);
                const bool take = ov < mv || (ov == mv && oi < mi);
                m2 = fmin(fmin(m2, o2), take ? mv : ov);
                if (take) { mv = ov; mi = oi; }
            } else if (ov < mv || (ov == mv && oi < mi)) { mv = ov; mi = oi; }
        }
        if (lane == 0) { red[wid] = mv; ired[wid] = mi; if (FAST) red2[wid] = m2; }
        __syncthreads();
        if (tid == 0) {
            double bv = red[0];
            int bi = ired[0];
            if (FAST) {
                double b2 = red2[0];
                for (int q = 1; q < NW; ++q) {
                    const bool take = red[q] < bv || (red[q] == bv && ired[q] < bi);
                    b2 = fmin(fmin(b2, red2[q]), take ? bv : red[q]);
                    if (take) { bv = red[q]; bi = ired[q]; }
                }
                if (!(b2 - bv > delta2)) amb = 1;    // the runner-up is within the error bound (or nothing finite is left)
            } else {
                for (int q = 1; q < NW; ++q)
                    if (red[q] < bv || (red[q] == bv && ired[q] < bi)) { bv = red[q]; bi = ired[q]; }
            }
            if (bi >= N) {  // only NaN/inf distances left: take the lowest unmerged index
                for (int i = 0; i < N; ++i)
                    if (!merged[i] && i != x) { bi = i; break; }
            }
            ea[k] = x; eb[k] = bi; edist[k] = bv;
            s_x = bi;
        }
        __syncthreads();
        x = s_x;
    }
    }
    const int M = N - 1;
#ifdef HRP_TIMING
    tq1 = clock64();
#endif
    if (FAST) {
        if (MODE == 1) {
            const int flagged = __syncthreads_or(amb | force_slow);
            if (tid == 0) {
                slow[s] = flagged;
                if (flagged) atomicAdd(slow_count, 1);
            }
            if (flagged) return;                // redone by hrp_euclid_kernel + hrp_tree_kernel<0>
        }
        // merge heights with scipy's arithmetic: sequential over k, separate subtract / multiply / add, then sqrt;
        // D[y] of mst_single_linkage is the distance to the tree node that realised the minimum
        const double* D = Dall + (size_t)s * N * N;
        for (int k = tid; k < M; k += TR_THREADS) {
            const int b = eb[k], a = bestx[b];
            const double* ra = D + (size_t)max(a, b) * N;
            const double* rb = D + (size_t)min(a, b) * N;
            double acc = 0.0;
            if ((N & 3) == 0) {
                // rows are 32-byte aligned: one whole sector per lane and load (LDG.256).  Every lane walks its own two rows,
                // so a load instruction is 32 sector look-ups whatever its width, and the kernel is bound by exactly that
                // (L1 wavefronts: ~1 us per simulation with 16-byte loads)
                int kk = 0;
                for (; kk + 8 <= N; kk += 8) {
                    double va[8], vb[8];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                                     : "=d"(va[4 * q]), "=d"(va[4 * q + 1]), "=d"(va[4 * q + 2]), "=d"(va[4 * q + 3])
                                     : "l"(ra + kk + 4 * q));
                        asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                                     : "=d"(vb[4 * q]), "=d"(vb[4 * q + 1]), "=d"(vb[4 * q + 2]), "=d"(vb[4 * q + 3])
                                     : "l"(rb + kk + 4 * q));
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const double df = __dsub_rn(va[q], vb[q]);
                        acc = __dadd_rn(acc, __dmul_rn(df, df));
                    }
                }
                for (; kk < N; ++kk) {
                    const double df = __dsub_rn(ra[kk], rb[kk]);
                    acc = __dadd_rn(acc, __dmul_rn(df, df));
                }
            } else if ((N & 1) == 0) {          // rows are 16-byte aligned: half the load instructions (each touches 32 lines)
                const double2* pa = reinterpret_cast<const double2*>(ra);
                const double2* pb = reinterpret_cast<const double2*>(rb);
                // every lane walks its own two rows (32 sectors per load instruction): eight loads in flight per lane ahead of
                // the dependent chain of additions, whose order stays k = 0, 1, 2, ...
                int kk = 0;
                for (; kk + 4 <= N / 2; kk += 4) {
                    double2 va[4], vb[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) { va[q] = pa[kk + q]; vb[q] = pb[kk + q]; }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const double d0 = __dsub_rn(va[q].x, vb[q].x), d1 = __dsub_rn(va[q].y, vb[q].y);
                        acc = __dadd_rn(acc, __dmul_rn(d0, d0));
                        acc = __dadd_rn(acc, __dmul_rn(d1, d1));
                    }
                }
                for (; kk < N / 2; ++kk) {
                    const double2 va = pa[kk], vb = pb[kk];
                    const double d0 = __dsub_rn(va.x, vb.x), d1 = __dsub_rn(va.y, vb.y);
                    acc = __dadd_rn(acc, __dmul_rn(d0, d0));
                    acc = __dadd_rn(acc, __dmul_rn(d1, d1));
                }
            } else {
                for (int kk = 0; kk < N; ++kk) {
                    const double df = __dsub_rn(ra[kk], rb[kk]);
                    acc = __dadd_rn(acc, __dmul_rn(df, df));
                }
            }
            edist[k] = sqrt(acc);
        }
        __syncthreads();
    }
#ifdef HRP_TIMING
    tq2 = clock64();
#endif
    // ---- stable sort of the N-1 edges by length (rank sort; equal keys keep discovery order) ---------------------------
    for (int k = tid; k < M; k += TR_THREADS) {
        const double dk = edist[k];
        int rank = 0;
        for (int j = 0; j < M; ++j) {
            double dj = edist[j];
            rank += (dj < dk) || (dj == dk && j < k);
        }
        sdist[rank] = dk; sa[rank] = ea[k]; sb[rank] = eb[k];
    }
    __syncthreads();
    // ---- union-find relabelling (scipy _hierarchy.label) + depth-first leaf order ------------------------------------
    if (tid == 0) {
        for (int i = 0; i < 2 * N - 1; ++i) { parent[i] = i; csize[i] = (i < N) ? 1 : 0; }
        for (int r = 0; r < M; ++r) {
            int a = sa[r], b = sb[r];
            int ra = a; while (parent[ra] != ra) ra = parent[ra];
            while (parent[a] != ra) { int nx = parent[a]; parent[a] = ra; a = nx; }
            int rb = b; while (parent[rb] != rb) rb = parent[rb];
            while (parent[b] != rb) { int nx = parent[b]; parent[b] = rb; b = nx; }
            int lo = min(ra, rb), hi = max(ra, rb);
            int nw = N + r;
            parent[ra] = nw; parent[rb] = nw;
            csize[nw] = csize[ra] + csize[rb];
            sa[r] = lo; sb[r] = hi;
        }
        // depth-first from the last merge, column-0 child first; `parent` is reused as the explicit stack
        int sp = 0, pos = 0;
        parent[sp++] = N + M - 1;
        while (sp > 0) {
            int node = parent[--sp];
            if (node < N) order[pos++] = node;
            else { parent[sp++] = sb[node - N]; parent[sp++] = sa[node - N]; }
        }
    }
    __syncthreads();
    if (link_all) {
        double* Z = link_all + (size_t)s * M * 4;
        for (int r = tid; r < M; r += TR_THREADS) {
            Z[r * 4 + 0] = sa[r]; Z[r * 4 + 1] = sb[r]; Z[r * 4 + 2] = sdist[r]; Z[r * 4 + 3] = csize[N + r];
        }
    }
    for (int i = tid; i < N; i += TR_THREADS) order_all[(size_t)s * N + i] = order[i];
#ifdef HRP_TIMING
    tq3 = clock64();
    if (tid == 0 && (s == 7 || s == 3000))
        printf("hrp_tree<%d> sim %d: prim %lld heights %lld sort+label+dfs %lld cycles\n", (int)FAST, s, tq1 - tq0, tq2 - tq1, tq3 - tq2);
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// P[a][b] = C[order[a]][order[b]]: the covariance in quasi-diagonal order, so every cluster of the bisection is a
// contiguous square block that can be read coalesced.  grid = (ceil(N/8), S), one warp per row; the gather stays inside
// one 8N-byte row of C (L1-resident after the first touch).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hrp_permute_kernel(const double* __restrict__ cov_all,
                                                          const int* __restrict__ order_all, int N,
                                                          double* __restrict__ P_all) {
    const int s = blockIdx.y;
    const int a = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (a >= N) return;
    const int* order = order_all + (size_t)s * N;
    const double* row = cov_all + (size_t)s * N * N + (size_t)order[a] * N;
    double* dst = P_all + (size_t)s * N * N + (size_t)a * N;
    for (int b = lane; b < N; b += 32) dst[b] = row[order[b]];
}

// ---------------------------------------------------------------------------------------------------------------
// Recursive bisection (optimizer.py:256-279), level by level, on the permuted covariance.  One CTA per simulation.
// Cluster variance of [lo, hi): w = ivp / sum(ivp), ivp_a = 1 / P[a][a];  var = w' P[lo:hi, lo:hi] w.
// ---------------------------------------------------------------------------------------------------------------
#ifndef BS_THREADS
#define BS_THREADS 256
#endif
#define BS_BIG 64   // clusters longer than this are reduced by the whole CTA, shorter ones by one warp each

__global__ void __launch_bounds__(BS_THREADS) hrp_bisect_kernel(const double* __restrict__ P_all, int N,
                                                                double* __restrict__ w_all, int* __restrict__ status) {
    extern __shared__ double sm[];
    double* u = sm;            // [N] inverse variances in position order
    double* wpos = u + N;      // [N]
    double* cvar = wpos + N;   // [N]
    double* red = cvar + N;    // [32]
    int* segA = reinterpret_cast<int*>(red + 32);  // [N+1]
    int* segB = segA + N + 1;                      // [N+1]
    __shared__ int s_nseg;
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int NW = BS_THREADS / 32;
    const double* P = P_all + (size_t)s * N * N;
    for (int i = tid; i < N; i += BS_THREADS) {
        u[i] = 1.0 / P[(size_t)i * N + i];
        wpos[i] = 1.0;
    }
    if (tid == 0) { segA[0] = 0; segA[1] = N; s_nseg = 1; }
    int* cur = segA;
    int* nxt = segB;
    while (true) {
        __syncthreads();
        const int nseg = s_nseg;
        __syncthreads();  // everyone has read s_nseg before thread 0 overwrites it below
        bool any_split = false;
        for (int q = 0; q < nseg; ++q) any_split |= (cur[q + 1] - cur[q] >= 2);   // uniform
        if (!any_split) break;
        if (tid == 0) {
            int nn = 0;
            for (int q = 0; q < nseg; ++q) {
                int lo = cur[q], hi = cur[q + 1], len = hi - lo;
                nxt[nn++] = lo;
                if (len >= 2) nxt[nn++] = lo + (len + 1) / 2;   // np.array_split: the first half gets the extra item
            }
            nxt[nn] = N;
            s_nseg = nn;
        }
        __syncthreads();
        const int nchild = s_nseg;
        // big clusters: the whole CTA, one after the other (only the first few levels have any)
        for (int ch = 0; ch < nchild; ++ch) {
            const int lo = nxt[ch], hi = nxt[ch + 1], len = hi - lo;
            if (len <= BS_BIG) continue;
            double isum = 0.0;
            for (int a = tid; a < len; a += BS_THREADS) isum += u[lo + a];
            isum = block_sum(isum, red);
            double q = 0.0;
            for (int a = wid; a < len; a += NW) {
                const double ua = u[lo + a];
                const double* row = P + (size_t)(lo + a) * N + lo;
                double part = 0.0;
                for (int b = lane; b < len; b += 32) part = fma(u[lo + b], row[b], part);
                q = fma(ua, part, q);
            }
            q = block_sum(q, red);
            if (tid == 0) cvar[ch] = q / (isum * isum);
        }
        // small clusters: one warp each
        for (int ch = wid; ch < nchild; ch += NW) {
            const int lo = nxt[ch], hi = nxt[ch + 1], len = hi - lo;
            if (len > BS_BIG) continue;
            double isum = 0.0;
            for (int a = lane; a < len; a += 32) isum += u[lo + a];
            isum = warp_sum(isum);
            double q = 0.0;
            for (int idx = lane; idx < len * len; idx += 32) {
                const int a = idx / len, b = idx - a * len;
                q = fma(u[lo + a] * u[lo + b], P[(size_t)(lo + a) * N + lo + b], q);
            }
            q = warp_sum(q);
            if (lane == 0) cvar[ch] = q / (isum * isum);
        }
        __syncthreads();
        for (int i = tid; i < N; i += BS_THREADS) {
            int lo = 0, hi = nchild;
            while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (nxt[mid] <= i) lo = mid; else hi = mid; }
            const int ch = lo;
            int pl = 0, ph = nseg;
            while (ph - pl > 1) { int mid = (pl + ph) >> 1; if (cur[mid] <= i) pl = mid; else ph = mid; }
            const int plo = cur[pl], phi = cur[pl + 1];
            if (phi - plo >= 2) {
                const bool left = (nxt[ch] == plo);
                const double lv = left ? cvar[ch] : cvar[ch - 1];
                const double rv = left ? cvar[ch + 1] : cvar[ch];
                const double alpha = 1.0 - lv / (lv + rv);
                wpos[i] *= left ? alpha : (1.0 - alpha);
            }
        }
        __syncthreads();
        int* t = cur; cur = nxt; nxt = t;
    }
    double part = 0.0;
    for (int i = tid; i < N; i += BS_THREADS) {
        double wv = wpos[i];
        w_all[(size_t)s * N + i] = wv;
        part += wv;
    }
    const double tot = block_sum(part, red);
    if (tid == 0 && status && !(tot <= 1.001 && tot >= 0.999)) atomicOr(&status[s], MCOSB_ST_HRP_SUM);
}

// Euclidean distances between the rows of D for S matrices (also used by NCO's silhouettes); the caller accounts for it.
// Euclidean distances between the rows of D through the Gram matrix on the FP64 tensor cores: E_ij = sqrt(max(0, n_i + n_j -
// 2 D_i . D_j)), rn = squared row norms.  What scikit-learn's euclidean_distances computes (same identity, same clipping at
// zero) -- used by NCO's k-means++ / silhouettes, where scipy's sequential pdist arithmetic is not the reference.
int mcosb_launch_gram_dist(mcosb_ctx* ctx, int S, const double* D, const double* rn, double* E) {
    const int N = ctx->N;
    const int ng = (N + DM_BM - 1) / DM_BM;
    const size_t smem_g = (size_t)DM_STAGES * GR_STAGE * sizeof(double);
    if (N % 2 == 0 && (uintptr_t)D % 16 == 0) {
        MCOSB_CUDA(ctx, cudaFuncSetAttribute(hrp_gram_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
        hrp_gram_kernel<true, true><<<dim3(ng * (ng + 1), S), GR_THREADS, smem_g, ctx->stream>>>(D, rn, N, E);
    } else {
        MCOSB_CUDA(ctx, cudaFuncSetAttribute(hrp_gram_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
        hrp_gram_kernel<false, true><<<dim3(ng * (ng + 1), S), GR_THREADS, smem_g, ctx->stream>>>(D, rn, N, E);
    }
    return MCOSB_OK;
}

void mcosb_launch_euclid(mcosb_ctx* ctx, int S, const double* D, double* E) {
    const int nt = (ctx->N + EU_T - 1) / EU_T;
    hrp_euclid_kernel<<<dim3(nt * (nt + 1) / 2, S), 256, 0, ctx->stream>>>(D, ctx->N, E, nullptr);
}

int mcosb_dev_hrp(mcosb_ctx* ctx, int S, const double* dcov, double* dw, int* dorder, double* dlink, int* dstatus) {
    if (S == 0) return MCOSB_OK;
    const int N = ctx->N;
    const size_t n = (size_t)N;
    const bool fast = ctx->hrp_mode != 1;
    double *D, *E, *rn = nullptr;
    int* slow = nullptr;
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_HRP_WORK, (size_t)S * n * n, &D));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_HRP_E, (size_t)S * n * n, &E));
    if (fast) {
        MCOSB_TRY(mcosb_scratch_t(ctx, SL_HRP_RN, (size_t)S * n, &rn));
        MCOSB_TRY(mcosb_scratch_t(ctx, SL_HRP_FLAG, (size_t)S, &slow));
    }
    hrp_dist_kernel<<<dim3((N + HD_ROWS - 1) / HD_ROWS, S), 256, n * sizeof(double), ctx->stream>>>(dcov, N, D, rn);
    MCOSB_LAUNCHED(ctx, MK_HRP_DIST);
    int* order = dorder;
    if (!order) MCOSB_TRY(mcosb_scratch_t(ctx, SL_HRP_ORDER, (size_t)S * n, &order));
    if (N > 32767) return mcosb_fail(ctx, MCOSB_ERR_ARG, "n_assets too large for hrp_tree_kernel (16-bit node ids)");
    size_t smem = (3 * n + 32) * sizeof(double) + 32 * sizeof(int) + 11 * n * sizeof(unsigned short);
    if (smem > (size_t)ctx->smem_optin) return mcosb_fail(ctx, MCOSB_ERR_ARG, "n_assets too large for hrp_tree_kernel");
    const int nt = (N + EU_T - 1) / EU_T;
    if (fast) {
        // approximate squared distances on the tensor cores, Prim with guarded decisions, exact merge heights
        const int ng = (N + DM_BM - 1) / DM_BM;
        const size_t smem_g = (size_t)DM_STAGES * GR_STAGE * sizeof(double);
        if (N % 2 == 0 && (uintptr_t)D % 16 == 0) {
            MCOSB_CUDA(ctx, cudaFuncSetAttribute(hrp_gram_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
            hrp_gram_kernel<true, false><<<dim3(ng * (ng + 1), S), GR_THREADS, smem_g, ctx->stream>>>(D, rn, N, E);
        } else {
            MCOSB_CUDA(ctx, cudaFuncSetAttribute(hrp_gram_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
            hrp_gram_kernel<false, false><<<dim3(ng * (ng + 1), S), GR_THREADS, smem_g, ctx->stream>>>(D, rn, N, E);
        }
        MCOSB_LAUNCHED(ctx, MK_HRP_GRAM);
        const int force = ctx->hrp_mode == 2 ? 1 : 0;
        if (N <= 1024 && ctx->hrp_mode != 3) {
            // Prim with one warp per simulation, then heights / sort / labelling with one CTA per simulation
            unsigned short* edges;
            MCOSB_TRY(mcosb_scratch_t(ctx, SL_HRP_EDGES, (size_t)S * 3 * n, &edges));
            const unsigned gp = (unsigned)((S + PR_WARPS - 1) / PR_WARPS);
#define PR_LAUNCH(M) hrp_prim_kernel<M><<<gp, PR_WARPS * 32, 0, ctx->stream>>>(E, rn, N, S, edges, slow, force, ctx->d_hrp_slow)
            if (N <= 128) PR_LAUNCH(4);
            else if (N <= 256) PR_LAUNCH(8);
            else if (N <= 512) PR_LAUNCH(16);
            else if (N <= 768) PR_LAUNCH(24);
            else PR_LAUNCH(32);
#undef PR_LAUNCH
            MCOSB_LAUNCHED(ctx, MK_HRP_PRIM);
            MCOSB_CUDA(ctx, cudaFuncSetAttribute(hrp_tree_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            hrp_tree_kernel<2><<<S, TR_THREADS, smem, ctx->stream>>>(E, dcov, N, dw, order, dlink, dstatus, D, rn, slow, nullptr,
                                                                    force, ctx->d_hrp_slow, edges);
        } else {
            MCOSB_CUDA(ctx, cudaFuncSetAttribute(hrp_tree_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            hrp_tree_kernel<1><<<S, TR_THREADS, smem, ctx->stream>>>(E, dcov, N, dw, order, dlink, dstatus, D, rn, slow, nullptr,
                                                                    force, ctx->d_hrp_slow, nullptr);
        }
        MCOSB_LAUNCHED(ctx, MK_HRP_TREE);
    }
    // exact distances (bit-identical to scipy's pdist): every simulation, or only those the fast pass flagged
    hrp_euclid_kernel<<<dim3(nt * (nt + 1) / 2, S), 256, 0, ctx->stream>>>(D, N, E, slow);
    MCOSB_LAUNCHED(ctx, MK_HRP_EUCLID);
    MCOSB_CUDA(ctx, cudaFuncSetAttribute(hrp_tree_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hrp_tree_kernel<0><<<S, TR_THREADS, smem, ctx->stream>>>(E, dcov, N, dw, order, dlink, dstatus, D, rn, nullptr, slow, 0,
                                                            nullptr, nullptr);
    MCOSB_LAUNCHED(ctx, MK_HRP_TREE);
    double* P = D;  // the distance matrix is dead once the trees exist: reuse its buffer for the permuted covariance
    hrp_permute_kernel<<<dim3((N + 7) / 8, S), 256, 0, ctx->stream>>>(dcov, order, N, P);
    MCOSB_LAUNCHED(ctx, MK_HRP_PERMUTE);
    size_t smem_b = (3 * n + 32) * sizeof(double) + (2 * (n + 1)) * sizeof(int);
    MCOSB_CUDA(ctx, cudaFuncSetAttribute(hrp_bisect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    hrp_bisect_kernel<<<S, BS_THREADS, smem_b, ctx->stream>>>(P, N, dw, dstatus);
    MCOSB_LAUNCHED(ctx, MK_HRP_BISECT);
    return MCOSB_OK;
}

// 0 = fast (tensor-core distances, guarded Prim, exact fallback for flagged simulations), 1 = exact distances for every
// simulation, 2 = fast pass with every simulation sent to the fallback (exercises the gating; results as mode 1),
// 3 = fast path with Prim inside the per-CTA tree kernel (what n_assets > 1024 uses), for any size (test aid)
extern "C" int mcosb_set_hrp_mode(mcosb_ctx* ctx, int mode) {
    MCOSB_CHECK_CTX(ctx);
    if (mode < 0 || mode > 3) return mcosb_fail(ctx, MCOSB_ERR_ARG, "hrp mode must be 0, 1, 2 or 3");
    ctx->hrp_mode = mode;
    return MCOSB_OK;
}

// Number of simulations the fast pass handed to the exact fallback since the last call (synchronises the stream).
extern "C" int mcosb_hrp_fallback_count(mcosb_ctx* ctx, int* count) {
    MCOSB_CHECK_CTX(ctx);
    if (!count) return mcosb_fail(ctx, MCOSB_ERR_ARG, "bad argument to mcosb_hrp_fallback_count");
    MCOSB_CUDA(ctx, cudaMemcpyAsync(count, ctx->d_hrp_slow, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MCOSB_CUDA(ctx, cudaMemsetAsync(ctx->d_hrp_slow, 0, sizeof(int), ctx->stream));
    MCOSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MCOSB_OK;
}

extern "C" int mcosb_hrp(mcosb_ctx* ctx, int S, const double* cov, double* w, int* order, double* linkage,
                         int* status) {
    MCOSB_CHECK_CTX(ctx);
    if (S < 0 || !cov || !w) return mcosb_fail(ctx, MCOSB_ERR_ARG, "bad argument to mcosb_hrp");
    const size_t N = ctx->N;
    Staging st(ctx);
    const double* dcov;
    double *dw, *dl;
    int *dord, *dst;
    MCOSB_TRY(st.in(cov, (size_t)S * N * N, SL_STAGE_IN0, &dcov));
    MCOSB_TRY(st.out(w, (size_t)S * N, SL_STAGE_OUT0, &dw));
    MCOSB_TRY(st.out(order, (size_t)S * N, SL_STAGE_OUT1, &dord));
    MCOSB_TRY(st.out(linkage, (size_t)S * (N > 1 ? N - 1 : 0) * 4, SL_STAGE_OUT2, &dl));
    MCOSB_TRY(st.out(status, (size_t)S, SL_STAGE_OUT3, &dst));
    if (dst) MCOSB_CUDA(ctx, cudaMemsetAsync(dst, 0, (size_t)S * sizeof(int), ctx->stream));
    MCOSB_TRY(mcosb_dev_hrp(ctx, S, dcov, dw, dord, dl, dst));
    return st.finish();
}
