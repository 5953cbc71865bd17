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
#include "common.cuh"

// ---------------------------------------------------------------------------------------------------------------
// grid = (ceil(N/32), S), 4 rows per warp; the square roots of the diagonal are taken once per CTA (shared memory).
// (1 - r) / 2 is an exact halving, so the multiplication by 0.5 rounds like numpy's division.
__global__ void __launch_bounds__(256) hrp_dist_kernel(const double* __restrict__ cov, int N, double* __restrict__ D) {
    extern __shared__ double sg[];   // [N]
    const int s = blockIdx.y;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double* c = cov + (size_t)s * N * N;
    double* d = D + (size_t)s * N * N;
    for (int j = threadIdx.x; j < N; j += 256) sg[j] = sqrt(c[(size_t)j * N + j]);
    __syncthreads();
    for (int q = 0; q < 4; ++q) {
        const int row = blockIdx.x * 32 + wid * 4 + q;
        if (row >= N) return;
        const double si = sg[row];
        for (int j = lane; j < N; j += 32) {
            double r = __ddiv_rn(c[(size_t)row * N + j], __dmul_rn(si, sg[j]));
            r = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
            double v = sqrt(__dmul_rn(__dsub_rn(1.0, r), 0.5));
            d[(size_t)row * N + j] = (j == row) ? 0.0 : v;
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
                                                         double* __restrict__ Eall) {
    __shared__ double Si[EU_K][EU_LD];
    __shared__ double Sj[EU_K][EU_LD];
    const int s = blockIdx.y;
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
#define TR_THREADS 256

__global__ void __launch_bounds__(TR_THREADS) hrp_tree_kernel(const double* __restrict__ Eall,
                                                              const double* __restrict__ cov_all, int N,
                                                              double* __restrict__ w_all, int* __restrict__ order_all,
                                                              double* __restrict__ link_all, int* __restrict__ status) {
    extern __shared__ double sm[];
    // doubles
    double* best = sm;              // [N]   Prim: distance to the tree;  later: position weights
    double* edist = best + N;       // [N]   edge lengths in discovery order
    double* sdist = edist + N;      // [N]   sorted edge lengths
    double* red = sdist + N;        // [32]
    // ints
    int* ea = reinterpret_cast<int*>(red + 32);  // [N] edge endpoint x (discovery order)
    int* eb = ea + N;               // [N] edge endpoint y
    int* sa = eb + N;               // [N] sorted endpoints / linkage column 0
    int* sb = sa + N;               // [N]                   linkage column 1
    int* parent = sb + N;           // [2N] union-find
    int* csize = parent + 2 * N;    // [2N] cluster sizes
    int* order = csize + 2 * N;     // [N]
    int* segA = order + N;          // [N+1]
    int* segB = segA + N + 1;       // [N+1]
    int* merged = segB + N + 1;     // [N]
    int* ired = merged + N;         // [32]
    __shared__ int s_x;

    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int NW = TR_THREADS / 32;
    const double* E = Eall + (size_t)s * N * N;

    if (N == 1) {
        if (tid == 0) order_all[s] = 0;
        return;
    }
    // ---- Prim's MST from node 0 (scipy _hierarchy.mst_single_linkage) -------------------------------------------------
    for (int i = tid; i < N; i += TR_THREADS) { best[i] = INFINITY; merged[i] = 0; }
    __syncthreads();
    int x = 0;
    for (int k = 0; k < N - 1; ++k) {
        if (tid == 0) merged[x] = 1;
        __syncthreads();
        const double* row = E + (size_t)x * N;
        double mv = INFINITY;
        int mi = N;
        for (int i = tid; i < N; i += TR_THREADS) {
            if (!merged[i]) {
                double dd = row[i];
                double b = best[i];
                if (b > dd) { b = dd; best[i] = b; }
                if (b < mv) { mv = b; mi = i; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, mv, o);
            int oi = __shfl_xor_sync(0xffffffffu, mi, o);
            if (ov < mv || (ov == mv && oi < mi)) { mv = ov; mi = oi; }
        }
        if (lane == 0) { red[wid] = mv; ired[wid] = mi; }
        __syncthreads();
        if (tid == 0) {
            double bv = red[0];
            int bi = ired[0];
            for (int q = 1; q < NW; ++q)
                if (red[q] < bv || (red[q] == bv && ired[q] < bi)) { bv = red[q]; bi = ired[q]; }
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
    // ---- stable sort of the N-1 edges by length (rank sort; equal keys keep discovery order) ---------------------------
    const int M = N - 1;
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
#define BS_THREADS 256
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
void mcosb_launch_euclid(mcosb_ctx* ctx, int S, const double* D, double* E) {
    const int nt = (ctx->N + EU_T - 1) / EU_T;
    hrp_euclid_kernel<<<dim3(nt * (nt + 1) / 2, S), 256, 0, ctx->stream>>>(D, ctx->N, E);
}

int mcosb_dev_hrp(mcosb_ctx* ctx, int S, const double* dcov, double* dw, int* dorder, double* dlink, int* dstatus) {
    if (S == 0) return MCOSB_OK;
    const int N = ctx->N;
    const size_t n = (size_t)N;
    double *D, *E;
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_HRP_WORK, (size_t)S * n * n, &D));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_HRP_E, (size_t)S * n * n, &E));
    hrp_dist_kernel<<<dim3((N + 31) / 32, S), 256, n * sizeof(double), ctx->stream>>>(dcov, N, D);
    MCOSB_LAUNCHED(ctx, MK_HRP_DIST);
    const int nt = (N + EU_T - 1) / EU_T;
    hrp_euclid_kernel<<<dim3(nt * (nt + 1) / 2, S), 256, 0, ctx->stream>>>(D, N, E);
    MCOSB_LAUNCHED(ctx, MK_HRP_EUCLID);
    int* order = dorder;
    if (!order) MCOSB_TRY(mcosb_scratch_t(ctx, SL_HRP_ORDER, (size_t)S * n, &order));
    size_t smem = (3 * n + 32) * sizeof(double) + (4 * n + 4 * n + n + 2 * (n + 1) + n + 32) * sizeof(int);
    if (smem > (size_t)ctx->smem_optin) return mcosb_fail(ctx, MCOSB_ERR_ARG, "n_assets too large for hrp_tree_kernel");
    MCOSB_CUDA(ctx, cudaFuncSetAttribute(hrp_tree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hrp_tree_kernel<<<S, TR_THREADS, smem, ctx->stream>>>(E, dcov, N, dw, order, dlink, dstatus);
    MCOSB_LAUNCHED(ctx, MK_HRP_TREE);
    double* P = D;  // the distance matrix is dead once E exists: reuse its buffer for the permuted covariance
    hrp_permute_kernel<<<dim3((N + 7) / 8, S), 256, 0, ctx->stream>>>(dcov, order, N, P);
    MCOSB_LAUNCHED(ctx, MK_HRP_PERMUTE);
    size_t smem_b = (3 * n + 32) * sizeof(double) + (2 * (n + 1)) * sizeof(int);
    MCOSB_CUDA(ctx, cudaFuncSetAttribute(hrp_bisect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    hrp_bisect_kernel<<<S, BS_THREADS, smem_b, ctx->stream>>>(P, N, dw, dstatus);
    MCOSB_LAUNCHED(ctx, MK_HRP_BISECT);
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
