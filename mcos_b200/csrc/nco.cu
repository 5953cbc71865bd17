// K4e-f: NCOOptimizer.allocate (optimizer.py:56-195), batched.
//
//   nco_dist_kernel     D = sqrt((1 - corr)/2)  (diagonal NOT zeroed, optimizer.py:148), rows of D are the points
//   hrp_euclid_kernel   (shared with HRP) Euclidean distances between rows of D, for the silhouettes
//   nco_cluster_kernel  one CTA per simulation: for k = 2..max_k  KMeans(k, n_init=1, random_state=42) -- k-means++
//                       seeding driven by the MT19937(42) uniform stream (data-independent, drawn on the host), Lloyd
//                       iterations with sklearn's stopping rules -- then silhouette_samples and the quality
//                       mean/std; keeps the best partition (optimizer.py:156-167).  Every clustering trial of the
//                       reference re-runs the identical fits (fixed seed) and can never win the strict `>` test, so
//                       one trial is evaluated.
//   nco_alloc_kernel    intra-cluster closed-form portfolios (Cholesky solve per cluster), reduced K x K problem,
//                       final weights (optimizer.py:119-138, 176-195).
//
// PARITY: the allocation algebra is pinned by the reference's goldens given a partition (tests/test_optimizer.py:30-47);
// the clustering follows scikit-learn 1.9's algorithm, not bit for bit (sklearn computes distances through a GEMM
// identity); agreement with sklearn's labels is measured in tests/test_gpu_nco.py.  "parity unpinned" vs sklearn 0.22.1.
#include <cstdlib>
#include <random>

#include "common.cuh"

#define NC_THREADS 256
#define NC_THREADS_MAX 512   // nco_cluster_kernel runs with 512 threads from N = 512 (one CTA per SM there anyway)
#define NCT ((int)blockDim.x)
#define NC_MAX_ITER 300
#define NC_CR 8            // rows of the trailing triangle a warp updates at once in nc_chol_solve
#define NC_MP 4            // (cluster, coordinate) pairs a thread sums at once in the M-step

__global__ void __launch_bounds__(256) nco_dist_kernel(const double* __restrict__ cov, int N, double* __restrict__ D,
                                                       double* __restrict__ rn) {
    extern __shared__ double sg[];   // [N] square roots of the diagonal, once per CTA (32 rows)
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
        double nrm = 0.0;
        // 8 independent loads per lane ahead of the divisions / square roots (as in hrp_dist_kernel); nrm keeps its order
        for (int j0 = lane; j0 < N; j0 += 256) {
            double x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = j0 + 32 * u;
                x[u] = (j < N) ? c[(size_t)row * N + j] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = j0 + 32 * u;
                if (j < N) {
                    double r = __ddiv_rn(x[u], __dmul_rn(si, sg[j]));
                    r = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
                    if (r != r) r = 0.0;  // corr.fillna(0)
                    const double v = sqrt(__dmul_rn(__dsub_rn(1.0, r), 0.5));
                    d[(size_t)row * N + j] = v;
                    nrm = fma(v, v, nrm);
                }
            }
        }
        nrm = warp_sum(nrm);           // squared row norm for the Gram form of the row distances
        if (lane == 0) rn[(size_t)s * N + row] = nrm;
    }
}

// declared in hrp.cu
void mcosb_launch_euclid(mcosb_ctx* ctx, int S, const double* D, double* E);
int mcosb_launch_gram_dist(mcosb_ctx* ctx, int S, const double* D, const double* rn, double* E);

struct NcSm {
    double* centers;   // [kmax][N]
    double* cnew;      // [kmax][N]   (aliased by the silhouette per-cluster distance sums [kmax][N])
    double* xnorm;     // [N]  (k-means++: distances to the best candidate so far)
    double* closest;   // [N]  k-means++ closest squared distance / generic per-point scratch
    double* cand;      // [NC_CG][N]  candidate distances (k-means++) / per-point scratch
    double* bestd;     // [N]  distances to the best candidate so far
    double* colmean;   // [N]
    double* cn2;       // [kmax] squared norms of centres
    double* shift;     // [kmax]
    double* red;       // [32]
    int* labels;       // [N]
    int* lold;         // [N]
    int* lbest;        // [N]
    int* counts;       // [kmax]
    int* ired;         // [32]
    int* picks;        // [NC_CG] candidate ids of the current k-means++ group
    int* members;      // [N] point ids grouped by cluster (ascending inside a cluster)
    int* cstart;       // [kmax + 1]
};

__device__ __forceinline__ double nc_block_sum(double v, NcSm& s) { return block_sum(v, s.red); }

// squared distance of every point to point `c`: the Euclidean distances between the rows of D are already there for the
// silhouettes (E, hrp_euclid_kernel), and centring does not change a difference, so ||x_i - x_c||^2 = E[c][i]^2 -- one
// coalesced row read instead of an N-term dot product per point (sklearn evaluates ||x||^2 - 2 x.y + ||y||^2, which agrees
// to rounding; the clustering is "not bit for bit" anyway, see the header).
__device__ void nc_dist_to_point(const double* __restrict__ E, int N, int c, double* out) {
    for (int i = threadIdx.x; i < N; i += NCT) {
        const double e = E[(size_t)c * N + i];
        out[i] = e * e;
    }
    __syncthreads();
}

// k-means++ draws its local trials all at once (candidate_ids = searchsorted(cumsum(closest), rand_vals)); they are
// handled NC_CG at a time: one parallel prefix scan finds the picks (the serial scan by one thread was 60 % of the kernel),
// one pass over the points gives the distances to all of them.
#define NC_CG 4

// picks[t] = smallest i with closest[0] + ... + closest[i] >= vals[t] (N - 1 if none), t < nv.
__device__ void nc_pick_multi(const double* __restrict__ wgt, int N, const double (&vals)[NC_CG], int nv, NcSm& s) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int chunk = (N + NCT - 1) / NCT;   // consecutive points per thread
    const int i0 = tid * chunk;
    double ls = 0.0;
    for (int e = 0; e < chunk; ++e)
        if (i0 + e < N) ls += wgt[i0 + e];
    double inc = ls;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double tt = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += tt;
    }
    __syncthreads();                       // red / picks free
    if (lane == 31) s.red[wid] = inc;
    if (tid < NC_CG) s.picks[tid] = N - 1;
    __syncthreads();
    double acc = inc - ls;
    for (int ww = 0; ww < wid; ++ww) acc += s.red[ww];
    bool done[NC_CG];
#pragma unroll
    for (int t = 0; t < NC_CG; ++t) done[t] = t >= nv;
    for (int e = 0; e < chunk; ++e) {
        if (i0 + e < N) {
            acc += wgt[i0 + e];
#pragma unroll
            for (int t = 0; t < NC_CG; ++t)
                if (!done[t] && acc >= vals[t]) { done[t] = true; atomicMin(&s.picks[t], i0 + e); }
        }
    }
    __syncthreads();
}

// out[t][i] = squared distance of point i to candidate s.picks[t], t < nv
__device__ void nc_dist_multi(const double* __restrict__ E, int N, int nv, NcSm& s, double* __restrict__ out) {
    for (int idx = threadIdx.x; idx < nv * N; idx += NCT) {
        const int t = idx / N, i = idx - t * N;
        const double e = E[(size_t)s.picks[t] * N + i];
        out[idx] = e * e;
    }
    __syncthreads();
}

// block sums of NC_CG values at once (all threads get all results)
__device__ void nc_block_sum_multi(double (&v)[NC_CG], NcSm& s) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < NC_CG; ++t) v[t] = warp_sum(v[t]);
    __syncthreads();
    if (lane < NC_CG) {
        double mine = v[0];
#pragma unroll
        for (int t = 1; t < NC_CG; ++t) mine = (lane == t) ? v[t] : mine;
        s.red[wid * NC_CG + lane] = mine;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < NC_CG; ++t) {
        double a = 0.0;
        for (int ww = 0; ww < NCT / 32; ++ww) a += s.red[ww * NC_CG + t];
        v[t] = a;
    }
}

// E-step: labels[i] = argmin_j (||c_j||^2 - 2 x_i.c_j), lowest j on ties.
// One thread per point and chunk of NC_JC centres: the point's column of X is read once per chunk and feeds NC_JC x 4
// accumulator chains (the (point, centre)-pair version re-read the column for every centre with 4 loads in flight per
// thread: 100 of 138 M cycles per simulation at N = 1000, all of it L2 latency).  Per pair the arithmetic is unchanged --
// four chains over d mod 4, (a0 + a1) + (a2 + a3), strict < over ascending j -- so the labels are bit for bit the old ones.
#define NC_JC 8
__device__ void nc_assign(const double* __restrict__ X, int N, int k, NcSm& s) {
    for (int j = threadIdx.x; j < k; j += NCT) {
        double q = 0.0;
        for (int d = 0; d < N; ++d) q = fma(s.centers[j * N + d], s.centers[j * N + d], q);
        s.cn2[j] = q;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += NCT) {
        double bestv = INFINITY;
        int bestj = 0;
        for (int j0 = 0; j0 < k; j0 += NC_JC) {
            const int nj = min(NC_JC, k - j0);
            const double* ca = s.centers + (size_t)j0 * N;
            double a[NC_JC][4];
#pragma unroll
            for (int jj = 0; jj < NC_JC; ++jj) a[jj][0] = a[jj][1] = a[jj][2] = a[jj][3] = 0.0;
            int d = 0;
            for (; d + 3 < N; d += 4) {
                double x[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) x[e] = X[(size_t)(d + e) * N + i] - s.colmean[d + e];
#pragma unroll
                for (int jj = 0; jj < NC_JC; ++jj) {
                    if (jj < nj) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) a[jj][e] = fma(x[e], ca[(size_t)jj * N + d + e], a[jj][e]);
                    }
                }
            }
            for (; d < N; ++d) {
                const double x0 = X[(size_t)d * N + i] - s.colmean[d];
#pragma unroll
                for (int jj = 0; jj < NC_JC; ++jj)
                    if (jj < nj) a[jj][0] = fma(x0, ca[(size_t)jj * N + d], a[jj][0]);
            }
#pragma unroll
            for (int jj = 0; jj < NC_JC; ++jj) {
                if (jj < nj) {
                    const double v = s.cn2[j0 + jj] - 2.0 * ((a[jj][0] + a[jj][1]) + (a[jj][2] + a[jj][3]));
                    if (v < bestv) { bestv = v; bestj = j0 + jj; }
                }
            }
        }
        s.labels[i] = bestj;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(NC_THREADS_MAX) nco_cluster_kernel(const double* __restrict__ Dall,
                                                                 const double* __restrict__ Eall, int N, int max_k,
                                                                 const double* __restrict__ uni,  // [max_k+1][ustride]
                                                                 int ustride, int* __restrict__ labels_out,
                                                                 double* __restrict__ quality_out,
                                                                 double* __restrict__ gwork) {
    extern __shared__ double sm[];
    NcSm s;
    // centres / new centres / silhouette sums [2][max_k][N]: in shared memory when they fit, otherwise in per-simulation
    // global scratch (the reference's default max_num_clusters = N / 2 needs 8 N^2 bytes: 2 MB at N = 500)
    s.centers = gwork ? gwork + (size_t)blockIdx.x * 2 * max_k * N : sm;
    s.cnew = s.centers + (size_t)max_k * N;
    s.xnorm = gwork ? sm : s.cnew + (size_t)max_k * N;
    s.closest = s.xnorm + N;
    s.cand = s.cnew;          // k-means++ candidate distances [ncg][N]: cnew is free until Lloyd starts
    s.bestd = s.xnorm;
    s.colmean = s.closest + N;
    s.cn2 = s.colmean + N;
    s.shift = s.cn2 + max_k;
    s.red = s.shift + max_k;
    s.labels = reinterpret_cast<int*>(s.red + 64);   // red: (warps <= 16) x NC_CG partial sums
    s.lold = s.labels + N;
    s.lbest = s.lold + N;
    s.counts = s.lbest + N;
    s.ired = s.counts + max_k;
    s.picks = s.ired + 32;
    s.members = s.picks + NC_CG;
    s.cstart = s.members + N;
    __shared__ int s_flag;
#ifdef NC_TIMING
    long long nacc[6] = {0, 0, 0, 0, 0, 0}, nlast = clock64();
    int niter = 0;
#define NC_TICK(i) do { if (threadIdx.x == 0) { long long n__ = clock64(); nacc[i] += n__ - nlast; nlast = n__; } } while (0)
#else
#define NC_TICK(i) do { } while (0)
#endif
    const int sim = blockIdx.x, tid = threadIdx.x;
    const double* X = Dall + (size_t)sim * N * N;   // symmetric: X[d][i] == X[i][d]; threads read "columns" coalesced
    const double* E = Eall + (size_t)sim * N * N;

    // centring (KMeans.fit subtracts the column means) and tolerance tol = 1e-4 * mean(var(X, axis=0))
    double vsum = 0.0;
    for (int d = tid; d < N; d += NCT) {
        double m = 0.0;
        for (int i = 0; i < N; ++i) m += X[(size_t)i * N + d];
        m /= N;
        double v = 0.0;
        for (int i = 0; i < N; ++i) { double t = X[(size_t)i * N + d] - m; v = fma(t, t, v); }
        s.colmean[d] = m;
        vsum += v / N;
    }
    const double tol = 1e-4 * nc_block_sum(vsum, s) / N;
    __syncthreads();

    double best_q = NAN;
    NC_TICK(0);
    for (int k = 2; k <= max_k; ++k) {
        const double* u = uni + (size_t)k * ustride;
        const int ntrials = 2 + (int)log((double)k);
        // ---- k-means++ ------------------------------------------------------------------------------------------------
        // first centre: RandomState.choice(n, p=1/n) -- data independent, the index comes with the uniform table
        const int first = (int)u[0];
        for (int d = tid; d < N; d += NCT) s.centers[d] = X[(size_t)d * N + first] - s.colmean[d];
        nc_dist_to_point(E, N, first, s.closest);
        double pot = 0.0;
        for (int i = tid; i < N; i += NCT) pot += s.closest[i];
        pot = nc_block_sum(pot, s);
        for (int c = 1; c < k; ++c) {
            double best_pot = INFINITY;
            int best_id = -1;
            const int ncg = min(NC_CG, max_k);
            for (int t0 = 0; t0 < ntrials; t0 += ncg) {
                const int nv = min(ncg, ntrials - t0);
                double vals[NC_CG];
#pragma unroll
                for (int t = 0; t < NC_CG; ++t) vals[t] = (t < nv) ? u[1 + (c - 1) * ntrials + t0 + t] * pot : 0.0;
                nc_pick_multi(s.closest, N, vals, nv, s);
                nc_dist_multi(E, N, nv, s, s.cand);
                double cp[NC_CG] = {0.0, 0.0, 0.0, 0.0};
                for (int i = tid; i < N; i += NCT) {
                    const double cl = s.closest[i];
#pragma unroll
                    for (int t = 0; t < NC_CG; ++t)
                        if (t < nv) cp[t] += fmin(cl, s.cand[t * N + i]);
                }
                nc_block_sum_multi(cp, s);
                int win = -1;
#pragma unroll
                for (int t = 0; t < NC_CG; ++t)
                    if (t < nv && cp[t] < best_pot) { best_pot = cp[t]; best_id = s.picks[t]; win = t; }  // first minimum
                if (win >= 0)
                    for (int i = tid; i < N; i += NCT) s.bestd[i] = s.cand[win * N + i];
                __syncthreads();
            }
            for (int i = tid; i < N; i += NCT) s.closest[i] = fmin(s.closest[i], s.bestd[i]);
            for (int d = tid; d < N; d += NCT) s.centers[c * N + d] = X[(size_t)d * N + best_id] - s.colmean[d];
            pot = best_pot;
            __syncthreads();
        }
        NC_TICK(1);
        // ---- Lloyd -----------------------------------------------------------------------------------------------------
        for (int i = tid; i < N; i += NCT) s.lold[i] = -1;
        __syncthreads();
        bool strict = false;
        for (int it = 0; it < NC_MAX_ITER; ++it) {
            nc_assign(X, N, k, s);
            NC_TICK(2);
#ifdef NC_TIMING
            ++niter;
#endif
            // M-step: new centres = mean of the members; empty clusters take the points farthest from their centres
            for (int j = tid; j < k; j += NCT) s.counts[j] = 0;
            __syncthreads();
            for (int i = tid; i < N; i += NCT) atomicAdd(&s.counts[s.labels[i]], 1);
            __syncthreads();
            // points grouped by cluster (stable: thread j collects cluster j in ascending order), so a centre coordinate is a
            // sum over its own members only instead of a predicated sum over all points
            if (tid == 0) {
                int acc = 0;
                for (int j = 0; j < k; ++j) { s.cstart[j] = acc; acc += s.counts[j]; }
                s.cstart[k] = acc;
            }
            __syncthreads();
            for (int j = tid; j < k; j += NCT) {
                int pos = s.cstart[j];
                for (int i = 0; i < N; ++i)
                    if (s.labels[i] == j) s.members[pos++] = i;
            }
            __syncthreads();
            // new centres: every (cluster, coordinate) pair is a sum over the cluster's members in four chains (e mod 4);
            // a thread runs NC_MP pairs at once so that 4 NC_MP loads are in flight instead of 4
            for (int idx0 = tid; idx0 < k * N; idx0 += NC_MP * NCT) {
                double c[NC_MP][4];
                int e[NC_MP], e1[NC_MP], dd[NC_MP];
                double cmd[NC_MP];
                bool live[NC_MP];
#pragma unroll
                for (int u = 0; u < NC_MP; ++u) {
                    const int idx = idx0 + u * NCT;
                    live[u] = idx < k * N;
                    const int j = live[u] ? idx / N : 0;
                    dd[u] = live[u] ? idx % N : 0;
                    cmd[u] = s.colmean[dd[u]];
                    e[u] = live[u] ? s.cstart[j] : 0;
                    e1[u] = live[u] ? s.cstart[j + 1] : 0;
                    c[u][0] = c[u][1] = c[u][2] = c[u][3] = 0.0;
                }
                bool more = true;
                while (more) {
                    more = false;
#pragma unroll
                    for (int u = 0; u < NC_MP; ++u) {
                        if (e[u] + 3 < e1[u]) {
                            c[u][0] += X[(size_t)s.members[e[u]] * N + dd[u]] - cmd[u];
                            c[u][1] += X[(size_t)s.members[e[u] + 1] * N + dd[u]] - cmd[u];
                            c[u][2] += X[(size_t)s.members[e[u] + 2] * N + dd[u]] - cmd[u];
                            c[u][3] += X[(size_t)s.members[e[u] + 3] * N + dd[u]] - cmd[u];
                            e[u] += 4;
                            more |= e[u] + 3 < e1[u];
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < NC_MP; ++u) {
                    for (; e[u] < e1[u]; ++e[u]) c[u][0] += X[(size_t)s.members[e[u]] * N + dd[u]] - cmd[u];
                    if (live[u]) s.cnew[idx0 + u * NCT] = (c[u][0] + c[u][1]) + (c[u][2] + c[u][3]);
                }
            }
            __syncthreads();
            if (tid == 0) {
                int n_empty = 0;
                for (int j = 0; j < k; ++j) n_empty += (s.counts[j] == 0);
                s_flag = n_empty;
            }
            __syncthreads();
            if (s_flag > 0) {
                // _relocate_empty_clusters_dense: distance of each point to its (old) centre, farthest points first
                for (int i = tid; i < N; i += NCT) {
                    double q = 0.0;
                    const int j = s.labels[i];
                    for (int d = 0; d < N; ++d) {
                        double t = X[(size_t)d * N + i] - s.colmean[d] - s.centers[j * N + d];
                        q = fma(t, t, q);
                    }
                    s.cand[i] = q;
                }
                __syncthreads();
                if (tid == 0) {
                    for (int j = 0; j < k; ++j) {
                        if (s.counts[j] != 0) continue;
                        int far = 0;
                        for (int i = 1; i < N; ++i)
                            if (s.cand[i] > s.cand[far]) far = i;
                        s.cand[far] = -1.0;
                        const int old = s.labels[far];
                        for (int d = 0; d < N; ++d) {
                            double x = X[(size_t)far * N + d] - s.colmean[d];
                            s.cnew[old * N + d] -= x;
                            s.cnew[j * N + d] = x;
                        }
                        s.counts[old] -= 1;
                        s.counts[j] = 1;
                    }
                }
                __syncthreads();
            }
            for (int idx = tid; idx < k * N; idx += NCT) {
                const int j = idx / N;
                if (s.counts[j] > 0) s.cnew[idx] /= s.counts[j];
            }
            __syncthreads();
            for (int j = tid; j < k; j += NCT) {
                double q = 0.0;
                for (int d = 0; d < N; ++d) { double t = s.cnew[j * N + d] - s.centers[j * N + d]; q = fma(t, t, q); }
                s.shift[j] = q;
            }
            __syncthreads();
            for (int idx = tid; idx < k * N; idx += NCT) s.centers[idx] = s.cnew[idx];
            int diff = 0;
            for (int i = tid; i < N; i += NCT) diff |= (s.labels[i] != s.lold[i]);
            const int changed = __syncthreads_or(diff);
            if (!changed) { strict = true; break; }
            double tot = 0.0;
            for (int j = 0; j < k; ++j) tot += s.shift[j];
            if (tot <= tol) break;
            for (int i = tid; i < N; i += NCT) s.lold[i] = s.labels[i];
            __syncthreads();
            NC_TICK(3);
        }
        NC_TICK(3);
        if (!strict) nc_assign(X, N, k, s);
        // ---- silhouette_samples (Euclidean between rows of D) and the quality mean/std --------------------------------
        double* dsum = s.cnew;  // [k][N] sum of distances from point i to the members of cluster j
        for (int j = tid; j < k; j += NCT) s.counts[j] = 0;
        __syncthreads();
        for (int i = tid; i < N; i += NCT) atomicAdd(&s.counts[s.labels[i]], 1);
        for (int idx = tid; idx < k * N; idx += NCT) dsum[idx] = 0.0;
        __syncthreads();
        for (int i = tid; i < N; i += NCT) {       // adds in ascending jx as before; the loads run 8 ahead of them
            int jx = 0;
            for (; jx + 7 < N; jx += 8) {
                double ev[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) ev[u] = E[(size_t)(jx + u) * N + i];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (jx + u != i) dsum[s.labels[jx + u] * N + i] += ev[u];
            }
            for (; jx < N; ++jx)
                if (jx != i) dsum[s.labels[jx] * N + i] += E[(size_t)jx * N + i];
        }
        __syncthreads();
        double ssum = 0.0;
        for (int i = tid; i < N; i += NCT) {
            const int li = s.labels[i];
            const int nli = s.counts[li];
            double sv = 0.0;
            if (nli > 1) {
                const double a = dsum[li * N + i] / (nli - 1);
                double b = INFINITY;
                for (int j = 0; j < k; ++j)
                    if (j != li && s.counts[j] > 0) b = fmin(b, dsum[j * N + i] / s.counts[j]);
                sv = (b - a) / fmax(a, b);
                if (sv != sv) sv = 0.0;
            }
            s.cand[i] = sv;
            ssum += sv;
        }
        const double mean = nc_block_sum(ssum, s) / N;
        double vs = 0.0;
        for (int i = tid; i < N; i += NCT) { double t = s.cand[i] - mean; vs = fma(t, t, vs); }
        const double sd = sqrt(nc_block_sum(vs, s) / N);
        const double q = mean / sd;
        if (best_q != best_q || q > best_q) {  // optimizer.py:166
            best_q = q;
            for (int i = tid; i < N; i += NCT) s.lbest[i] = s.labels[i];
        }
        __syncthreads();
        NC_TICK(4);
    }
#ifdef NC_TIMING
    if (tid == 0 && blockIdx.x == 0) printf("nco_cluster cycles: setup %lld kmeans++ %lld assign %lld mstep %lld silhouette %lld; lloyd iterations %d\n", nacc[0], nacc[1], nacc[2], nacc[3], nacc[4], niter);
#endif
    {
    for (int i = tid; i < N; i += NCT) labels_out[(size_t)sim * N + i] = s.lbest[i];
    if (tid == 0 && quality_out) quality_out[sim] = best_q;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Allocation given the partition. One CTA per simulation. Work arrays in global scratch: W [N] intra weights,
// member lists, Cholesky factor of each cluster block (<= N x N) and of the reduced matrix.
// ---------------------------------------------------------------------------------------------------------------
__device__ bool nc_chol_solve(double* __restrict__ A, int n, int lda, double* __restrict__ b, double* __restrict__ colk) {
    // in-place Cholesky of the SPD matrix A (lower), then solves A x = b (b overwritten). Block-cooperative.
    // Right-looking; column k is scaled and staged in shared memory (colk[n]), then every warp takes rows of the trailing
    // lower triangle with its lanes along the row: coalesced, no index divisions, nothing above the diagonal touched (a
    // 500 x 500 cluster took 70 ms with one thread per element of the trailing SQUARE and both factors read by column).
    // Per element the operations and their order are the old ones.
    bool ok = true;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int k = 0; k < n; ++k) {
        __syncthreads();
        const double d = A[(size_t)k * lda + k];
        if (!(d > 0.0)) ok = false;
        const double piv = ok ? sqrt(d) : 1.0;
        __syncthreads();
        for (int i = k + threadIdx.x; i < n; i += blockDim.x) {
            const double v = (i == k) ? piv : A[(size_t)i * lda + k] / piv;
            A[(size_t)i * lda + k] = v;
            colk[i] = v;
        }
        __syncthreads();
        // NC_CR rows per warp at a time: their loads are independent, so NC_CR of them are in flight per lane (one row at a
        // time exposed the full L2 latency per 32 elements: 31 ms for a 500 x 500 cluster)
        for (int i0 = k + 1 + wid; i0 < n; i0 += NC_CR * nw) {
            double aik[NC_CR];
            int ii[NC_CR];
#pragma unroll
            for (int u = 0; u < NC_CR; ++u) {
                ii[u] = i0 + u * nw;
                aik[u] = ii[u] < n ? colk[ii[u]] : 0.0;
                if (ii[u] >= n) ii[u] = -1;              // no column j satisfies j <= -1
            }
            const int ilast = min(i0 + (NC_CR - 1) * nw, n - 1);
            for (int j = k + 1 + lane; j <= ilast; j += 32) {
                const double cj = colk[j];
                double v[NC_CR];
#pragma unroll
                for (int u = 0; u < NC_CR; ++u)
                    if (j <= ii[u]) v[u] = A[(size_t)ii[u] * lda + j];
#pragma unroll
                for (int u = 0; u < NC_CR; ++u)
                    if (j <= ii[u]) A[(size_t)ii[u] * lda + j] = v[u] - aik[u] * cj;
            }
        }
    }
    // triangular solves, column-oriented and block-parallel (one thread did both, 2 n^2 / 2 dependent global loads: 30 of
    // 31 M cycles per simulation at N = 1000 with ten 100 x 100 clusters).  Forward: row i still subtracts its products in the
    // order k = 0 .. i - 1 with one FMA each, as the serial loop did; backward: in the order k = n - 1 .. i + 1.
    for (int k = 0; k < n; ++k) {            // L t = b
        __syncthreads();
        const double xk = b[k] / A[(size_t)k * lda + k];
        __syncthreads();
        if (threadIdx.x == 0) b[k] = xk;
        for (int i = k + 1 + threadIdx.x; i < n; i += blockDim.x) b[i] = fma(-A[(size_t)i * lda + k], xk, b[i]);
    }
    for (int k = n - 1; k >= 0; --k) {       // L' x = t
        __syncthreads();
        const double xk = b[k] / A[(size_t)k * lda + k];
        __syncthreads();
        if (threadIdx.x == 0) b[k] = xk;
        for (int i = threadIdx.x; i < k; i += blockDim.x) b[i] = fma(-A[(size_t)k * lda + i], xk, b[i]);
    }
    __syncthreads();
    return ok;
}

__global__ void __launch_bounds__(NC_THREADS) nco_alloc_kernel(const double* __restrict__ mu_all,
                                                               const double* __restrict__ cov_all,
                                                               const int* __restrict__ labels_all, int N,
                                                               double* __restrict__ work_all,  // [S][N*N + 4N]
                                                               double* __restrict__ w_all, int* __restrict__ status) {
    extern __shared__ double sm[];
    __shared__ double red[32];
    __shared__ int s_nc;
    double* colk = sm;                                // [N] current Cholesky column (nc_chol_solve)
    int* members = reinterpret_cast<int*>(sm + N);    // [N] asset ids grouped by cluster
    int* cstart = members + N;                        // [N+1]
    int* cid_of = cstart + N + 1;                     // [N] compact cluster index of each asset
    const int sim = blockIdx.x, tid = threadIdx.x;
    const double* C = cov_all + (size_t)sim * N * N;
    const double* mu = mu_all ? mu_all + (size_t)sim * N : nullptr;
    const int* labels = labels_all + (size_t)sim * N;
    double* work = work_all + (size_t)sim * ((size_t)N * N + 4 * (size_t)N);
    double* A = work;                   // cluster block / reduced matrix
    double* wi = work + (size_t)N * N;  // [N] intra-cluster weight of each asset
    double* rhs = wi + N;               // [N]
    double* cw = rhs + N;               // [N] C * (intra weights, per cluster) scratch
    double* inter = cw + N;             // [N]
    bool bad = false;
    // clusters in ascending label order (np.unique), members in ascending asset order
    if (tid == 0) {
        int nc = 0, pos = 0;
        int maxlab = 0;
        for (int i = 0; i < N; ++i) maxlab = max(maxlab, labels[i]);
        for (int l = 0; l <= maxlab; ++l) {
            int before = pos;
            for (int i = 0; i < N; ++i)
                if (labels[i] == l) { members[pos++] = i; cid_of[i] = nc; }
            if (pos > before) cstart[nc++] = before;
        }
        cstart[nc] = pos;
        s_nc = nc;
    }
    __syncthreads();
    const int K = s_nc;
    for (int c = 0; c < K; ++c) {
        const int lo = cstart[c], n = cstart[c + 1] - lo;
        for (int idx = tid; idx < n * n; idx += NC_THREADS) {
            int i = idx / n, j = idx % n;
            A[(size_t)i * n + j] = C[(size_t)members[lo + i] * N + members[lo + j]];
        }
        for (int i = tid; i < n; i += NC_THREADS) rhs[i] = mu ? mu[members[lo + i]] : 1.0;
        __syncthreads();
        if (!nc_chol_solve(A, n, n, rhs, colk)) bad = true;
        double part = 0.0;
        for (int i = tid; i < n; i += NC_THREADS) part += rhs[i];
        const double tot = block_sum(part, red);
        for (int i = tid; i < n; i += NC_THREADS) wi[members[lo + i]] = rhs[i] / tot;
        __syncthreads();
    }
    // reduced covariance W' C W (K x K) and mean W' mu
    for (int c = 0; c < K; ++c) {
        // cw = C * (w restricted to cluster c)
        const int lo = cstart[c], hi = cstart[c + 1];
        for (int i = tid; i < N; i += NC_THREADS) {
            double acc = 0.0;
            for (int p = lo; p < hi; ++p) acc = fma(C[(size_t)members[p] * N + i], wi[members[p]], acc);
            cw[i] = acc;
        }
        __syncthreads();
        for (int r = tid; r < K; r += NC_THREADS) {
            double acc = 0.0;
            for (int p = cstart[r]; p < cstart[r + 1]; ++p) acc = fma(wi[members[p]], cw[members[p]], acc);
            A[(size_t)r * K + c] = acc;
        }
        __syncthreads();
    }
    for (int r = tid; r < K; r += NC_THREADS) {
        double acc = 1.0;
        if (mu) {
            acc = 0.0;
            for (int p = cstart[r]; p < cstart[r + 1]; ++p) acc = fma(wi[members[p]], mu[members[p]], acc);
        }
        inter[r] = acc;
    }
    __syncthreads();
    if (!nc_chol_solve(A, K, K, inter, colk)) bad = true;
    double part = 0.0;
    for (int r = tid; r < K; r += NC_THREADS) part += inter[r];
    const double tot = block_sum(part, red);
    for (int i = tid; i < N; i += NC_THREADS) w_all[(size_t)sim * N + i] = wi[i] * (inter[cid_of[i]] / tot);
    if (tid == 0 && bad && status) atomicOr(&status[sim], MCOSB_ST_NOT_PD);
}

// MT19937(42) uniform stream exactly as numpy.random.RandomState(42).random_sample() produces it
static void nco_uniform_table(int max_k, int ustride, int n_points, std::vector<double>& out) {
    out.assign((size_t)(max_k + 1) * ustride, 0.0);
    for (int k = 2; k <= max_k; ++k) {
        std::mt19937 gen(42u);
        const int ntrials = 2 + (int)std::log((double)k);
        const int need = 1 + (k - 1) * ntrials;
        for (int i = 0; i < need; ++i) {
            uint32_t a = gen() >> 5, b = gen() >> 6;
            out[(size_t)k * ustride + i] = (a * 67108864.0 + b) / 9007199254740992.0;
        }
        // slot 0 becomes the first centre itself: RandomState.choice(n, p=1/n) is cdf = cumsum(p) / cdf[-1];
        // searchsorted(cdf, u, side='right') -- nothing in it depends on the data
        const double u0 = out[(size_t)k * ustride], p = 1.0 / n_points;
        double last = 0.0, acc = 0.0;
        for (int i = 0; i < n_points; ++i) last += p;
        int idx = n_points - 1;
        for (int i = 0; i < n_points; ++i) { acc += p; if (!(acc / last <= u0)) { idx = i; break; } }
        out[(size_t)k * ustride] = (double)idx;
    }
}

int mcosb_dev_nco(mcosb_ctx* ctx, int S, const double* dmu, const double* dcov, int max_k, int trials,
                  const int* dlabels_in, double* dw, int* dlabels_out, int* dstatus) {
    if (S == 0) return MCOSB_OK;
    const int N = ctx->N;
    const size_t n = (size_t)N;
    (void)trials;  // every trial repeats the identical fits (fixed random_state); see the header comment
    int* labels = dlabels_out;
    if (!labels) MCOSB_TRY(mcosb_scratch_t(ctx, SL_NCO_LABELS, (size_t)S * n, &labels));
    if (dlabels_in) {
        if (labels != dlabels_in)
            MCOSB_CUDA(ctx, cudaMemcpyAsync(labels, dlabels_in, (size_t)S * n * sizeof(int), cudaMemcpyDeviceToDevice,
                                            ctx->stream));
    } else {
        if (max_k <= 0) max_k = N / 2;
        if (max_k > N) max_k = N;
        if (max_k < 2) return mcosb_fail(ctx, MCOSB_ERR_ARG, "NCO needs at least 2 clusters to search over");
        double *D, *E;
        MCOSB_TRY(mcosb_scratch_t(ctx, SL_HRP_WORK, (size_t)S * n * n, &D));
        MCOSB_TRY(mcosb_scratch_t(ctx, SL_HRP_E, (size_t)S * n * n, &E));
        double* rn;
        MCOSB_TRY(mcosb_scratch_t(ctx, SL_HRP_RN, (size_t)S * n, &rn));
        nco_dist_kernel<<<dim3((N + 31) / 32, S), 256, n * sizeof(double), ctx->stream>>>(dcov, N, D, rn);
        MCOSB_LAUNCHED(ctx, MK_NCO_DIST);
        // Euclidean distances between the rows of D (k-means++ potentials, silhouettes).  scikit-learn forms them through
        // the Gram identity with clipping at zero; so does this (FP64 tensor cores: 33 instead of 103 us per simulation at
        // N = 1000 for scipy's sequential pdist arithmetic, which is HRP's reference, not NCO's).  MCOSB_NCO_EXACT_DIST=1
        // keeps the exact kernel (A/B of the partitions).
        static const bool exact_dist = getenv("MCOSB_NCO_EXACT_DIST") != nullptr;
        if (exact_dist) {
            mcosb_launch_euclid(ctx, S, D, E);
            MCOSB_LAUNCHED(ctx, MK_HRP_EUCLID);
        } else {
            MCOSB_TRY(mcosb_launch_gram_dist(ctx, S, D, rn, E));
            MCOSB_LAUNCHED(ctx, MK_HRP_GRAM);
        }
        const int ntr_max = 2 + (int)std::log((double)max_k);
        const int ustride = 1 + (max_k - 1) * ntr_max;
        std::vector<double> tab;
        nco_uniform_table(max_k, ustride, N, tab);
        double* duni;
        MCOSB_TRY(mcosb_scratch_t(ctx, SL_NCO_E, tab.size(), &duni));
        MCOSB_CUDA(ctx, cudaMemcpyAsync(duni, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        MCOSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `tab` is a pageable stack-owned buffer
        const size_t small = (3 * n + 2 * (size_t)max_k + 64) * sizeof(double) +
                             (4 * n + 2 * (size_t)max_k + 33 + NC_CG + 1) * sizeof(int);
        size_t smem = small + 2 * (size_t)max_k * n * sizeof(double);
        double* gwork = nullptr;
        if (smem > (size_t)ctx->smem_optin) {
            smem = small;
            if (smem > (size_t)ctx->smem_optin)
                return mcosb_fail(ctx, MCOSB_ERR_ARG, "n_assets too large for nco_cluster_kernel's shared memory");
            MCOSB_TRY(mcosb_scratch_t(ctx, SL_NCO_CENTERS, (size_t)S * 2 * max_k * n, &gwork));
        }
        MCOSB_CUDA(ctx, cudaFuncSetAttribute(nco_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        nco_cluster_kernel<<<S, N >= 512 ? NC_THREADS_MAX : (N <= 128 ? 128 : NC_THREADS), smem, ctx->stream>>>(D, E, N, max_k, duni, ustride, labels,
                                                                                         nullptr, gwork);
        MCOSB_LAUNCHED(ctx, MK_NCO_CLUSTER);
    }
    double* work;
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_NCO_WORK, (size_t)S * (n * n + 4 * n), &work));
    size_t smem2 = n * sizeof(double) + (3 * n + 1) * sizeof(int) + 16;
    MCOSB_CUDA(ctx, cudaFuncSetAttribute(nco_alloc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    nco_alloc_kernel<<<S, NC_THREADS, smem2, ctx->stream>>>(dmu, dcov, labels, N, work, dw, dstatus);
    MCOSB_LAUNCHED(ctx, MK_NCO_ALLOC);
    return MCOSB_OK;
}

extern "C" int mcosb_nco(mcosb_ctx* ctx, int S, const double* mu, const double* cov, int max_clusters, int trials,
                         const int* labels_in, double* w, int* labels_out, int* status) {
    MCOSB_CHECK_CTX(ctx);
    if (S < 0 || !cov || !w) return mcosb_fail(ctx, MCOSB_ERR_ARG, "bad argument to mcosb_nco");
    const size_t N = ctx->N;
    Staging st(ctx);
    const double *dmu, *dcov;
    const int* dlin;
    double* dw;
    int *dlout, *dst;
    MCOSB_TRY(st.in(mu, (size_t)S * N, SL_STAGE_IN0, &dmu));
    MCOSB_TRY(st.in(cov, (size_t)S * N * N, SL_STAGE_IN1, &dcov));
    MCOSB_TRY(st.in(labels_in, (size_t)S * N, SL_STAGE_IN2, &dlin));
    MCOSB_TRY(st.out(w, (size_t)S * N, SL_STAGE_OUT0, &dw));
    MCOSB_TRY(st.out(labels_out, (size_t)S * N, SL_STAGE_OUT1, &dlout));
    MCOSB_TRY(st.out(status, (size_t)S, SL_STAGE_OUT2, &dst));
    if (dst) MCOSB_CUDA(ctx, cudaMemsetAsync(dst, 0, (size_t)S * sizeof(int), ctx->stream));
    MCOSB_TRY(mcosb_dev_nco(ctx, S, dmu, dcov, max_clusters, trials, dlin, dw, dlout, dst));
    return st.finish();
}
