// K1: Philox4x32-10 sampling of T x N multivariate-normal observations, X = mu + Z L' with L the once-computed
// Cholesky factor of the true covariance.  Replaces np.random.multivariate_normal (observation_simulator.py:27,39,51),
// which re-does an SVD of the true covariance on every call.
//
//   philox_normal_kernel : Z[s][t][2p..2p+1] = Box-Muller(Philox4x32-10(counter = (sim lo, sim hi, t, p), key = seed))
//                          -> pure RNG, one 128-bit Philox block per pair of normals; bound by the FP64 log/sincos
//                          pipe and by the 8TN bytes it writes.
//   trmm_kernel          : X = mu + Z L'  as one (S*T) x N x N FP64 tensor-core GEMM over all simulations at once
//                          (L is shared), skipping the k-slabs above the diagonal.  Flops 2 * T * N^2 / 2.
#include <cstdlib>

#include "common.cuh"
#include "dmma.cuh"

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Two standard normals from one Philox block: u1 in (0,1], u2 in [0,1) from 2 x 64 bits (top 53 used).
__device__ __forceinline__ void normal_pair(uint64_t sim, uint32_t t, uint32_t p, uint64_t seed, double& z0,
                                            double& z1) {
    uint32_t r[4];
    philox4x32_10((uint32_t)sim, (uint32_t)(sim >> 32), t, p, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    uint64_t a = ((uint64_t)r[1] << 32) | r[0];
    uint64_t b = ((uint64_t)r[3] << 32) | r[2];
    double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740992.0);  // (0, 1]
    double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);          // [0, 1)
    double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    z0 = rad * cs;
    z1 = rad * sn;
}

#ifndef PHILOX_CTAS_PER_SM
#define PHILOX_CTAS_PER_SM 16
#endif
// One warp per row (s, t) of Z[S][T][N], lanes over the pairs of that row, grid-stride over rows: the 64-bit divisions that
// turn a flat pair index into (s, t, p) -- about a third of the instructions when done per pair -- happen once per row.
__global__ void __launch_bounds__(256) philox_normal_kernel(double* __restrict__ Z, uint64_t sim_id0, int S, int T,
                                                            int N, uint64_t seed) {
    const uint32_t P = (uint32_t)(N + 1) / 2;  // pairs per row
    const size_t rows = (size_t)S * T;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool vec = (N & 1) == 0 && ((uintptr_t)Z & 15) == 0;   // rows 16-byte aligned: one 16-byte store per pair
    for (size_t row = (size_t)blockIdx.x * 8 + wid; row < rows; row += (size_t)gridDim.x * 8) {
        const uint64_t s = row / (uint32_t)T;
        const uint32_t t = (uint32_t)(row - s * (uint32_t)T);
        double* dst = Z + row * N;
        for (uint32_t p = lane; p < P; p += 32) {
            double z0, z1;
            normal_pair(sim_id0 + s, t, p, seed, z0, z1);
            if (vec) {
                reinterpret_cast<double2*>(dst)[p] = make_double2(z0, z1);
            } else {
                dst[2 * p] = z0;
                if (2 * p + 1 < (uint32_t)N) dst[2 * p + 1] = z1;
            }
        }
    }
}

// The same normals with their column sums: CTA (b, s) writes rows [b PRB, (b + 1) PRB) of simulation s, warp w the rows
// b PRB + w, + 8, ...; every warp adds what it writes into its own shared-memory row of N sums, the CTA folds the eight rows
// in warp order into part[s][b][N].  zbar_kernel adds the blocks in order and divides by T.  The column means of Y = Z L' are
// then zbar L' -- one more row per simulation for the TRMM instead of a pass over the 8 T N bytes of Y (colmean_kernel).
// Fixed summation order: bit-reproducible.
#define PRB 128
__global__ void __launch_bounds__(256) philox_normal_sums_kernel(double* __restrict__ Z, uint64_t sim_id0, int T, int N,
                                                                 uint64_t seed, double* __restrict__ part) {
    extern __shared__ double zacc[];               // [8][N2], N2 = 2 P
    const uint32_t P = (uint32_t)(N + 1) / 2;
    const int N2 = 2 * (int)P;
    const int s = blockIdx.y, b = blockIdx.x, nblk = gridDim.x;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int t0 = b * PRB, t1 = min(T, t0 + PRB);
    const bool vec = (N & 1) == 0 && ((uintptr_t)Z & 15) == 0;
    double* mine = zacc + (size_t)wid * N2;
    for (uint32_t p = lane; p < P; p += 32) mine[2 * p] = mine[2 * p + 1] = 0.0;      // the pairs this lane will add to
    for (int t = t0 + (int)wid; t < t1; t += 8) {
        double* dst = Z + ((size_t)s * T + t) * N;
        for (uint32_t p = lane; p < P; p += 32) {
            double z0, z1;
            normal_pair(sim_id0 + s, (uint32_t)t, p, seed, z0, z1);
            if (vec) {
                reinterpret_cast<double2*>(dst)[p] = make_double2(z0, z1);
            } else {
                dst[2 * p] = z0;
                if (2 * p + 1 < (uint32_t)N) dst[2 * p + 1] = z1;
            }
            double2 a = *reinterpret_cast<double2*>(mine + 2 * p);
            a.x += z0;
            a.y += z1;                             // column N of an odd N is never read back
            *reinterpret_cast<double2*>(mine + 2 * p) = a;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += 256) {
        double tot = zacc[i];
#pragma unroll
        for (int w = 1; w < 8; ++w) tot += zacc[(size_t)w * N2 + i];
        part[((size_t)s * nblk + b) * N + i] = tot;
    }
}

__global__ void __launch_bounds__(256) zbar_kernel(const double* __restrict__ part, int nblk, int T, int N,
                                                   double* __restrict__ zbar) {
    const int s = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    double tot = 0.0;
    for (int b = 0; b < nblk; ++b) tot += part[((size_t)s * nblk + b) * N + i];
    zbar[(size_t)s * N + i] = tot / (double)T;
}

// sample mean = true mean + ybar
__global__ void __launch_bounds__(256) add_shift_kernel(const double* __restrict__ ybar, const double* __restrict__ mu, int N,
                                                        size_t n, double* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) out[i] = mu[i % N] + ybar[i];
}

// X[m][j] = mu[j] + sum_{k <= j} Z[m][k] L[j][k];  M = S*T rows.  grid = (col tiles, row tiles)
// VEC: 16-byte cp.async (N even, Z and L 16-byte aligned), otherwise 8-byte copies.
template <bool VEC>
__global__ void __launch_bounds__(DM_THREADS, 1)
trmm_kernel(const double* __restrict__ Z, const double* __restrict__ L, const double* __restrict__ mu, size_t M, int N,
            double* __restrict__ X) {
    extern __shared__ double sm[];
    const int J = blockIdx.x;
    const size_t m0 = (size_t)blockIdx.y * DM_BM;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // column quarter of the warp: the two warps of a scheduler (wid and wid + 4) take quarters q and 3 - q, so that the
    // k-slabs a warp may skip inside the diagonal block of L (those entirely above its columns: 6, 4, 2, 0 of 8 for
    // quarters 0..3) take the same load off every scheduler
    const int cq = (wid >> 2) ? 3 - (wid & 3) : (wid & 3);
    const int wm0 = (wid >> 2) * DM_WM, wn0 = cq * DM_WN;
    // loader: VEC: k pair fixed, row = lr + 32 r; otherwise k fixed, row = lr + 16 r
    constexpr int LW = VEC ? 2 : 1, LRS = DM_THREADS * LW / DM_BK;
    const int lk = (tid % (DM_BK / LW)) * LW, lr = tid / (DM_BK / LW);

    double acc[DM_MB][DM_NB][2];
#pragma unroll
    for (int a = 0; a < DM_MB; ++a)
#pragma unroll
        for (int b = 0; b < DM_NB; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    auto issue = [&](int kt, int buf) {
#if defined(DM_ABLATE) && (DM_ABLATE & 4)   // timing experiment only: no global -> shared copies
        return;
#endif
        double* At = sm + (size_t)buf * 2 * DM_TILE_MMAJOR;
        double* Bt = At + DM_TILE_MMAJOR;
        const int k = kt * DM_BK + lk;
        const bool kok = k < N;
#pragma unroll
        for (int r = 0; r < DM_BM / LRS; ++r) {
            const int row = lr + LRS * r;
            const size_t m = m0 + row;
            const int j = J * DM_BN + row;
            const bool aok = kok && m < M, bok = kok && j < N;
            if (VEC) {
                cp_async16(&At[row * DM_LD_MMAJOR + lk], Z + (aok ? m * N + k : 0), aok);
                cp_async16(&Bt[row * DM_LD_MMAJOR + lk], L + (bok ? (size_t)j * N + k : 0), bok);
            } else {
                cp_async8(&At[row * DM_LD_MMAJOR + lk], Z + (aok ? m * N + k : 0), aok);
                cp_async8(&Bt[row * DM_LD_MMAJOR + lk], L + (bok ? (size_t)j * N + k : 0), bok);
            }
        }
    };
    const int kend = min(N, (J + 1) * DM_BN);  // L[j][k] = 0 for k > j
    const int nkt = (kend + DM_BK - 1) / DM_BK;
#pragma unroll
    for (int st = 0; st < DM_STAGES - 1; ++st) {
        if (st < nkt) issue(st, st);
        cp_async_commit();
    }
    for (int kt = 0; kt < nkt; ++kt) {
        cp_async_wait<DM_STAGES - 2>();
        __syncthreads();
        const double* At = sm + (size_t)(kt % DM_STAGES) * 2 * DM_TILE_MMAJOR;
        // the stage kt + 2 overwrites the buffer of slab kt - 1, which every warp has left (barrier above)
        auto next = [&] {
            if (kt + DM_STAGES - 1 < nkt) issue(kt + DM_STAGES - 1, (kt + DM_STAGES - 1) % DM_STAGES);
            cp_async_commit();
        };
        if (kt * DM_BK <= J * DM_BN + wn0 + DM_WN - 1)   // L[j][k] = 0 for k > j: nothing for this warp's columns
            dmma_slab<false, false, false>(At, At + DM_TILE_MMAJOR, wm0, wn0, lane, acc, next);
        else
            next();
    }
    const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
    for (int mb = 0; mb < DM_MB; ++mb)
#pragma unroll
        for (int nb = 0; nb < DM_NB; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                size_t row = m0 + wm0 + mb * 8 + g;
                int col = J * DM_BN + wn0 + nb * 8 + 2 * t4 + e;
                if (row < M && col < N) X[row * N + col] = acc[mb][nb][e] + mu[col];
            }
}

// The same product with 128 x 64 tiles and 4 warps (2 x 2, each 64 x 32), two CTAs per SM: while one CTA fills its
// pipeline or writes its tile the other keeps the tensor pipe busy (with one 256-thread CTA per SM about 10 % of the kernel
// was prologue / epilogue, most of it in the short tiles at the top of the triangle).
#define T2_THREADS 128
#define T2_BN 64
#define T2_STAGE ((DM_BM + T2_BN) * DM_LD_MMAJOR)
// ADDMU = false writes Y = Z L' = X - mu (the fused loop: its moments are taken from the pre-centred draws, mcosb_dev_moments
// with `shift`), which also takes the mu loads out of the epilogue.
// SPARSE: walk the list of non-zero k-slabs of L (mcosb_set_truth found whole slabs of zeros below the diagonal, e.g. a
// block-diagonal truth); otherwise every slab up to the diagonal, without the list reads.
template <bool VEC, bool ADDMU, bool SPARSE>
__global__ void __launch_bounds__(T2_THREADS, 2)
trmm2_kernel(const double* __restrict__ Z, const double* __restrict__ L, const double* __restrict__ mu, size_t M, int N,
             double* __restrict__ X, const unsigned short* __restrict__ klist, const int* __restrict__ nk) {
    extern __shared__ double sm[];
    const int J = blockIdx.x;
    const size_t m0 = (size_t)blockIdx.y * DM_BM;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int wm0 = (wid >> 1) * DM_WM, wn0 = (wid & 1) * DM_WN;
    constexpr int LW = VEC ? 2 : 1, LRS = T2_THREADS * LW / DM_BK;   // rows per loader pass
    const int lk = (tid % (DM_BK / LW)) * LW, lr = tid / (DM_BK / LW);

    double acc[DM_MB][DM_NB][2];
#pragma unroll
    for (int a = 0; a < DM_MB; ++a)
#pragma unroll
        for (int b = 0; b < DM_NB; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    auto issue = [&](int kt, int buf) {
        double* At = sm + (size_t)buf * T2_STAGE;
        double* Bt = At + DM_TILE_MMAJOR;
        const int k = kt * DM_BK + lk;
        const bool kok = k < N;
#pragma unroll
        for (int r = 0; r < DM_BM / LRS; ++r) {
            const int row = lr + LRS * r;
            const size_t m = m0 + row;
            const bool aok = kok && m < M;
            if (VEC) cp_async16(&At[row * DM_LD_MMAJOR + lk], Z + (aok ? m * N + k : 0), aok);
            else cp_async8(&At[row * DM_LD_MMAJOR + lk], Z + (aok ? m * N + k : 0), aok);
        }
#pragma unroll
        for (int r = 0; r < T2_BN / LRS; ++r) {
            const int row = lr + LRS * r;
            const int j = J * T2_BN + row;
            const bool bok = kok && j < N;
            if (VEC) cp_async16(&Bt[row * DM_LD_MMAJOR + lk], L + (bok ? (size_t)j * N + k : 0), bok);
            else cp_async8(&Bt[row * DM_LD_MMAJOR + lk], L + (bok ? (size_t)j * N + k : 0), bok);
        }
    };
    // SPARSE: the k-slabs of L with a non-zero in this tile's 64 rows, each with a bit per half of 32 rows = per warp
    // column set.  Dense: L[j][k] = 0 for k > j only.
    const unsigned short* kl = klist + (size_t)J * ((N + DM_BK - 1) / DM_BK);
    const int nkt = SPARSE ? nk[J] : (min(N, (J + 1) * T2_BN) + DM_BK - 1) / DM_BK;
    const unsigned mine = (wid & 1) ? 0x8000u : 0x4000u;
    auto slab = [&](int it) { return SPARSE ? (int)(kl[it] & 0x3fff) : it; };
#pragma unroll
    for (int st = 0; st < DM_STAGES - 1; ++st) {
        if (st < nkt) issue(slab(st), st);
        cp_async_commit();
    }
    for (int it = 0; it < nkt; ++it) {
        cp_async_wait<DM_STAGES - 2>();
        __syncthreads();
        const double* At = sm + (size_t)(it % DM_STAGES) * T2_STAGE;
        auto next = [&] {
            if (it + DM_STAGES - 1 < nkt) issue(slab(it + DM_STAGES - 1), (it + DM_STAGES - 1) % DM_STAGES);
            cp_async_commit();
        };
        // otherwise the slab holds only zeros for this warp's columns
        if (SPARSE ? (kl[it] & mine) != 0 : it * DM_BK <= J * T2_BN + wn0 + DM_WN - 1)
            dmma_slab<false, false, false>(At, At + DM_TILE_MMAJOR, wm0, wn0, lane, acc, next);
        else
            next();
    }
    const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
    for (int mb = 0; mb < DM_MB; ++mb)
#pragma unroll
        for (int nb = 0; nb < DM_NB; ++nb) {
            const size_t row = m0 + wm0 + mb * 8 + g;
            const int col = J * T2_BN + wn0 + nb * 8 + 2 * t4;
            if (row < M) {
                if (VEC && col + 1 < N) {
                    *reinterpret_cast<double2*>(X + row * N + col) =
                        ADDMU ? make_double2(acc[mb][nb][0] + mu[col], acc[mb][nb][1] + mu[col + 1])
                              : make_double2(acc[mb][nb][0], acc[mb][nb][1]);
                } else {
                    if (col < N) X[row * N + col] = ADDMU ? acc[mb][nb][0] + mu[col] : acc[mb][nb][0];
                    if (col + 1 < N) X[row * N + col + 1] = ADDMU ? acc[mb][nb][1] + mu[col + 1] : acc[mb][nb][1];
                }
            }
        }
}

// X = (mu +) Z L' for M rows of Z through trmm2_kernel (the 128 x 64 / 4-warp TRMM)
static int launch_trmm2(mcosb_ctx* ctx, const double* dZ, size_t M, double* dX, bool add_mu) {
    const int N = ctx->N;
    const size_t row_tiles = (M + DM_BM - 1) / DM_BM;
    if (row_tiles > 65535) return mcosb_fail(ctx, MCOSB_ERR_ARG, "sample batch too large; split it");
    const bool vec = N % 2 == 0 && (uintptr_t)dZ % 16 == 0 && (uintptr_t)ctx->d_chol % 16 == 0 && (uintptr_t)dX % 16 == 0;
    const size_t smem = (size_t)DM_STAGES * T2_STAGE * sizeof(double);
    dim3 g((N + T2_BN - 1) / T2_BN, (unsigned)row_tiles);
#define T2_LAUNCH(V, A, SP)                                                                                           \
    do {                                                                                                              \
        MCOSB_CUDA(ctx, cudaFuncSetAttribute(trmm2_kernel<V, A, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        trmm2_kernel<V, A, SP><<<g, T2_THREADS, smem, ctx->stream>>>(dZ, ctx->d_chol, ctx->d_mu, M, N, dX, ctx->d_klist, \
                                                                     ctx->d_nk);                                      \
    } while (0)
#define T2_LAUNCH2(V, A)                                                                                              \
    do {                                                                                                              \
        if (ctx->chol_sparse) T2_LAUNCH(V, A, true);                                                                  \
        else T2_LAUNCH(V, A, false);                                                                                  \
    } while (0)
    if (vec && add_mu) T2_LAUNCH2(true, true);
    else if (vec) T2_LAUNCH2(true, false);
    else if (add_mu) T2_LAUNCH2(false, true);
    else T2_LAUNCH2(false, false);
#undef T2_LAUNCH2
#undef T2_LAUNCH
    MCOSB_LAUNCHED(ctx, MK_TRMM);
    return MCOSB_OK;
}

int mcosb_dev_sample(mcosb_ctx* ctx, uint64_t sim_id0, int S, double* dX, bool add_mu, double** dybar, double* dmu_out,
                     bool* have_ybar) {
    if (have_ybar) *have_ybar = false;
    if (!ctx->have_truth) return mcosb_fail(ctx, MCOSB_ERR_STATE, "mcosb_set_truth has not been called");
    if (S == 0) return MCOSB_OK;
    const int N = ctx->N, T = ctx->T;
    const size_t M = (size_t)S * T;
    double* dZ = nullptr;
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_Z, (M + S) * N, &dZ));      // + one row of column means per simulation
    static const bool no_sums = std::getenv("MCOSB_NO_ZSUMS") != nullptr;   // A/B runs: mean pass over X instead
    const size_t smem_sums = (size_t)8 * 2 * ((N + 1) / 2) * sizeof(double);
    const int nblk = (T + PRB - 1) / PRB;
    const bool sums = dybar && dmu_out && !add_mu && !no_sums && S <= 65535 && smem_sums <= (size_t)ctx->smem_optin;
    double* part = nullptr;
    if (sums) {
        MCOSB_TRY(mcosb_scratch_t(ctx, SL_ZSUM_PART, (size_t)S * nblk * N, &part));
        MCOSB_CUDA(ctx, cudaFuncSetAttribute(philox_normal_sums_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sums));
        philox_normal_sums_kernel<<<dim3(nblk, S), 256, smem_sums, ctx->stream>>>(dZ, sim_id0, T, N, ctx->seed, part);
        MCOSB_LAUNCHED(ctx, MK_PHILOX_NORMAL);
        // zbar goes behind the last row of Z: the TRMM below turns it into ybar = zbar L' behind the last row of Y
        zbar_kernel<<<dim3((N + 255) / 256, S), 256, 0, ctx->stream>>>(part, nblk, T, N, dZ + M * N);
    } else {
        int grid = (int)std::min<size_t>((M + 7) / 8, (size_t)ctx->num_sms * PHILOX_CTAS_PER_SM);   // 8 warps = 8 rows per CTA and pass
        philox_normal_kernel<<<grid, 256, 0, ctx->stream>>>(dZ, sim_id0, S, T, N, ctx->seed);
    }
    MCOSB_LAUNCHED(ctx, MK_PHILOX_NORMAL);
    static const bool one_cta = std::getenv("MCOSB_TRMM_256") != nullptr;   // the 128 x 128 / 256-thread variant (A/B runs)
    if (one_cta && add_mu) {
        size_t row_tiles = (M + DM_BM - 1) / DM_BM;
        if (row_tiles > 65535) return mcosb_fail(ctx, MCOSB_ERR_ARG, "sample batch too large; split it");
        const bool vec = N % 2 == 0 && (uintptr_t)dZ % 16 == 0 && (uintptr_t)ctx->d_chol % 16 == 0 && (uintptr_t)dX % 16 == 0;
        const size_t smem = (size_t)DM_STAGES * 2 * DM_TILE_MMAJOR * sizeof(double);
        dim3 g((N + DM_BN - 1) / DM_BN, (unsigned)row_tiles);
        if (vec) {
            MCOSB_CUDA(ctx, cudaFuncSetAttribute(trmm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            trmm_kernel<true><<<g, DM_THREADS, smem, ctx->stream>>>(dZ, ctx->d_chol, ctx->d_mu, M, N, dX);
        } else {
            MCOSB_CUDA(ctx, cudaFuncSetAttribute(trmm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            trmm_kernel<false><<<g, DM_THREADS, smem, ctx->stream>>>(dZ, ctx->d_chol, ctx->d_mu, M, N, dX);
        }
        MCOSB_LAUNCHED(ctx, MK_TRMM);
    } else {
        MCOSB_TRY(launch_trmm2(ctx, dZ, sums ? M + S : M, dX, add_mu));
    }
    if (sums) {
        *dybar = dX + M * N;
        const size_t n = (size_t)S * N;
        add_shift_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(*dybar, ctx->d_mu, N, n, dmu_out);
        MCOSB_LAUNCHED(ctx, MK_COLMEAN);
        if (have_ybar) *have_ybar = true;
    }
    return MCOSB_OK;
}

extern "C" int mcosb_sample(mcosb_ctx* ctx, uint64_t sim_id0, int S, double* X) {
    MCOSB_CHECK_CTX(ctx);
    if (S < 0 || !X) return mcosb_fail(ctx, MCOSB_ERR_ARG, "bad argument to mcosb_sample");
    Staging st(ctx);
    double* dX;
    MCOSB_TRY(st.out(X, (size_t)S * ctx->T * ctx->N, SL_STAGE_OUT0, &dX));
    MCOSB_TRY(mcosb_dev_sample(ctx, sim_id0, S, dX, true));
    return st.finish();
}
