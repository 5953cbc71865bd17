// Shared internals of libmcosb200: context, scratch arena, host/device pointer staging, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mcosb.h"

#define MCOSB_SMS 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

enum ScratchSlot {
    SL_STAGE_IN0 = 0, SL_STAGE_IN1, SL_STAGE_IN2, SL_STAGE_IN3,
    SL_STAGE_OUT0, SL_STAGE_OUT1, SL_STAGE_OUT2, SL_STAGE_OUT3, SL_STAGE_OUT4,
    SL_Z, SL_X, SL_MUHAT, SL_COV, SL_COLMEAN, SL_LWPART, SL_LWSTAT, SL_LWCOEF, SL_HRP_RN, SL_HRP_FLAG, SL_HRP_EDGES,
    SL_DN_A, SL_DN_SIGMA, SL_DN_D, SL_DN_E, SL_DN_TAU, SL_DN_LAM, SL_DN_VAR, SL_DN_NF, SL_DN_VEC, SL_DN_WORK,
    SL_MK_L, SL_MK_W, SL_MK_ITERS,
    SL_HRP_E, SL_HRP_W, SL_HRP_ORDER, SL_HRP_LINK, SL_HRP_WORK,
    SL_NCO_W, SL_NCO_LABELS, SL_NCO_WORK, SL_NCO_E,
    SL_ERR, SL_STATUS, SL_IDEAL, SL_RUN_W, SL_DN_Q2, SL_NCO_CENTERS, SL_S2_ZP, SL_S2_ZC, SL_S3_V, SL_S3_W, SL_S3_GT, SL_S3_ZP,
    SL_ZSUM_PART,
    SL_COUNT
};

enum KernelId {
    MK_CHOL_TRUTH,
    MK_ERRORS,
    MK_COLMEAN,
    MK_SYRK,
    MK_LW_ROWNORM,
    MK_LW_FINISH,
    MK_FILL,
    MK_PHILOX_NORMAL,
    MK_TRMM,
    MK_CORR,
    MK_TRIDIAG,
    MK_EIGVALS,
    MK_MPFIT,
    MK_EIGVEC,
    MK_MARKOWITZ,
    MK_HRP_DIST,
    MK_HRP_EUCLID,
    MK_HRP_TREE,
    MK_NCO_DIST,
    MK_NCO_CLUSTER,
    MK_NCO_ALLOC,
    MK_TRIDIAG2,
    MK_HRP_PERMUTE,
    MK_HRP_BISECT,
    MK_INVITER,
    MK_BACKTRANSFORM,
    MK_RECONSTRUCT,
    MK_DETONE,
    MK_RISK_PARITY,
    MK_SY2SB,
    MK_SB2ST,
    MK_Q2BACK,
    MK_HRP_GRAM,
    MK_HRP_PRIM,
    MK_SY2SB_PANEL,
    MK_SY2SB_PASS,
    MK_COUNT
};
static const char* const MCOSB_KERNEL_NAMES[MK_COUNT] = {
    "chol_truth",
    "errors",
    "colmean",
    "syrk",
    "lw_rownorm",
    "lw_finish",
    "fill",
    "philox_normal",
    "trmm",
    "corr",
    "tridiag",
    "eigvals",
    "mpfit",
    "eigvec",
    "markowitz",
    "hrp_dist",
    "hrp_euclid",
    "hrp_tree",
    "nco_dist",
    "nco_cluster",
    "nco_alloc",
    "tridiag2",
    "hrp_permute",
    "hrp_bisect",
    "inviter",
    "backtransform",
    "reconstruct",
    "detone",
    "risk_parity",
    "sy2sb",
    "sb2st",
    "q2back",
    "hrp_gram",
    "hrp_prim",
    "sy2sb_panel",
    "sy2sb_pass"
};

struct mcosb_ctx {
    int device = 0;
    int N = 0, T = 0;
    uint64_t seed = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    // the ideal allocations on the true (mu, cov) -- a handful of one-CTA kernels, ~2 ms -- run on this stream while the first
    // wave samples; the wave waits for ev_ideal before its first transformer / optimizer touches the scratch they share
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_ideal = nullptr, ev_main = nullptr;
    bool ideal_pending = false;
    double* d_mu = nullptr;    // true mu [N]
    double* d_cov = nullptr;   // true cov [N][N]
    double* d_chol = nullptr;  // lower Cholesky factor of the true cov [N][N]
    // k-slabs (16 columns of L) that are not entirely zero, per tile of 64 rows of L, written by mcosb_set_truth: the TRMM
    // walks this list instead of 0 .. diagonal (a block-diagonal truth has a block-diagonal factor).  Entry = slab index
    // | 0x4000 if rows 0..31 of the tile have a non-zero in the slab | 0x8000 for rows 32..63.
    unsigned short* d_klist = nullptr;   // [ceil(N / 64)][ceil(N / 16)]
    int* d_nk = nullptr;                 // [ceil(N / 64)] entries used, then [1] the number of zero slabs found below the diagonal
    bool chol_sparse = false;            // the list skips something: the TRMM reads it
    bool have_truth = false;
    void* slot_ptr[SL_COUNT] = {};
    size_t slot_bytes[SL_COUNT] = {};
    std::string err;
    long long launches = 0;
    double stage_ms[MCOSB_N_STAGES] = {};
    int smem_optin = 0;
    int num_sms = MCOSB_SMS;
    std::string ideal_key;  // optimizer list whose ideal allocations (on the current truth) sit in SL_IDEAL
    size_t mem_budget = 0;  // free HBM + context-held scratch at the first mcosb_run (wave sizing)
    int hrp_mode = 0;    // 0 fast MST + exact fallback, 1 exact distances only, 2 fast with every simulation sent to the fallback
    int* d_hrp_slow = nullptr;   // [1] simulations that took the exact fallback since the counter was last read
    int eig_method = 0;  // 0 auto, 1 one-stage Householder (tridiag / tridiag2), 2 two-stage (sy2sb + sb2st)
    bool s3_attr_set = false;   // shared-memory attributes of the sy2sb3 kernels set on this context's device
    // per-kernel timing (mcosb_set_profiling): one event after every launch; a kernel's time is the gap to the
    // previous event on the stream
    bool profile = false;
    std::vector<std::pair<int, cudaEvent_t>> kev;
    double kernel_ms[MK_COUNT] = {};
    long long kernel_launches[MK_COUNT] = {};
};

inline int mcosb_fail(mcosb_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

#define MCOSB_CUDA(ctx, call)                                                                        \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            char buf__[512];                                                                         \
            snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),   \
                     __FILE__, __LINE__);                                                            \
            return mcosb_fail(ctx, MCOSB_ERR_CUDA, buf__);                                           \
        }                                                                                            \
    } while (0)

#define MCOSB_TRY(expr)            \
    do {                           \
        int rc__ = (expr);         \
        if (rc__ != MCOSB_OK) return rc__; \
    } while (0)

// After every kernel launch: count it and surface launch-configuration errors immediately.
#define MCOSB_LAUNCHED(ctx, kid)                                        \
    do {                                                                \
        (ctx)->launches++;                                              \
        (ctx)->kernel_launches[kid]++;                                  \
        MCOSB_CUDA(ctx, cudaGetLastError());                            \
        if ((ctx)->profile) {                                           \
            cudaEvent_t ev__;                                            \
            MCOSB_CUDA(ctx, cudaEventCreate(&ev__));                     \
            MCOSB_CUDA(ctx, cudaEventRecord(ev__, (ctx)->stream));       \
            (ctx)->kev.push_back({kid, ev__});                           \
        }                                                               \
    } while (0)

// Marks a point on the stream that is not a kernel (start of a wave, after a copy) so the next kernel's gap is clean.
#define MCOSB_PROFILE_MARK(ctx)                                         \
    do {                                                                \
        if ((ctx)->profile) {                                           \
            cudaEvent_t ev__;                                            \
            MCOSB_CUDA(ctx, cudaEventCreate(&ev__));                     \
            MCOSB_CUDA(ctx, cudaEventRecord(ev__, (ctx)->stream));       \
            (ctx)->kev.push_back({-1, ev__});                            \
        }                                                               \
    } while (0)

// Grow-only device scratch, one buffer per slot.
inline int mcosb_scratch(mcosb_ctx* ctx, int slot, size_t bytes, void** out) {
    if (bytes == 0) bytes = 8;
    if (ctx->slot_bytes[slot] < bytes) {
        if (ctx->slot_ptr[slot]) {
            MCOSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            MCOSB_CUDA(ctx, cudaFree(ctx->slot_ptr[slot]));
            ctx->slot_ptr[slot] = nullptr;
            ctx->slot_bytes[slot] = 0;
        }
        size_t want = bytes + bytes / 8;
        MCOSB_CUDA(ctx, cudaMalloc(&ctx->slot_ptr[slot], want));
        ctx->slot_bytes[slot] = want;
    }
    *out = ctx->slot_ptr[slot];
    return MCOSB_OK;
}

template <typename T>
inline int mcosb_scratch_t(mcosb_ctx* ctx, int slot, size_t count, T** out) {
    void* p = nullptr;
    MCOSB_TRY(mcosb_scratch(ctx, slot, count * sizeof(T), &p));
    *out = static_cast<T*>(p);
    return MCOSB_OK;
}

inline bool mcosb_is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Tracks host buffers staged through device scratch within one API call.
struct Staging {
    mcosb_ctx* ctx;
    struct Out { void* host; void* dev; size_t bytes; };
    std::vector<Out> outs;
    bool any_host = false;
    explicit Staging(mcosb_ctx* c) : ctx(c) {}

    // Input: returns a device pointer holding `bytes` of `p` (p itself if it already is one). nullptr stays nullptr.
    template <typename T>
    int in(const T* p, size_t count, int slot, const T** out) {
        if (!p) { *out = nullptr; return MCOSB_OK; }
        if (mcosb_is_device_ptr(p)) { *out = p; return MCOSB_OK; }
        any_host = true;
        T* d = nullptr;
        MCOSB_TRY(mcosb_scratch_t(ctx, slot, count, &d));
        MCOSB_CUDA(ctx, cudaMemcpyAsync(d, p, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        *out = d;
        return MCOSB_OK;
    }
    // Output: device pointer to write into; copied back to the host buffer by finish().
    template <typename T>
    int out(T* p, size_t count, int slot, T** outp) {
        if (!p) { *outp = nullptr; return MCOSB_OK; }
        if (mcosb_is_device_ptr(p)) { *outp = p; return MCOSB_OK; }
        any_host = true;
        T* d = nullptr;
        MCOSB_TRY(mcosb_scratch_t(ctx, slot, count, &d));
        outs.push_back({p, d, count * sizeof(T)});
        *outp = d;
        return MCOSB_OK;
    }
    // In/out buffer.
    template <typename T>
    int inout(T* p, size_t count, int slot, T** outp) {
        if (!p) { *outp = nullptr; return MCOSB_OK; }
        if (mcosb_is_device_ptr(p)) { *outp = p; return MCOSB_OK; }
        any_host = true;
        T* d = nullptr;
        MCOSB_TRY(mcosb_scratch_t(ctx, slot, count, &d));
        MCOSB_CUDA(ctx, cudaMemcpyAsync(d, p, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        outs.push_back({p, d, count * sizeof(T)});
        *outp = d;
        return MCOSB_OK;
    }
    int finish() {
        for (auto& o : outs)
            MCOSB_CUDA(ctx, cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, ctx->stream));
        if (any_host) MCOSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return MCOSB_OK;
    }
};

#define MCOSB_CHECK_CTX(ctx)                          \
    do {                                              \
        if (!(ctx)) return MCOSB_ERR_ARG;             \
        cudaError_t e__ = cudaSetDevice((ctx)->device); \
        if (e__ != cudaSuccess) return mcosb_fail(ctx, MCOSB_ERR_CUDA, cudaGetErrorString(e__)); \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block-wide sum; `red` must hold >= 32 doubles of shared memory. All threads get the result.
__device__ __forceinline__ double block_sum(double v, double* red) {
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect red from a previous use
    if (lane == 0) red[wid] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    double t = (lane < nw) ? red[lane] : 0.0;
    t = warp_sum(t);
    return t;
}

// Ledoit-Wolf shrink of one covariance entry, (1 - sh) c (+ sh m on the diagonal).  Spelled with explicit roundings so
// that lw_finish_kernel (in place) and corr_kernel (on the fly, fused loop) produce the same bits.
__device__ __forceinline__ double lw_shrink(double c, double sh, double m, bool diag) {
    const double v = __dmul_rn(__dsub_rn(1.0, sh), c);
    return diag ? __fma_rn(sh, m, v) : v;
}

// ---------------------------------------------------------------------------------------------------------------
// internal (device-pointer) stage entry points used by the public wrappers and by mcosb_run
// ---------------------------------------------------------------------------------------------------------------
// add_mu = false: dX receives Y = Z L' = X - mu (the draws centred by the TRUE mean), for mcosb_dev_moments(..., shift = true mu)
// dybar != nullptr (with add_mu = false): the column means of Y are produced on the way -- the normal generator sums the columns
// of Z while it writes them, and zbar rides through the TRMM as one more row per simulation behind the last draw: ybar = zbar L'
// lands at dX + S T N (dX must have room for S (T + 1) N doubles), *dybar points there, mu + ybar goes to dmu_out[S][N]; *have_ybar
// tells whether that happened (N small enough for the generator's shared-memory sums), so that mcosb_dev_moments can skip its mean
// pass over X.
int mcosb_dev_sample(mcosb_ctx* ctx, uint64_t sim_id0, int S, double* dX, bool add_mu = true, double** dybar = nullptr,
                     double* dmu_out = nullptr, bool* have_ybar = nullptr);
// lwcoef_defer (Ledoit-Wolf only): instead of shrinking dcov in place, leave S_b there and write (shrinkage, mean
// variance) per simulation to lwcoef_defer[2 s .. 2 s + 1]; mcosb_dev_denoise applies them while it forms the correlation
// shift (device, [N]) != nullptr: dX holds the draws MINUS this vector (mcosb_dev_sample with add_mu = false); the means get
// it added back, and the covariance is formed as (Y'Y - T ybar ybar') / (T - 1) without centring in the GEMM loop -- exact
// algebra, and well conditioned because ybar is O(sigma / sqrt(T)).
// ybar_pre (with shift): the column means of dX are already known (mcosb_dev_sample) and dmu is already written.
int mcosb_dev_moments(mcosb_ctx* ctx, int S, const double* dX, int kind, double* dmu, double* dcov, double* dshrink,
                      double* lwcoef_defer = nullptr, const double* shift = nullptr, const double* ybar_pre = nullptr);
int mcosb_dev_denoise(mcosb_ctx* ctx, int S, double* dcov, int n_obs, double bandwidth, double* deig, double* dvar,
                      int* dnf, int* dstatus, const double* lwcoef = nullptr);
// *q2 receives the stage-2 reflectors when the two-stage tridiagonalisation was used (then A holds the panel reflectors
// of sy2sb_kernel, offset 8), nullptr for the one-stage kernels (offset 1)
int mcosb_dev_eig_prepare(mcosb_ctx* ctx, int S, const double* dcov, double* A, double* sigma, double* d, double* e,
                          double* tau, double* lam, const double** q2, const double* lwcoef = nullptr);
int mcosb_dev_detone(mcosb_ctx* ctx, int S, double* dcov, int n_remove, int* dstatus);
int mcosb_dev_risk_parity(mcosb_ctx* ctx, int S, const double* dcov, const double* dbudget, double* dw, int* diters,
                          int* dstatus);
int mcosb_dev_markowitz(mcosb_ctx* ctx, int S, const double* dmu, const double* dcov, double rf, int clean,
                        double* dw, int* diters, int* dstatus);
int mcosb_dev_hrp(mcosb_ctx* ctx, int S, const double* dcov, double* dw, int* dorder, double* dlink, int* dstatus);
int mcosb_dev_nco(mcosb_ctx* ctx, int S, const double* dmu, const double* dcov, int max_k, int trials,
                  const int* dlabels_in, double* dw, int* dlabels_out, int* dstatus);
// derr[s * ld + k] receives estimator kinds[k] (k < nk <= 3, host array) of simulation s
int mcosb_dev_errors(mcosb_ctx* ctx, int S, const double* dw, const double* dwstar, int batched, int nk,
                     const int* kinds, double* derr, int ld);
