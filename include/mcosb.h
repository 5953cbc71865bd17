/* mcosb.h -- C ABI of the B200-native engine for the MCOS Monte-Carlo loop.
 *
 * The reference (enjine-com/mcos) is pure Python with no FFI of its own; the boundary this library sits behind is the
 * set of plugin methods the loop `mcos/mcos.py:22-33` calls.  Each entry point below is the batched form of one of
 * those methods and names the reference interface it replaces (paths relative to the reference repo).  The binding a
 * maintainer would add on the reference side (ctypes) is shown in INTEGRATION.md.
 *
 * Conventions
 *   - all matrices are row-major FP64, batch-major: X[S][T][N], cov[S][N][N], mu/w[S][N];
 *   - every data pointer may be a HOST pointer or a DEVICE pointer on the context's GPU; the library detects which
 *     (cudaPointerGetAttributes) and stages host buffers through context-owned device scratch.  With device pointers
 *     a call is asynchronous on the context's stream; with any host pointer it returns after the copies complete;
 *   - the caller owns every buffer it passes; the context owns scratch, freed by mcosb_destroy;
 *   - every function returns 0 (MCOSB_OK) or a negative code; mcosb_last_error() gives the text.  Per-simulation
 *     numerical conditions never abort a batch: they are OR-ed into the optional status[S] array;
 *   - a context is bound to one device and is not thread-safe (one context per GPU / per thread);
 *   - nullable output pointers are marked "nullable".
 */
#ifndef MCOSB_H
#define MCOSB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcosb_ctx mcosb_ctx;

/* return codes */
enum {
    MCOSB_OK = 0,
    MCOSB_ERR_ARG = -1,      /* bad argument */
    MCOSB_ERR_CUDA = -2,     /* CUDA runtime error (text in mcosb_last_error) */
    MCOSB_ERR_STATE = -3,    /* e.g. truth not set */
    MCOSB_ERR_NOT_PD = -4,   /* true covariance is not positive definite (Cholesky failed) */
    MCOSB_ERR_IDEAL = -5     /* mcosb_run: an optimizer failed on the TRUE (mu, cov) (mcos.py:30 would raise); the
                                MCOSB_ST_* bits are the last integer of mcosb_last_error.  MCOSB_ST_NO_EXCESS alone is
                                not a failure */
};

/* per-simulation status bits */
enum {
    MCOSB_ST_NOT_PD = 1,          /* a Cholesky pivot was <= 0 (Markowitz free-set system, NCO cluster solve) */
    MCOSB_ST_NO_EXCESS = 2,       /* Markowitz: no asset with mu_i > rf; the allocation is the single asset with the best
                                     (least negative) stand-alone Sharpe ratio -- informational, not a failure */
    MCOSB_ST_QP_CAP = 4,          /* Markowitz: active-set iteration cap reached */
    MCOSB_ST_MPFIT_FAIL = 8,      /* DeNoiser: line search failed -> var = 1 (covariance_transformer.py:127-130) */
    MCOSB_ST_EIGVEC = 16,         /* DeNoiser: an eigenvector residual check failed */
    MCOSB_ST_HRP_SUM = 32,        /* HRP: weights do not sum to 1 +- 1e-3 (optimizer.py:220-221 raises) */
    MCOSB_ST_NONFINITE = 64       /* NaN/Inf encountered in an input or result */
};

/* moment estimators: observation_simulator.py:31-40 / 19-28 / 43-57 */
enum { MCOSB_MOMENTS_MUCOV = 0, MCOSB_MOMENTS_LEDOIT_WOLF = 1, MCOSB_MOMENTS_JACKKNIFE = 2 };
/* error estimators: error_estimator.py:21-25 / 28-32 / 35-41 */
enum { MCOSB_ERR_EXPECTED_OUTCOME = 0, MCOSB_ERR_VARIANCE = 1, MCOSB_ERR_SHARPE = 2 };
/* optimizers for mcosb_run: optimizer.py:41-53 / 198-287 / 56-195 / 290-359 */
enum { MCOSB_OPT_MARKOWITZ = 0, MCOSB_OPT_HRP = 1, MCOSB_OPT_NCO = 2, MCOSB_OPT_RISK_PARITY = 3 };
/* covariance transformers for mcosb_run: covariance_transformer.py:55-208 / 211-250 */
enum { MCOSB_TR_DENOISE = 0, MCOSB_TR_DETONE = 1 };
#define MCOSB_MAX_TRANSFORMERS 4

int mcosb_version(void);

/* Creates a context on `device` for problems with n_assets variables and n_obs observations per simulation.
 * `seed` keys the Philox4x32-10 stream (counter = simulation id, row, column quad). */
int mcosb_create(mcosb_ctx** out, int device, int n_assets, int n_obs, uint64_t seed);
int mcosb_destroy(mcosb_ctx* ctx);
const char* mcosb_last_error(const mcosb_ctx* ctx);

/* Binds the context to a caller-owned cudaStream_t (e.g. torch's current stream); NULL = the context's own
 * (non-blocking) stream.  The CUDA default stream has handle 0 = NULL: to run on it pass cudaStreamLegacy ((void*)0x1)
 * or cudaStreamPerThread ((void*)0x2).  Calls with device pointers are asynchronous on the bound stream. */
int mcosb_set_stream(mcosb_ctx* ctx, void* cuda_stream);
/* Blocks until all work queued on the context's stream has finished. */
int mcosb_synchronize(mcosb_ctx* ctx);

/* Uploads the true (mu[N], cov[N][N]) -- the simulator attributes .mu/.cov (observation_simulator.py:21-24) -- and
 * factors cov = L L' once (replaces the SVD numpy redoes on every multivariate_normal call). */
int mcosb_set_truth(mcosb_ctx* ctx, const double* mu, const double* cov);

/* X[S][T][N] ~ N(mu, cov): replaces np.random.multivariate_normal(mu, cov, size=T) at
 * observation_simulator.py:27,39,51.  Simulation s uses Philox counter (sim_id0 + s, t, j/4), so results do not
 * depend on batch size or GPU count. */
int mcosb_sample(mcosb_ctx* ctx, uint64_t sim_id0, int S, double* X);

/* (mu_hat[S][N], cov_hat[S][N][N]) from injected draws X[S][T][N]: replaces x.mean(0) + np.cov / LedoitWolf / the
 * jackknife loops (observation_simulator.py:28, 40, 53-57).  shrink[S] (nullable) receives the Ledoit-Wolf shrinkage
 * (0 for the other kinds). */
int mcosb_moments(mcosb_ctx* ctx, int S, const double* X, int kind, double* mu_hat, double* cov_hat, double* shrink);

/* simulate() for simulations sim_id0 .. sim_id0 + S - 1 in one call (observation_simulator.py:26-28 / 38-40 / 50-57): the
 * draws never leave the device, and -- as inside mcosb_run -- they are kept centred by the TRUE mean (Y = X - mu), so the
 * sample mean is mu + mean(Y) and the covariance (Y'Y - T ybar ybar') / (T - 1) needs no centring pass.  Composing this
 * call with the stage entry points below reproduces mcosb_run bit for bit; mcosb_sample + mcosb_moments agree with it to
 * rounding (1e-13).  shrink[S] nullable. */
int mcosb_sample_moments(mcosb_ctx* ctx, uint64_t sim_id0, int S, int kind, double* mu_hat, double* cov_hat,
                         double* shrink);

/* DeNoiserCovarianceTransformer.transform (covariance_transformer.py:62-97), in place on cov[S][N][N].
 * Nullable outputs: eigvals[S][N] (descending), var[S] (fitted MP variance), n_facts[S], status[S]. */
int mcosb_denoise(mcosb_ctx* ctx, int S, double* cov_inout, int n_obs, double bandwidth, double* eigvals, double* var,
                  int* n_facts, int* status);

/* Selects the tridiagonalisation behind np.linalg.eigh (covariance_transformer.py:105) in mcosb_denoise / mcosb_detone /
 * mcosb_run: 0 = automatic (two-stage for 128 <= n_assets <= 696), 1 = one-stage blocked Householder, 2 = two-stage
 * (dense -> band on the FP64 tensor cores, band -> tridiagonal by bulge chasing in shared memory). */
int mcosb_set_eig_method(mcosb_ctx* ctx, int method);

/* Diagnostic: tridiagonal form (d[S][N], e[S][N] with e[i] = T[i+1][i], e[N-1] = 0) of symmetric matrices sym[S][N][N]
 * by the selected method -- the first half of np.linalg.eigh (covariance_transformer.py:105).  a_work[S][N][N]
 * (nullable) receives the work matrix holding the Householder reflectors. */
int mcosb_tridiagonalize(mcosb_ctx* ctx, int S, const double* sym, double* d, double* e, double* a_work);

/* DetoneCovarianceTransformer.transform (covariance_transformer.py:222-250), in place on cov[S][N][N]: removes the
 * n_remove (0..16) largest eigen-components of the correlation matrix, renormalises it to unit diagonal and rescales by
 * the input standard deviations.  n_remove == 0 leaves cov untouched.  status[S] nullable. */
int mcosb_detone(mcosb_ctx* ctx, int S, double* cov_inout, int n_remove, int* status);

/* MarkowitzOptimizer.allocate (optimizer.py:44-49 + PyPortfolioOpt 0.5.2 max_sharpe): long-only max-Sharpe weights
 * w[S][N] by an exact active-set QP.  clean != 0 applies clean_weights (|w| < 1e-4 -> 0, round to 5 decimals).
 * Nullable: iters[S], status[S]. */
int mcosb_markowitz(mcosb_ctx* ctx, int S, const double* mu, const double* cov, double rf, int clean, double* w,
                    int* iters, int* status);

/* HRPOptimizer.allocate (optimizer.py:204-223): w[S][N] in quasi-diagonal POSITION order, as the reference returns
 * it.  Nullable: order[S][N] (asset index at each position), linkage[S][N-1][4] (scipy layout), status[S]. */
int mcosb_hrp(mcosb_ctx* ctx, int S, const double* cov, double* w, int* order, double* linkage, int* status);

/* How mcosb_hrp / mcosb_run obtain scipy's single-linkage tree (sch.linkage on the square distance matrix,
 * optimizer.py:216).  0 (default): approximate squared row distances on the FP64 tensor cores, Prim's algorithm with
 * every comparison guarded by a rigorous error bound, merge heights recomputed with scipy's arithmetic; a simulation
 * with any comparison inside the bound is redone with exact distances, so the linkage is scipy's bit for bit either
 * way.  1: exact distances (scipy's pdist arithmetic) for every simulation.  2: as 0 with every simulation sent to the
 * exact fallback (test aid).  3: as 0 with the variant of the fast path used above 1024 assets (test aid). */
int mcosb_set_hrp_mode(mcosb_ctx* ctx, int mode);

/* Number of simulations that took the exact fallback since the last call (host int; synchronises the stream). */
int mcosb_hrp_fallback_count(mcosb_ctx* ctx, int* count);

/* NCOOptimizer.allocate (optimizer.py:74-138).  mu nullable (minimum-variance variant).  labels_in[S][N] nullable:
 * when given, the KMeans/silhouette search (optimizer.py:140-174) is skipped and this partition is used.
 * Nullable: labels_out[S][N], status[S]. */
int mcosb_nco(mcosb_ctx* ctx, int S, const double* mu, const double* cov, int max_clusters, int trials,
              const int* labels_in, double* w, int* labels_out, int* status);

/* RiskParityOptimizer.allocate (optimizer.py:298-359).  budget[N] nullable (equal risk budget 1/N).  Returns the
 * budget itself when SLSQP's first convergence test would fire at it (as the reference does on covariances of
 * realistic scale), otherwise the exact risk-budgeting weights (Newton on Spinu's convex formulation).
 * Nullable: iters[S] (0 = returned at the starting point), status[S]. */
int mcosb_risk_parity(mcosb_ctx* ctx, int S, const double* cov, const double* budget, double* w, int* iters,
                      int* status);

/* AbstractErrorEstimator.estimate against the TRUE (mu, cov) set by mcosb_set_truth (error_estimator.py:21-49).
 * w[S][N]; w_star is [S][N] if w_star_batched else [N] (the loop-invariant ideal allocation, mcos.py:30). */
int mcosb_errors(mcosb_ctx* ctx, int S, const double* w, const double* w_star, int w_star_batched, int kind,
                 double* err);
/* The same for several estimators in one pass over the weights: err[S][n_kinds], kinds[k] = MCOSB_ERR_*, 1 <= n_kinds
 * <= 3.  Column k is bit-identical to mcosb_errors(..., kinds[k], ...) (the three estimators share delta.mu and
 * delta' Cov delta, error_estimator.py:44-49). */
int mcosb_errors_multi(mcosb_ctx* ctx, int S, const double* w, const double* w_star, int w_star_batched, int n_kinds,
                       const int* kinds, double* err);

/* The whole loop body of mcos.py:22-33 for simulations sim_id0 .. sim_id0+S-1, stages chained on the device. */
typedef struct mcosb_plan {
    int moments_kind;        /* MCOSB_MOMENTS_* */
    int denoise;             /* 0/1: apply DeNoiserCovarianceTransformer */
    double bandwidth;        /* DeNoiser KDE bandwidth (0.25) */
    int n_optimizers;        /* <= 4 */
    int optimizers[4];       /* MCOSB_OPT_* in output-column order */
    int error_kind;          /* MCOSB_ERR_* */
    double risk_free_rate;   /* Markowitz (0.02) */
    int nco_max_clusters;    /* NCO: 0 = N/2 */
    int nco_trials;          /* NCO: num_clustering_trials */
    int wave;                /* simulations per wave (0 = choose from free memory) */
    int detone_n_remove;     /* > 0: apply DetoneCovarianceTransformer(n_remove) after the DeNoiser */
    /* ---- version >= 200 (zero-initialise for the behaviour above) ------------------------------------------------- */
    int n_error_kinds;       /* 0: one estimator, `error_kind`.  1..3: all of error_kinds[] from ONE pass over the
                                simulations (the estimators are an O(N^2) epilogue on the same weights) */
    int error_kinds[3];      /* MCOSB_ERR_* */
    int n_transformers;      /* 0: `denoise` / `detone_n_remove` above.  1..MCOSB_MAX_TRANSFORMERS: the list below,
                                applied in order (mcos.py:25-26 applies any list in order) */
    int transformer_kind[MCOSB_MAX_TRANSFORMERS];    /* MCOSB_TR_* */
    double transformer_arg[MCOSB_MAX_TRANSFORMERS];  /* DeNoiser: bandwidth; Detone: n_remove */
    const double* risk_budget;  /* RiskParityOptimizer(target_risk)[N] (optimizer.py:295-309), host or device pointer;
                                   NULL = equal budget 1/N */
} mcosb_plan;

/* err[S][n_optimizers] (err[S][n_optimizers][n_error_kinds] when plan->n_error_kinds > 0); status[S] nullable.  The
 * ideal allocations (optimizer.allocate(true mu, true cov)) are computed once per call instead of once per simulation. */
int mcosb_run(mcosb_ctx* ctx, const mcosb_plan* plan, uint64_t sim_id0, int S, double* err, int* status);

/* Per-stage device time (ms) of the last mcosb_run, accumulated over waves, measured with CUDA events on the
 * context's stream.  stage_ms must hold MCOSB_N_STAGES doubles. */
enum { MCOSB_STAGE_SAMPLE = 0, MCOSB_STAGE_MOMENTS = 1, MCOSB_STAGE_DENOISE = 2, MCOSB_STAGE_MARKOWITZ = 3,
       MCOSB_STAGE_HRP = 4, MCOSB_STAGE_NCO = 5, MCOSB_STAGE_ERRORS = 6, MCOSB_N_STAGES = 7 };
int mcosb_last_stage_ms(mcosb_ctx* ctx, double* stage_ms);
/* Number of kernel launches issued by this context since creation. */
long long mcosb_launch_count(const mcosb_ctx* ctx);

/* Per-kernel device time.  mcosb_set_profiling(ctx, 1) resets the counters and makes every launch record a CUDA
 * event on the context's stream; mcosb_kernel_times() synchronizes and fills ms[k] / launches[k] for
 * k < mcosb_kernel_count(), named by mcosb_kernel_name(k). */
int mcosb_set_profiling(mcosb_ctx* ctx, int on);
int mcosb_kernel_count(void);
const char* mcosb_kernel_name(int id);
int mcosb_kernel_times(mcosb_ctx* ctx, double* ms, long long* launches);

#ifdef __cplusplus
}
#endif
#endif /* MCOSB_H */
