// mcosb_run: the body of the Monte-Carlo loop (mcos/mcos.py:22-33) for a range of simulation ids, all stages chained on
// the device, in waves sized to HBM.  The ideal allocation optimizer.allocate(true mu, true cov) (mcos.py:30) is
// loop-invariant and computed once per call instead of once per simulation.
#include <algorithm>
#include <chrono>
#include <cstdlib>

#include "common.cuh"
#include "sbr.cuh"

// X and Z of the sampling + moments stage are materialised per sub-chunk of this many bytes (each); larger chunks mean
// fewer, longer SYRK / TRMM launches (less tail), at 2 x this much HBM
#ifndef RUN_CHUNK_BYTES
#define RUN_CHUNK_BYTES ((size_t)4 << 30)
#endif

// The plan with its defaults resolved: transformer list (legacy fields or the explicit list), estimator kinds, device
// pointer of the risk budget.
struct RunPlan {
    int ntr = 0;
    int tr_kind[MCOSB_MAX_TRANSFORMERS] = {};
    double tr_arg[MCOSB_MAX_TRANSFORMERS] = {};
    int nk = 1;
    int kinds[3] = {0, 0, 0};
    const double* d_budget = nullptr;
};

static int resolve_plan(mcosb_ctx* ctx, const mcosb_plan* plan, RunPlan& rp) {
    if (plan->n_transformers < 0 || plan->n_transformers > MCOSB_MAX_TRANSFORMERS)
        return mcosb_fail(ctx, MCOSB_ERR_ARG, "n_transformers must be 0..4");
    if (plan->n_transformers > 0) {
        for (int i = 0; i < plan->n_transformers; ++i) {
            const int k = plan->transformer_kind[i];
            if (k != MCOSB_TR_DENOISE && k != MCOSB_TR_DETONE) return mcosb_fail(ctx, MCOSB_ERR_ARG, "unknown transformer kind");
            if (k == MCOSB_TR_DETONE && (int)plan->transformer_arg[i] == 0) continue;   // the identity
            rp.tr_kind[rp.ntr] = k;
            rp.tr_arg[rp.ntr++] = plan->transformer_arg[i];
        }
    } else {
        if (plan->denoise) { rp.tr_kind[rp.ntr] = MCOSB_TR_DENOISE; rp.tr_arg[rp.ntr++] = plan->bandwidth; }
        if (plan->detone_n_remove > 0) { rp.tr_kind[rp.ntr] = MCOSB_TR_DETONE; rp.tr_arg[rp.ntr++] = plan->detone_n_remove; }
    }
    if (plan->n_error_kinds < 0 || plan->n_error_kinds > 3) return mcosb_fail(ctx, MCOSB_ERR_ARG, "n_error_kinds must be 0..3");
    if (plan->n_error_kinds == 0) { rp.nk = 1; rp.kinds[0] = plan->error_kind; }
    else { rp.nk = plan->n_error_kinds; for (int k = 0; k < rp.nk; ++k) rp.kinds[k] = plan->error_kinds[k]; }
    for (int k = 0; k < rp.nk; ++k)
        if (rp.kinds[k] < 0 || rp.kinds[k] > 2) return mcosb_fail(ctx, MCOSB_ERR_ARG, "unknown error estimator kind");
    return MCOSB_OK;
}

int mcosb_run_wave(mcosb_ctx* ctx, const mcosb_plan* plan, const RunPlan& rp, uint64_t sim0, int W, const double* d_ideal,
                   double* derr, int* dstatus, std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>>& marks) {
    const size_t N = ctx->N, T = ctx->T;
    double *dmu, *dcov;
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_MUHAT, (size_t)W * N, &dmu));
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_COV, (size_t)W * N * N, &dcov));
    auto mark = [&](int stage, bool begin) -> int {
        cudaEvent_t e;
        MCOSB_CUDA(ctx, cudaEventCreate(&e));
        MCOSB_CUDA(ctx, cudaEventRecord(e, ctx->stream));
        if (begin) marks.push_back({stage, {e, nullptr}});
        else marks.back().second.second = e;
        return MCOSB_OK;
    };
    MCOSB_PROFILE_MARK(ctx);
    // ---- sampling + moments in sub-chunks so X (8TN bytes per simulation) stays small ---------------------------------
    const size_t chunk_bytes = RUN_CHUNK_BYTES;
    int CH = (int)std::max<size_t>(1, std::min<size_t>((size_t)W, chunk_bytes / (T * N * sizeof(double))));
    CH = (W + (W + CH - 1) / CH - 1) / ((W + CH - 1) / CH);   // the same number of sub-chunks, equal sizes
    double *dX, *dyb = nullptr;
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_X, (size_t)CH * (T + 1) * N, &dX));   // + one row per simulation: the column means
    // Ledoit-Wolf followed by the de-noiser: the shrink (one read-modify-write of every covariance matrix) is folded
    // into the de-noiser's covariance -> correlation pass, which is the only consumer of the shrunk matrix
    double* lwcoef = nullptr;
    if (plan->moments_kind == MCOSB_MOMENTS_LEDOIT_WOLF && rp.ntr > 0 && rp.tr_kind[0] == MCOSB_TR_DENOISE)
        MCOSB_TRY(mcosb_scratch_t(ctx, SL_LWCOEF, 2 * (size_t)W, &lwcoef));
    for (int c0 = 0; c0 < W; c0 += CH) {
        const int nc = std::min(CH, W - c0);
        MCOSB_TRY(mark(MCOSB_STAGE_SAMPLE, true));
        // the draws stay centred by the TRUE mean (Y = Z L'): the moment kernels add it back to the sample mean and take the
        // covariance of Y without centring inside the GEMM loop (mcosb_sample_moments is the same pair as a stage call)
        // ... and their column means come out of the sampler (sums of Z while it is written, then zbar L'): no mean pass
        bool have_ybar = false;
        MCOSB_TRY(mcosb_dev_sample(ctx, sim0 + c0, nc, dX, false, &dyb, dmu + (size_t)c0 * N, &have_ybar));
        MCOSB_TRY(mark(MCOSB_STAGE_SAMPLE, false));
        MCOSB_TRY(mark(MCOSB_STAGE_MOMENTS, true));
        MCOSB_TRY(mcosb_dev_moments(ctx, nc, dX, plan->moments_kind, dmu + (size_t)c0 * N, dcov + (size_t)c0 * N * N,
                                    nullptr, lwcoef ? lwcoef + 2 * (size_t)c0 : nullptr, ctx->d_mu, have_ybar ? dyb : nullptr));
        MCOSB_TRY(mark(MCOSB_STAGE_MOMENTS, false));
    }
    if (ctx->ideal_pending) {            // the ideal allocations (auxiliary stream) share scratch with the stages below
        MCOSB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_ideal, 0));
        ctx->ideal_pending = false;
    }
    for (int i = 0; i < rp.ntr; ++i) {   // any list, in order (mcos.py:25-26)
        MCOSB_TRY(mark(MCOSB_STAGE_DENOISE, true));
        if (rp.tr_kind[i] == MCOSB_TR_DENOISE)
            MCOSB_TRY(mcosb_dev_denoise(ctx, W, dcov, ctx->T, rp.tr_arg[i], nullptr, nullptr, nullptr, dstatus,
                                        i == 0 ? lwcoef : nullptr));
        else
            MCOSB_TRY(mcosb_dev_detone(ctx, W, dcov, (int)rp.tr_arg[i], dstatus));
        MCOSB_TRY(mark(MCOSB_STAGE_DENOISE, false));
    }
    double* dw;
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_RUN_W, (size_t)W * N, &dw));
    for (int o = 0; o < plan->n_optimizers; ++o) {
        const int kind = plan->optimizers[o];
        int stage;
        if (kind == MCOSB_OPT_MARKOWITZ) {
            stage = MCOSB_STAGE_MARKOWITZ;
            MCOSB_TRY(mark(stage, true));
            MCOSB_TRY(mcosb_dev_markowitz(ctx, W, dmu, dcov, plan->risk_free_rate, 1, dw, nullptr, dstatus));
        } else if (kind == MCOSB_OPT_HRP) {
            stage = MCOSB_STAGE_HRP;
            MCOSB_TRY(mark(stage, true));
            MCOSB_TRY(mcosb_dev_hrp(ctx, W, dcov, dw, nullptr, nullptr, dstatus));
        } else if (kind == MCOSB_OPT_RISK_PARITY) {
            stage = MCOSB_STAGE_NCO;   // reported with the "other optimizers" stage timer
            MCOSB_TRY(mark(stage, true));
            MCOSB_TRY(mcosb_dev_risk_parity(ctx, W, dcov, rp.d_budget, dw, nullptr, dstatus));
        } else {
            stage = MCOSB_STAGE_NCO;
            MCOSB_TRY(mark(stage, true));
            MCOSB_TRY(mcosb_dev_nco(ctx, W, dmu, dcov, plan->nco_max_clusters, plan->nco_trials, nullptr, dw, nullptr,
                                    dstatus));
        }
        MCOSB_TRY(mark(stage, false));
        MCOSB_TRY(mark(MCOSB_STAGE_ERRORS, true));
        MCOSB_TRY(mcosb_dev_errors(ctx, W, dw, d_ideal + (size_t)o * N, 0, rp.nk, rp.kinds, derr + (size_t)o * rp.nk,
                                   plan->n_optimizers * rp.nk));
        MCOSB_TRY(mark(MCOSB_STAGE_ERRORS, false));
    }
    return MCOSB_OK;
}

// Sampling + moment estimation exactly as mcosb_run chains them (draws centred by the true mean, sub-chunks of
// RUN_CHUNK_BYTES), as a stage call: what a caller composes with mcosb_denoise / mcosb_markowitz / ... to reproduce
// mcosb_run bit for bit.
extern "C" int mcosb_sample_moments(mcosb_ctx* ctx, uint64_t sim_id0, int S, int kind, double* mu_hat, double* cov_hat,
                                    double* shrink) {
    MCOSB_CHECK_CTX(ctx);
    if (S < 0 || !mu_hat || !cov_hat) return mcosb_fail(ctx, MCOSB_ERR_ARG, "bad argument to mcosb_sample_moments");
    if (!ctx->have_truth) return mcosb_fail(ctx, MCOSB_ERR_STATE, "mcosb_set_truth has not been called");
    const size_t N = ctx->N, T = ctx->T;
    Staging st(ctx);
    double *dmu, *dcov, *dsh;
    MCOSB_TRY(st.out(mu_hat, (size_t)S * N, SL_STAGE_OUT0, &dmu));
    MCOSB_TRY(st.out(cov_hat, (size_t)S * N * N, SL_STAGE_OUT1, &dcov));
    MCOSB_TRY(st.out(shrink, (size_t)S, SL_STAGE_OUT2, &dsh));
    if (S > 0) {
        const int CH = (int)std::max<size_t>(1, std::min<size_t>((size_t)S, (size_t)RUN_CHUNK_BYTES / (T * N * sizeof(double))));
        double *dX, *dyb = nullptr;
        MCOSB_TRY(mcosb_scratch_t(ctx, SL_X, (size_t)CH * (T + 1) * N, &dX));
        for (int c0 = 0; c0 < S; c0 += CH) {
            const int nc = std::min(CH, S - c0);
            bool have_ybar = false;
            MCOSB_TRY(mcosb_dev_sample(ctx, sim_id0 + c0, nc, dX, false, &dyb, dmu + (size_t)c0 * N, &have_ybar));
            MCOSB_TRY(mcosb_dev_moments(ctx, nc, dX, kind, dmu + (size_t)c0 * N, dcov + (size_t)c0 * N * N,
                                        dsh ? dsh + c0 : nullptr, nullptr, ctx->d_mu, have_ybar ? dyb : nullptr));
        }
    }
    return st.finish();
}

extern "C" int mcosb_run(mcosb_ctx* ctx, const mcosb_plan* plan, uint64_t sim_id0, int S, double* err, int* status) {
    MCOSB_CHECK_CTX(ctx);
    if (!plan || S < 0 || !err) return mcosb_fail(ctx, MCOSB_ERR_ARG, "bad argument to mcosb_run");
    if (!ctx->have_truth) return mcosb_fail(ctx, MCOSB_ERR_STATE, "mcosb_set_truth has not been called");
    if (plan->n_optimizers < 1 || plan->n_optimizers > 4) return mcosb_fail(ctx, MCOSB_ERR_ARG, "n_optimizers must be 1..4");
    for (int o = 0; o < plan->n_optimizers; ++o)
        if (plan->optimizers[o] < 0 || plan->optimizers[o] > 3) return mcosb_fail(ctx, MCOSB_ERR_ARG, "unknown optimizer");
    const size_t N = ctx->N;
    const int nopt = plan->n_optimizers;
    RunPlan rp;
    MCOSB_TRY(resolve_plan(ctx, plan, rp));
    for (int i = 0; i < MCOSB_N_STAGES; ++i) ctx->stage_ms[i] = 0.0;
    if (S == 0) return MCOSB_OK;

    static const bool trace = getenv("MCOSB_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_start = now();
    Staging st(ctx);
    double* derr;
    int* dstatus;
    MCOSB_TRY(st.out(err, (size_t)S * nopt * rp.nk, SL_ERR, &derr));
    bool uses_rp = false;
    for (int o = 0; o < nopt; ++o) uses_rp |= plan->optimizers[o] == MCOSB_OPT_RISK_PARITY;
    if (uses_rp && plan->risk_budget) MCOSB_TRY(st.in(plan->risk_budget, N, SL_STAGE_IN0, &rp.d_budget));
    if (status) MCOSB_TRY(st.out(status, (size_t)S, SL_STATUS, &dstatus));
    else MCOSB_TRY(mcosb_scratch_t(ctx, SL_STATUS, (size_t)S, &dstatus));
    MCOSB_CUDA(ctx, cudaMemsetAsync(dstatus, 0, (size_t)S * sizeof(int), ctx->stream));

    if (ctx->ideal_pending) {   // left over from a call that failed before its first wave
        MCOSB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_ideal, 0));
        ctx->ideal_pending = false;
    }
    MCOSB_PROFILE_MARK(ctx);
    // ---- ideal allocations on the true (mu, cov): once per (truth, optimizer list), kept in the context -----------------
    double* d_ideal;
    int* d_ideal_status;
    MCOSB_TRY(mcosb_scratch_t(ctx, SL_IDEAL, (size_t)4 * N + 8, &d_ideal));
    d_ideal_status = reinterpret_cast<int*>(d_ideal + (size_t)4 * N);
    char key[160];
    snprintf(key, sizeof key, "%d|%d,%d,%d,%d|%.17g|%d|%d", nopt, plan->optimizers[0], plan->optimizers[1],
             plan->optimizers[2], plan->optimizers[3], plan->risk_free_rate, plan->nco_max_clusters, plan->nco_trials);
    if (ctx->ideal_key != key || rp.d_budget) {   // a custom risk budget is not part of the key: recompute
        ctx->ideal_key.clear();
        // On the auxiliary (high-priority) stream, behind everything already queued on the main one: the one-CTA kernels of
        // the ideal allocation then overlap the sampling of the first wave instead of idling the GPU ahead of it.  With
        // per-kernel profiling on, everything stays on the main stream (a kernel's time is the gap between its events).
        static const bool ideal_serial = getenv("MCOSB_IDEAL_SERIAL") != nullptr;
        cudaStream_t main_stream = ctx->stream;
        const bool overlap = !ctx->profile && !ideal_serial && ctx->aux_stream;
        if (overlap) {
            MCOSB_CUDA(ctx, cudaEventRecord(ctx->ev_main, main_stream));
            MCOSB_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_main, 0));
            ctx->stream = ctx->aux_stream;
        }
        struct Restore {
            mcosb_ctx* c; cudaStream_t s;
            ~Restore() { c->stream = s; }
        } restore{ctx, main_stream};
        MCOSB_CUDA(ctx, cudaMemsetAsync(d_ideal_status, 0, sizeof(int), ctx->stream));
        for (int o = 0; o < nopt; ++o) {
            double* wi = d_ideal + (size_t)o * N;
            const int kind = plan->optimizers[o];
            if (kind == MCOSB_OPT_MARKOWITZ)
                MCOSB_TRY(mcosb_dev_markowitz(ctx, 1, ctx->d_mu, ctx->d_cov, plan->risk_free_rate, 1, wi, nullptr, d_ideal_status));
            else if (kind == MCOSB_OPT_HRP)
                MCOSB_TRY(mcosb_dev_hrp(ctx, 1, ctx->d_cov, wi, nullptr, nullptr, d_ideal_status));
            else if (kind == MCOSB_OPT_RISK_PARITY)
                MCOSB_TRY(mcosb_dev_risk_parity(ctx, 1, ctx->d_cov, rp.d_budget, wi, nullptr, d_ideal_status));
            else
                MCOSB_TRY(mcosb_dev_nco(ctx, 1, ctx->d_mu, ctx->d_cov, plan->nco_max_clusters, plan->nco_trials, nullptr, wi,
                                        nullptr, d_ideal_status));
        }
        if (overlap) {
            MCOSB_CUDA(ctx, cudaEventRecord(ctx->ev_ideal, ctx->aux_stream));
            ctx->ideal_pending = true;
        }
        if (!rp.d_budget) ctx->ideal_key = key;
    }

    const double t_ideal = now();
    // ---- wave size from free HBM -----------------------------------------------------------------------------------------
    int wave = plan->wave;
    if (wave <= 0) {
        // cudaMemGetInfo takes 0.1 - 40 ms on this platform and would leave the GPU idle between the ideal allocation and
        // the first wave: ask once per context and remember (free + what the context already holds)
        size_t held = 0;
        for (int i = 0; i < SL_COUNT; ++i) held += ctx->slot_bytes[i];
        if (ctx->mem_budget == 0) {
            size_t free_once = 0, total_b = 0;
            MCOSB_CUDA(ctx, cudaMemGetInfo(&free_once, &total_b));
            ctx->mem_budget = free_once + held;
        }
        const size_t free_b = ctx->mem_budget > held ? ctx->mem_budget - held : 0;
        // per simulation: cov + denoise work matrix + HRP D and E (4 N^2), inverse-iteration scratch, small vectors
        size_t per_sim = (5 * N * N + 6 * N * 16 + 16 * N) * sizeof(double) + sbr3_scratch_bytes(N);   // + panel arrays and Z partials of the band reduction
        for (int o = 0; o < nopt; ++o)
            if (plan->optimizers[o] == MCOSB_OPT_NCO) {   // NCO: cluster-block work matrix, centres when they spill to HBM
                const size_t mk = plan->nco_max_clusters > 0 ? (size_t)plan->nco_max_clusters : N / 2;
                per_sim += (N * N + 4 * N + (2 * mk * N * sizeof(double) > 200000 ? 2 * mk * N : 0)) * sizeof(double);
                break;
            }
        size_t budget = (free_b + held) / 2;
        size_t fixed = 2 * RUN_CHUNK_BYTES + ((size_t)1 << 28);  // X + Z chunk buffers
        size_t w = budget > fixed ? (budget - fixed) / per_sim : 1;
        wave = (int)std::max<size_t>(1, std::min<size_t>(w, 8192));
    }
    wave = std::min(wave, S);
    if (plan->wave <= 0 && S > wave) {
        // equal waves instead of full ones and a short remainder (every kernel of a short wave ends on a mostly idle GPU):
        // the simulations divided over the same number of waves, rounded up to whole CTAs-per-SM rounds
        const int nw = (S + wave - 1) / wave;
        const int even = (S + nw - 1) / nw;
        wave = std::min(wave, (even + ctx->num_sms - 1) / ctx->num_sms * ctx->num_sms);
    }

    const double t_mem = now();
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> marks;
    int rc = MCOSB_OK;
    for (int s0 = 0; s0 < S && rc == MCOSB_OK; s0 += wave) {
        const int W = std::min(wave, S - s0);
        rc = mcosb_run_wave(ctx, plan, rp, sim_id0 + s0, W, d_ideal, derr + (size_t)s0 * nopt * rp.nk, dstatus + s0, marks);
    }
    if (rc == MCOSB_OK) rc = st.finish();
    const double t_launched = now();
    cudaError_t se = cudaStreamSynchronize(ctx->stream);
    const double t_synced = now();
    for (auto& m : marks) {
        if (m.second.first && m.second.second && se == cudaSuccess) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, m.second.first, m.second.second) == cudaSuccess) ctx->stage_ms[m.first] += ms;
        }
        if (m.second.first) cudaEventDestroy(m.second.first);
        if (m.second.second) cudaEventDestroy(m.second.second);
    }
    if (rc != MCOSB_OK) return rc;
    if (se != cudaSuccess) return mcosb_fail(ctx, MCOSB_ERR_CUDA, cudaGetErrorString(se));
    int ideal_status = 0;
    MCOSB_CUDA(ctx, cudaMemcpy(&ideal_status, d_ideal_status, sizeof(int), cudaMemcpyDeviceToHost));
    if (trace)
        fprintf(stderr, "mcosb_run host ms: ideal-launch %.2f memgetinfo %.2f wave-launch %.2f sync-wait %.2f tail %.2f\n",
                t_ideal - t_start, t_mem - t_ideal, t_launched - t_mem, t_synced - t_launched, now() - t_synced);
    ideal_status &= ~MCOSB_ST_NO_EXCESS;   // informational: the Markowitz vertex allocation is defined (markowitz.cu)
    if (ideal_status) {
        char buf[128];
        snprintf(buf, sizeof buf, "ideal allocation on the true (mu, cov) failed with status bits %d", ideal_status);
        return mcosb_fail(ctx, MCOSB_ERR_IDEAL, buf);
    }
    return MCOSB_OK;
}
