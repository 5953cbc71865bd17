"""GPU parity at the PUBLISHED configurations, through the path bench.py times (mcosb_run), against the CPU oracle on the
same device-drawn observations (the loop body of mcos/mcos.py:22-33):

  config 3   N = 500, T = 1000, Ledoit-Wolf + DeNoiser, Markowitz + HRP, Sharpe error         (the headline; two-stage
             eigensolver, Ledoit-Wolf shrink folded into the correlation pass, HRP tensor-core fast path -- the defaults)
  config 5   N = 1000, T = 2000, Ledoit-Wolf + DeNoiser, Markowitz + HRP + NCO(10), all three estimators in one pass
  config 4   N = 100, T = 200, Jackknife + NCO(10 clusters, 10 trials) in the fused loop
  config 2   N = 100, T = 200, MuCov + DeNoiser + Markowitz

Bars.  HRP errors 1e-7 relative (cluster order bit-exact, so only the de-noised covariance differs).  Markowitz errors:
`clean_weights` rounds to 5 decimals, so a 1e-12 difference before rounding can move a weight by 1e-5; every simulation
whose oracle pre-rounding weights stay 1e-7 clear of every rounding boundary must match at 1e-9, the others are counted
and their cleaned weights may differ from the oracle's only by single 1e-5 flips at those near-boundary entries.
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose

import oracle
import mcos_b200 as mb
from mcos_b200 import _lib
from mcos_b200.engine import Engine
from mcos_b200.synthetic import block_diagonal_truth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1800)]

KIND_NAME = {_lib.ERR_EXPECTED_OUTCOME: "expected_outcome", _lib.ERR_VARIANCE: "variance", _lib.ERR_SHARPE: "sharpe"}


def near_rounding_boundary(w, margin=1e-7):
    """Entries of pre-rounding weights within `margin` of a clean_weights decision: the 1e-4 cutoff or a half-integer
    multiple of 1e-5 (np.round(w, 5))."""
    w = np.asarray(w)
    scaled = np.abs(w) * 1e5
    to_half = np.abs(scaled - np.floor(scaled) - 0.5) * 1e-5
    return (np.abs(np.abs(w) - 1e-4) < margin) | ((np.abs(w) >= 1e-4 - margin) & (to_half < margin))


def _same_partition(a, b):
    return len(set(zip(a.tolist(), b.tolist()))) == len(set(a.tolist())) == len(set(b.tolist()))


def _oracle_moments(kind, x):
    return {"mucov": oracle.moments_mucov, "mucovledoitwolf": oracle.moments_ledoit_wolf,
            "jackknife": oracle.moments_jackknife}[kind](x)


def _check_markowitz(eng, mu, cov, mu_hat_gpu, cov_gpu, err_gpu, mu_hat_o, cov_o, ideal_o, kinds):
    """err_gpu [S, n_kinds] of the fused run against the exact-QP oracle per the rule in the module docstring.
    Returns (n_safe, n_near)."""
    s = err_gpu.shape[0]
    # the ideal allocation (mcos.py:30) obeys the same rule: equal to the oracle's unless the truth itself sits on a
    # rounding boundary, in which case the comparison below uses the kernel's (one flip at a near-boundary entry)
    ideal_gpu = eng.markowitz(mu[None], cov[None], clean=True)[0][0]
    near_ideal = near_rounding_boundary(oracle.markowitz_exact_qp(mu, cov)[0])
    assert (ideal_gpu[~near_ideal] == ideal_o[~near_ideal]).all() and np.abs(ideal_gpu - ideal_o).max() <= 1e-5 + 1e-12
    ideal_o = ideal_gpu
    raw_gpu, info = eng.markowitz(mu_hat_gpu, cov_gpu, clean=False)
    clean_gpu, _ = eng.markowitz(mu_hat_gpu, cov_gpu, clean=True)
    assert (info["status"] == 0).all()
    n_safe = n_near = 0
    for i in range(s):
        raw_o, _ = oracle.markowitz_exact_qp(mu_hat_o[i], cov_o[i])
        assert_allclose(raw_gpu[i], raw_o, rtol=1e-6, atol=1e-10)          # 1e-6 after de-noising (north_star)
        near = near_rounding_boundary(raw_o)
        clean_o = oracle.clean_weights(raw_o)
        want = np.array([oracle.estimate_error(KIND_NAME[k], mu, cov, clean_o, ideal_o) for k in kinds])
        if not near.any():
            n_safe += 1
            assert np.array_equal(clean_gpu[i], clean_o), f"sim {i}: cleaned weights differ away from any boundary"
            assert_allclose(err_gpu[i], want, rtol=1e-9, atol=1e-15, err_msg=f"sim {i}")
        else:
            n_near += 1
            diff = clean_gpu[i] - clean_o
            assert (diff[~near] == 0).all(), f"sim {i}: flip away from a rounding boundary"
            assert np.abs(diff).max() <= 1e-5 + 1e-12, f"sim {i}: more than one 1e-5 flip per entry"
            got = np.array([oracle.estimate_error(KIND_NAME[k], mu, cov, clean_gpu[i], ideal_o) for k in kinds])
            assert_allclose(err_gpu[i], got, rtol=1e-9, atol=1e-15, err_msg=f"sim {i} (own cleaned weights)")
    return n_safe, n_near


def _run_config(n, t, s, sim_name, moments_kind, denoise, optimizers, kinds, seed, nco=(10, 10), wave=0):
    """Runs mcosb_run and the oracle on the same draws; returns everything the checks need."""
    mu, cov = block_diagonal_truth(n)
    eng = Engine(n, t, seed=seed)
    eng.set_truth(mu, cov)
    plan = _lib.Plan()
    plan.moments_kind, plan.denoise, plan.bandwidth, plan.n_optimizers = moments_kind, int(denoise), 0.25, len(optimizers)
    for i, o in enumerate(optimizers):
        plan.optimizers[i] = o
    plan.error_kind, plan.risk_free_rate, plan.wave = kinds[0], 0.02, wave
    plan.nco_max_clusters, plan.nco_trials = nco
    if len(kinds) > 1:
        plan.n_error_kinds = len(kinds)
        for k, kind in enumerate(kinds):
            plan.error_kinds[k] = kind
    errs, status = eng.run(plan, s, sim_id0=100)
    errs = errs.reshape(s, len(optimizers), len(kinds))
    assert (status == 0).all(), status
    x = eng.sample(s, sim_id0=100)
    # the GPU's own stage outputs (fused == stage by stage, bit for bit: tests/test_gpu_run.py)
    mu_hat_gpu, cov_gpu, _ = eng.sample_moments(s, 100, moments_kind)
    if denoise:
        cov_gpu, _ = eng.denoise(cov_gpu, t)
    # the oracle's
    mu_hat_o, cov_o = [], []
    for i in range(s):
        m, c = _oracle_moments(sim_name, x[i])
        mu_hat_o.append(m.reshape(-1))
        cov_o.append(oracle.denoise(c, t) if denoise else c)
    return dict(mu=mu, cov=cov, eng=eng, errs=errs, x=x, mu_hat_gpu=mu_hat_gpu, cov_gpu=cov_gpu,
                mu_hat_o=np.stack(mu_hat_o), cov_o=np.stack(cov_o))


def test_config3_headline_through_the_fused_path():
    """BASELINE config 3 at full size, default eigensolver (two-stage) and HRP mode (tensor-core fast path)."""
    n, t, s = 500, 1000, 8
    kinds = [_lib.ERR_SHARPE]
    r = _run_config(n, t, s, "mucovledoitwolf", _lib.MOMENTS_LEDOIT_WOLF, True, [_lib.OPT_MARKOWITZ, _lib.OPT_HRP], kinds, 2024)
    mu, cov = r["mu"], r["cov"]
    assert_allclose(r["cov_gpu"], r["cov_o"], rtol=1e-6, atol=1e-12)
    ideal_m = oracle.markowitz_exact(mu, cov)
    ideal_h = oracle.hrp(cov)
    n_safe, n_near = _check_markowitz(r["eng"], mu, cov, r["mu_hat_gpu"], r["cov_gpu"], r["errs"][:, 0, :],
                                      r["mu_hat_o"], r["cov_o"], ideal_m, kinds)
    print(f"config 3: Markowitz {n_safe} simulations matched at 1e-9, {n_near} near a rounding boundary")
    assert n_safe >= 1
    for i in range(s):
        d = oracle.hrp_details(r["cov_o"][i])
        want = oracle.err_sharpe(mu, cov, d["weights"], ideal_h)
        assert_allclose(r["errs"][i, 1, 0], want, rtol=1e-7, atol=1e-14)
    w, info = r["eng"].hrp(r["cov_gpu"])
    for i in range(s):
        assert np.array_equal(info["order"][i], oracle.hrp_details(r["cov_o"][i])["order"])


def test_config5_shape_all_three_estimators_one_pass():
    """BASELINE config 5's shape on one GPU: N = 1000, T = 2000, Markowitz + HRP + NCO(10), the three estimators from one
    pass.  NCO: the kernel's partition is compared with sklearn's; its weights must be the oracle's for ITS partition."""
    n, t, s = 1000, 2000, 2
    kinds = [_lib.ERR_EXPECTED_OUTCOME, _lib.ERR_VARIANCE, _lib.ERR_SHARPE]
    opts = [_lib.OPT_MARKOWITZ, _lib.OPT_HRP, _lib.OPT_NCO]
    r = _run_config(n, t, s, "mucovledoitwolf", _lib.MOMENTS_LEDOIT_WOLF, True, opts, kinds, 7)
    mu, cov, eng = r["mu"], r["cov"], r["eng"]
    assert_allclose(r["cov_gpu"], r["cov_o"], rtol=1e-6, atol=1e-12)
    ideal_m = oracle.markowitz_exact(mu, cov)
    ideal_h = oracle.hrp(cov)
    n_safe, n_near = _check_markowitz(eng, mu, cov, r["mu_hat_gpu"], r["cov_gpu"], r["errs"][:, 0, :], r["mu_hat_o"],
                                      r["cov_o"], ideal_m, kinds)
    print(f"config 5: Markowitz {n_safe} matched at 1e-9, {n_near} near a rounding boundary")
    for i in range(s):
        wh = oracle.hrp(r["cov_o"][i])
        want = [oracle.estimate_error(KIND_NAME[k], mu, cov, wh, ideal_h) for k in kinds]
        assert_allclose(r["errs"][i, 1, :], want, rtol=1e-7, atol=1e-14)
    # NCO: ideal allocation and per-simulation allocations for the kernel's own partitions
    wi, ii = eng.nco(mu[None], cov[None], 10, 10)
    assert_allclose(wi[0], oracle.nco(mu, cov, labels=ii["labels"][0]), rtol=1e-8, atol=1e-11)
    wn, info = eng.nco(r["mu_hat_gpu"], r["cov_gpu"], 10, 10)
    agree = 0
    for i in range(s):
        got = info["labels"][i]
        w_o = oracle.nco(r["mu_hat_o"][i], r["cov_o"][i], labels=got)
        assert_allclose(wn[i], w_o, rtol=1e-6, atol=1e-9)
        want = [oracle.estimate_error(KIND_NAME[k], mu, cov, w_o, wi[0]) for k in kinds]
        assert_allclose(r["errs"][i, 2, :], want, rtol=1e-6, atol=1e-12)
        ref_labels, _ = oracle.nco_cluster(oracle.cov_to_corr(r["cov_o"][i]), 10, 1)
        agree += _same_partition(got, ref_labels)
    print(f"config 5: NCO partition equal to sklearn's in {agree}/{s} simulations")
    assert agree == s


def test_config5_one_pass_equals_three_single_estimator_runs():
    """err[S][n_opt][3] of one run is bit-equal to three runs with one estimator each (error_estimator.py:21-49)."""
    n, t, s = 200, 400, 10
    mu, cov = block_diagonal_truth(n)
    sim = mb.MuCovLedoitWolfObservationSimulator(mu, cov, t, seed=3)
    opts = [mb.MarkowitzOptimizer(), mb.HRPOptimizer(), mb.NCOOptimizer(10), mb.RiskParityOptimizer()]
    ests = [mb.ExpectedOutcomeErrorEstimator(), mb.VarianceErrorEstimator(), mb.SharpeRatioErrorEstimator()]
    tr = [mb.DeNoiserCovarianceTransformer()]
    all3, st = mb.simulate_errors(sim, s, opts, ests, tr)
    assert all3.shape == (s, 4, 3) and (st == 0).all()
    for k, est in enumerate(ests):
        one, _ = mb.simulate_errors(sim, s, opts, est, tr)
        assert one.shape == (s, 4)
        assert np.array_equal(one, all3[:, :, k], equal_nan=True)
    tabs = mb.simulate_optimizations(sim, s, opts, ests, tr)
    assert set(tabs) == {"ExpectedOutcomeErrorEstimator", "VarianceErrorEstimator", "SharpeRatioErrorEstimator"}
    assert list(tabs["VarianceErrorEstimator"].index) == ["markowitz", "HRP", "NCO", "Risk Parity"]
    eng = Engine(n, t)
    eng.set_truth(mu, cov)
    w = np.random.RandomState(0).dirichlet(np.ones(n), size=5)
    multi = eng.errors_multi(w, w[0], [2, 0, 1])
    for col, kind in enumerate([2, 0, 1]):
        assert np.array_equal(multi[:, col], eng.errors(w, w[0], kind), equal_nan=True)


def test_config4_jackknife_nco_in_the_fused_loop():
    """BASELINE config 4: N = 100, T = 200, Jackknife + NCO(max_num_clusters = 10, num_clustering_trials = 10)."""
    n, t, s = 100, 200, 12
    kinds = [_lib.ERR_VARIANCE]
    r = _run_config(n, t, s, "jackknife", _lib.MOMENTS_JACKKNIFE, False, [_lib.OPT_NCO], kinds, 11, wave=5)
    mu, cov, eng = r["mu"], r["cov"], r["eng"]
    assert_allclose(r["cov_gpu"], r["cov_o"], rtol=1e-10, atol=1e-16)
    wi, ii = eng.nco(mu[None], cov[None], 10, 10)
    assert_allclose(wi[0], oracle.nco(mu, cov, labels=ii["labels"][0]), rtol=1e-8, atol=1e-11)
    wn, info = eng.nco(r["mu_hat_gpu"], r["cov_gpu"], 10, 10)
    agree = 0
    for i in range(s):
        got = info["labels"][i]
        w_o = oracle.nco(r["mu_hat_o"][i], r["cov_o"][i], labels=got)
        assert_allclose(wn[i], w_o, rtol=1e-7, atol=1e-10)
        assert_allclose(r["errs"][i, 0, 0], oracle.err_variance(mu, cov, w_o, wi[0]), rtol=1e-7, atol=1e-14)
        ref_labels, _ = oracle.nco_cluster(oracle.cov_to_corr(r["cov_o"][i]), 10, 1)
        agree += _same_partition(got, ref_labels)
    print(f"config 4: NCO partition equal to sklearn's in {agree}/{s} simulations")
    assert agree >= s - 1


def test_config2_mucov_denoise_markowitz():
    """BASELINE config 2: N = 100, T = 200, MuCov + DeNoiser + Markowitz, Variance error."""
    n, t, s = 100, 200, 24
    kinds = [_lib.ERR_VARIANCE]
    r = _run_config(n, t, s, "mucov", _lib.MOMENTS_MUCOV, True, [_lib.OPT_MARKOWITZ], kinds, 5, wave=7)
    ideal_m = oracle.markowitz_exact(r["mu"], r["cov"])
    n_safe, n_near = _check_markowitz(r["eng"], r["mu"], r["cov"], r["mu_hat_gpu"], r["cov_gpu"], r["errs"][:, 0, :],
                                      r["mu_hat_o"], r["cov_o"], ideal_m, kinds)
    print(f"config 2: Markowitz {n_safe} matched at 1e-9, {n_near} near a rounding boundary")
    assert n_safe >= 1 and n_safe + n_near == s


def test_nco_partition_agreement_rate_config4_shape():
    """The KMeans / silhouette search against scikit-learn on 500 sampled covariances of config 4's shape: the partitions
    must be sklearn's in >= 98 % of the simulations; every disagreement is reported with both silhouette qualities
    (evaluated by sklearn on the same distance matrix) so a tie-break can be told from a worse partition.  Reference:
    mcos/optimizer.py:140-174 (parity vs the reference's sklearn 0.22.1 stream stays unpinned, see oracle/)."""
    from sklearn.metrics import silhouette_samples
    n, t, s = 100, 200, 500
    mu, cov = block_diagonal_truth(n)
    eng = Engine(n, t, seed=17)
    eng.set_truth(mu, cov)
    x = eng.sample(s)
    mu_hat, covs, _ = eng.moments(x, _lib.MOMENTS_JACKKNIFE)
    w, info = eng.nco(mu_hat, covs, 10, 10)
    bad = []
    for i in range(s):
        corr = oracle.cov_to_corr(covs[i])
        ref_labels, ref_q = oracle.nco_cluster(corr, 10, 1)
        if not _same_partition(info["labels"][i], ref_labels):
            dist = ((1.0 - corr) / 2.0) ** 0.5
            sil = silhouette_samples(dist, info["labels"][i])
            bad.append((i, len(set(info["labels"][i].tolist())), len(set(ref_labels.tolist())), sil.mean() / sil.std(), ref_q))
    for i, kg, kr, qg, qr in bad:
        print(f"  sim {i}: kernel k={kg} quality {qg:.12g} | sklearn k={kr} quality {qr:.12g}")
    print(f"NCO partition agreement with sklearn: {s - len(bad)}/{s}")
    assert len(bad) <= s // 50


def test_nco_default_cluster_count_at_n256():
    """NCOOptimizer() -- max_num_clusters = None = N / 2 (optimizer.py:152-153) -- needs 8 N^2 bytes for the centres: they
    spill from shared memory to per-simulation global scratch above N ~ 165.  The partition and weights must not depend
    on where the centres live: N = 256 with max_k = 12 (shared memory) equals the same search capped at 12 run through
    the spill path, and the default search runs to completion."""
    n, t, s = 256, 512, 2
    mu, cov = block_diagonal_truth(n, n_blocks=8)
    eng = Engine(n, t, seed=1)
    eng.set_truth(mu, cov)
    x = eng.sample(s)
    mu_hat, covs, _ = eng.moments(x, _lib.MOMENTS_MUCOV)
    w_def, info_def = eng.nco(mu_hat, covs, None, 10)            # 2 * 128 * 256 doubles = 512 KB -> global scratch
    assert (info_def["status"] == 0).all() and np.allclose(w_def.sum(1), 1.0, atol=1e-9)
    for i in range(s):
        assert_allclose(w_def[i], oracle.nco(mu_hat[i], covs[i], labels=info_def["labels"][i]), rtol=1e-8, atol=1e-11)
        ref_labels, _ = oracle.nco_cluster(oracle.cov_to_corr(covs[i]), None, 1)
        assert _same_partition(info_def["labels"][i], ref_labels)
