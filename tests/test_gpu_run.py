"""GPU: the fused loop (mcosb_run / simulate_optimizations) against the stage-by-stage path and the CPU oracle."""
import numpy as np
import pandas as pd
import pytest
from numpy.testing import assert_allclose, assert_almost_equal

import oracle
import mcos_b200 as mb
from mcos_b200 import _lib
from mcos_b200.engine import Engine, get_engine
from mcos_b200.synthetic import block_diagonal_truth

pytestmark = pytest.mark.gpu


def _plan(moments, denoise, opts, err, wave=0):
    p = _lib.Plan()
    p.moments_kind, p.denoise, p.bandwidth, p.n_optimizers = moments, int(denoise), 0.25, len(opts)
    for i, o in enumerate(opts):
        p.optimizers[i] = o
    p.error_kind, p.risk_free_rate, p.nco_max_clusters, p.nco_trials, p.wave = err, 0.02, 0, 10, wave
    return p


@pytest.mark.parametrize("n,t,moments,denoise,err", [
    (100, 200, _lib.MOMENTS_MUCOV, True, _lib.ERR_VARIANCE),            # BASELINE config 2 shape
    (60, 120, _lib.MOMENTS_LEDOIT_WOLF, True, _lib.ERR_SHARPE),
    (20, 50, _lib.MOMENTS_JACKKNIFE, False, _lib.ERR_EXPECTED_OUTCOME),
])
def test_run_equals_stage_by_stage_and_oracle(n, t, moments, denoise, err):
    mu, cov = block_diagonal_truth(n, n_blocks=10 if n % 10 == 0 and n >= 100 else 2)
    s = 12
    eng = Engine(n, t, seed=11)
    eng.set_truth(mu, cov)
    plan = _plan(moments, denoise, [_lib.OPT_MARKOWITZ, _lib.OPT_HRP], err, wave=5)   # 3 ragged waves
    errs, status = eng.run(plan, s, sim_id0=3)
    assert (status == 0).all()
    # the same simulations, one stage at a time through the C ABI.  mcosb_sample_moments is the sampling + moments stage as
    # the fused loop runs it (draws kept centred by the true mean); the draws themselves (mcosb_sample) feed the oracle, and
    # mcosb_moments on them agrees to rounding
    x = eng.sample(s, sim_id0=3)
    mu_hat, cov_hat, _ = eng.sample_moments(s, 3, moments)
    mu_x, cov_x, _ = eng.moments(x, moments)
    assert_allclose(mu_hat, mu_x, rtol=1e-12, atol=1e-15)
    assert_allclose(cov_hat, cov_x, rtol=1e-11, atol=1e-12 * np.abs(cov_x).max())
    if denoise:
        cov_hat, _ = eng.denoise(cov_hat, t)
    wm, _ = eng.markowitz(mu_hat, cov_hat)
    wh, _ = eng.hrp(cov_hat)
    im, _ = eng.markowitz(mu[None], cov[None])
    ih, _ = eng.hrp(cov[None])
    assert np.array_equal(errs[:, 0], eng.errors(wm, im[0], err))
    assert np.array_equal(errs[:, 1], eng.errors(wh, ih[0], err))
    # and the CPU oracle on the same draws (Markowitz through the exact QP + clean_weights)
    kind = {0: "expected_outcome", 1: "variance", 2: "sharpe"}[err]
    sim = {0: "mucov", 1: "mucovledoitwolf", 2: "jackknife"}[moments]
    tab, samples = oracle.simulate_optimizations(mu, cov, t, s, simulator=sim,
                                                 transformers=[("denoise", 0.25)] if denoise else [],
                                                 optimizers=("markowitz_exact", "HRP"), estimator=kind,
                                                 observations=list(x), return_samples=True)
    # Markowitz: clean_weights rounds to 1e-5, so a 1e-12 difference before rounding can move a weight by 1e-5
    assert_allclose(errs[:, 0], np.ravel(samples["markowitz_exact"]), rtol=2e-3, atol=1e-7)
    assert_allclose(errs[:, 1], np.ravel(samples["HRP"]), rtol=1e-7, atol=1e-14)


def test_run_is_independent_of_wave_and_batch_split():
    mu, cov = block_diagonal_truth(40, n_blocks=4)
    eng = Engine(40, 80, seed=5)
    eng.set_truth(mu, cov)
    a, _ = eng.run(_plan(1, True, [0, 1], 2, wave=0), 20)
    b, _ = eng.run(_plan(1, True, [0, 1], 2, wave=7), 20)
    c0, _ = eng.run(_plan(1, True, [0, 1], 2, wave=0), 9, sim_id0=0)
    c1, _ = eng.run(_plan(1, True, [0, 1], 2, wave=0), 11, sim_id0=9)
    assert np.array_equal(a, b) and np.array_equal(a, np.concatenate([c0, c1]))


def test_automatic_waves_are_equal_and_do_not_change_results():
    """More simulations than one automatic wave (capped at 8192): mcosb_run cuts the job into equal waves rounded to the SM
    count instead of full waves plus a short remainder; every simulation must come out as in an explicit small-wave run."""
    mu, cov = block_diagonal_truth(20, n_blocks=2)
    eng = Engine(20, 40, seed=3)
    eng.set_truth(mu, cov)
    s = 17000                                             # 3 automatic waves
    a, sa = eng.run(_plan(0, True, [0, 1], 1, wave=0), s)
    b, sb = eng.run(_plan(0, True, [0, 1], 1, wave=2999), s)
    assert np.array_equal(a, b, equal_nan=True) and np.array_equal(sa, sb)


def test_simulate_optimizations_readme_sample(stock):
    """BASELINE config 1: the README sample (reference README.md / tests/test_mcos.py shape): 20 stocks, MuCov,
    DeNoiser, Markowitz + HRP, VarianceErrorEstimator -> DataFrame[mean, stdev] indexed by optimizer."""
    sim = mb.MuCovObservationSimulator(stock["mu"], stock["cov"], n_observations=50)
    df = mb.simulate_optimizations(sim, n_sims=50, optimizers=[mb.MarkowitzOptimizer(), mb.HRPOptimizer()],
                                   error_estimator=mb.VarianceErrorEstimator(),
                                   covariance_transformers=[mb.DeNoiserCovarianceTransformer()])
    assert isinstance(df, pd.DataFrame) and list(df.index) == ["markowitz", "HRP"]
    assert df.index.name == "optimizer" and list(df.columns) == ["mean", "stdev"]
    assert np.isfinite(df.values).all() and (df["mean"] > 0).all()


def test_error_distribution_matches_cpu_reference_path(stock):
    """Distributional agreement with the GPU RNG (north_star): mean error of HRP / Markowitz over many GPU simulations
    agrees with the CPU path (numpy RNG, reference algorithms) within Monte-Carlo standard error."""
    mu, cov = stock["mu"], stock["cov"]
    t, n_gpu, n_cpu = 50, 4000, 60
    sim = mb.MuCovObservationSimulator(mu, cov, n_observations=t, seed=3)
    errs, _ = mb.simulate_errors(sim, n_gpu, [mb.MarkowitzOptimizer(), mb.HRPOptimizer()], mb.VarianceErrorEstimator(),
                                 [mb.DeNoiserCovarianceTransformer()])
    np.random.seed(0)
    _, cpu = oracle.simulate_optimizations(mu, cov, t, n_cpu, simulator="mucov", transformers=[("denoise", 0.25)],
                                           optimizers=("markowitz", "HRP"), estimator="variance", return_samples=True)
    for col, name in enumerate(("markowitz", "HRP")):
        c = np.ravel(cpu[name])
        se = np.sqrt(errs[:, col].var() / n_gpu + c.var() / n_cpu)
        assert abs(errs[:, col].mean() - c.mean()) < 5 * se


def test_plugin_methods_single_inputs(stock, literals):
    """The reference's unit tests pointed at the GPU classes (tests/test_optimizer.py, test_error_estimator.py,
    test_covariance_transformer.py of the reference)."""
    mu, cov = stock["mu"], stock["cov"]
    assert np.abs(mb.MarkowitzOptimizer().allocate(mu, cov) - literals["markowitz_w"]).max() <= 2e-4
    assert_almost_equal(mb.HRPOptimizer().allocate(mu, cov), literals["hrp_w"])        # 8-decimal golden
    out = mb.DeNoiserCovarianceTransformer().transform(cov, int(literals["denoise_n_observations"]))
    assert out.shape == (20, 20)
    assert_almost_equal(out, literals["denoised_cov"])
    a, o = literals["est_allocation"], literals["est_optimal_allocation"]
    for est, want in zip((mb.ExpectedOutcomeErrorEstimator(), mb.VarianceErrorEstimator(), mb.SharpeRatioErrorEstimator()),
                         literals["est_expected"]):
        assert abs(est.estimate(mu, cov, a, o) - want) < 1.5e-7   # assert_almost_equal(decimal=7), as the reference
    assert mb.MarkowitzOptimizer().name == "markowitz" and mb.HRPOptimizer().name == "HRP" and mb.NCOOptimizer().name == "NCO"
    mh, ch = mb.MuCovObservationSimulator(mu, cov, 50).simulate()
    assert mh.shape == (20, 1) and ch.shape == (20, 20)
    with pytest.raises(ValueError):
        mb.simulate_optimizations_from_price_history(pd.DataFrame(stock["prices"]), "nope", 5, 2,
                                                     [mb.HRPOptimizer()], mb.VarianceErrorEstimator(), [])


def test_from_price_history(stock):
    prices = pd.DataFrame(stock["prices"], columns=[str(x) for x in stock["tickers"]], index=pd.to_datetime(stock["dates"]))
    df = mb.simulate_optimizations_from_price_history(prices, "MuCovLedoitWolf", 60, 16, [mb.MarkowitzOptimizer(), mb.HRPOptimizer()],
                                                      mb.ExpectedOutcomeErrorEstimator(), [mb.DeNoiserCovarianceTransformer()])
    assert list(df.index) == ["markowitz", "HRP"] and np.isfinite(df.values).all()


def test_user_plugins_run_on_the_host_around_the_gpu_stages():
    """The reference exists to compare arbitrary AbstractOptimizer / AbstractCovarianceTransformer implementations
    (mcos/mcos.py:22-33 calls whatever it is given).  A plugin type the engine has no kernel for is called on the host
    once per simulation with the reference's arguments; the built-in plugins around it stay batched on the GPU."""
    n, t, s = 40, 90, 10
    mu, cov = block_diagonal_truth(n, n_blocks=4)
    sim = mb.MuCovObservationSimulator(mu, cov, t, seed=4)
    calls = {"alloc": 0, "transform": 0, "estimate": 0}

    class InverseVariance(mb.AbstractOptimizer):
        name = "IVP"

        def allocate(self, mu, cov):
            calls["alloc"] += 1
            assert np.asarray(cov).shape == (n, n)
            ivp = 1.0 / np.diag(cov)
            return ivp / ivp.sum()

    class HalfShrink(mb.AbstractCovarianceTransformer):
        def transform(self, cov, n_observations):
            calls["transform"] += 1
            assert n_observations == t
            return 0.5 * cov + 0.5 * np.diag(np.diag(cov))

    class L1Error(mb.AbstractErrorEstimator):
        def estimate(self, mu, cov, allocation, optimal_allocation):
            calls["estimate"] += 1
            return float(np.abs(optimal_allocation - allocation).sum())

    # (1) user optimizer next to a built-in one, built-in transformer and estimator
    opts = [InverseVariance(), mb.HRPOptimizer()]
    errs, status = mb.simulate_errors(sim, s, opts, mb.VarianceErrorEstimator(), [mb.DeNoiserCovarianceTransformer()])
    assert errs.shape == (s, 2) and (status == 0).all() and calls["alloc"] == s + 1      # + the ideal allocation
    fused, _ = mb.simulate_errors(sim, s, [mb.HRPOptimizer()], mb.VarianceErrorEstimator(), [mb.DeNoiserCovarianceTransformer()])
    assert np.array_equal(errs[:, 1], fused[:, 0])                  # the built-in column is what the fused loop gives
    eng = get_engine(n, t, 0, 4)
    _, cov_hat, _ = eng.sample_moments(s, 0, _lib.MOMENTS_MUCOV)
    cov_dn, _ = eng.denoise(cov_hat, t)
    ideal = InverseVariance().allocate(mu, cov)
    want = [oracle.err_variance(mu, cov, InverseVariance().allocate(None, c), ideal) for c in cov_dn]
    assert_allclose(errs[:, 0], want, rtol=1e-12, atol=1e-18)
    # (2) user transformer between two built-in ones, user estimator, a list of estimators
    trs = [mb.DetoneCovarianceTransformer(1), HalfShrink(), mb.DeNoiserCovarianceTransformer()]
    e3, st3 = mb.simulate_errors(sim, s, [mb.HRPOptimizer(), mb.MarkowitzOptimizer()], [L1Error(), mb.VarianceErrorEstimator()], trs)
    assert e3.shape == (s, 2, 2) and calls["transform"] == s and calls["estimate"] == 2 * s
    c1, _ = eng.detone(cov_hat, 1)
    c2 = np.stack([HalfShrink().transform(c, t) for c in c1])
    c3, _ = eng.denoise(c2, t)
    wh, _ = eng.hrp(c3)
    ih, _ = eng.hrp(cov[None])
    assert_allclose(e3[:, 0, 0], np.abs(ih[0] - wh).sum(1), rtol=1e-13)
    assert np.array_equal(e3[:, 0, 1], eng.errors(wh, ih[0], _lib.ERR_VARIANCE))
    # (3) a subclass that overrides nothing stays on the fused path; the table has the reference's layout
    class MyHRP(mb.HRPOptimizer):
        pass
    df = mb.simulate_optimizations(sim, s, [MyHRP(), InverseVariance()], mb.VarianceErrorEstimator(), [])
    assert list(df.index) == ["HRP", "IVP"] and list(df.columns) == ["mean", "stdev"]
    # (4) a user-defined simulator (global numpy RNG, as the reference's own classes)
    class NumpySimulator(mb.AbstractObservationSimulator):
        def __init__(self, mu, cov, n_observations):
            self.mu, self.cov, self.n_observations = mu, cov, n_observations

        def simulate(self):
            x = np.random.multivariate_normal(self.mu, self.cov, size=self.n_observations)
            return x.mean(axis=0).reshape(-1, 1), np.cov(x, rowvar=False)

    np.random.seed(3)
    e4, _ = mb.simulate_errors(NumpySimulator(mu, cov, t), 5, [mb.HRPOptimizer()], mb.VarianceErrorEstimator(), [])
    np.random.seed(3)
    tab = oracle.simulate_optimizations(mu, cov, t, 5, simulator="mucov", optimizers=("HRP",), estimator="variance",
                                        return_samples=True)[1]
    assert_allclose(e4[:, 0], np.ravel(tab["HRP"]), rtol=1e-9)


def test_more_than_four_optimizers_and_risk_budget_in_the_fused_loop():
    """Five optimizers take two passes over the same simulations (Philox draws are keyed by id); a RiskParityOptimizer
    with a custom target_risk (optimizer.py:295-309) travels through the plan."""
    n, t, s = 40, 90, 9
    mu, cov = block_diagonal_truth(n, n_blocks=4)
    sim = mb.MuCovLedoitWolfObservationSimulator(mu, cov, t, seed=8)
    budget = np.linspace(1.0, 2.0, n)
    budget /= budget.sum()
    opts = [mb.MarkowitzOptimizer(), mb.HRPOptimizer(), mb.NCOOptimizer(5), mb.RiskParityOptimizer(),
            mb.RiskParityOptimizer(budget)]
    errs, st = mb.simulate_errors(sim, s, opts, mb.VarianceErrorEstimator(), [mb.DeNoiserCovarianceTransformer()])
    assert errs.shape == (s, 5)
    first4, _ = mb.simulate_errors(sim, s, opts[:4], mb.VarianceErrorEstimator(), [mb.DeNoiserCovarianceTransformer()])
    assert np.array_equal(errs[:, :4], first4)
    eng = get_engine(n, t, 0, 8)
    _, cov_hat, _ = eng.sample_moments(s, 0, _lib.MOMENTS_LEDOIT_WOLF)
    cov_dn, _ = eng.denoise(cov_hat, t)
    w, _ = eng.risk_parity(cov_dn, budget)
    iw, _ = eng.risk_parity(cov[None], budget)
    assert np.array_equal(errs[:, 4], eng.errors(w, iw[0], _lib.ERR_VARIANCE))
    # the budget reached the kernel: at this covariance scale SLSQP's first convergence test fires at the starting point,
    # which is the budget itself (optimizer.py:349-359; riskparity.cu), not 1/N
    assert_allclose(iw[0], budget, rtol=0, atol=1e-15)
    assert not np.allclose(iw[0], 1.0 / n)


def test_transformer_lists_in_any_order():
    """mcos.py:25-26 applies the transformer list in order; [Detone, DeNoiser] and [DeNoiser, Detone, DeNoiser] go
    through the fused loop and equal the stage-by-stage result bit for bit."""
    n, t, s = 60, 150, 7
    mu, cov = block_diagonal_truth(n, n_blocks=6)
    sim = mb.MuCovObservationSimulator(mu, cov, t, seed=6)
    eng = get_engine(n, t, 0, 6)
    eng.ensure_truth(mu, cov)
    _, cov_hat, _ = eng.sample_moments(s, 0, _lib.MOMENTS_MUCOV)
    ih, _ = eng.hrp(cov[None])
    for trs, stages in (([mb.DetoneCovarianceTransformer(1), mb.DeNoiserCovarianceTransformer()], "TD"),
                        ([mb.DeNoiserCovarianceTransformer(), mb.DetoneCovarianceTransformer(2), mb.DeNoiserCovarianceTransformer(0.3)], "DtD")):
        errs, _ = mb.simulate_errors(sim, s, [mb.HRPOptimizer()], mb.VarianceErrorEstimator(), trs)
        c = cov_hat
        for tr in trs:
            c = tr.transform_batch(c, t) if isinstance(tr, mb.DeNoiserCovarianceTransformer) else tr.transform_batch(c)
        wh, _ = eng.hrp(c)
        assert np.array_equal(errs[:, 0], eng.errors(wh, ih[0], _lib.ERR_VARIANCE)), stages
        # and the oracle, same order
        o = cov_hat[0]
        for tr in trs:
            o = oracle.denoise(o, t, tr.bandwidth) if isinstance(tr, mb.DeNoiserCovarianceTransformer) else oracle.detone(o, tr.n_remove)
        assert_allclose(c[0], o, rtol=1e-6, atol=1e-12)


@pytest.mark.gpu
def test_checkpointed_run_resumes_exactly(tmp_path):
    """a run cut off after a chunk and restarted returns what the uninterrupted run returns (simulations
    are keyed by id); a checkpoint of a different run is ignored."""
    mu, cov = block_diagonal_truth(40)
    sim = mb.MuCovLedoitWolfObservationSimulator(mu, cov, 80)
    args = (sim, 300, [mb.MarkowitzOptimizer(), mb.HRPOptimizer()], mb.SharpeRatioErrorEstimator(),
            [mb.DeNoiserCovarianceTransformer()])
    ref, st_ref = mb.simulate_errors(*args)
    ck = str(tmp_path / "run")
    full, st = mb.simulate_errors(*args, checkpoint=ck, checkpoint_every=128)
    np.testing.assert_array_equal(full, ref)
    np.testing.assert_array_equal(st, st_ref)
    path = ck + ".rank0.npz"
    z = dict(np.load(path))
    assert int(z["done"]) == 300
    # pretend the run died after the first chunk, with rubbish where the later chunks would go
    np.savez(path, signature=z["signature"], done=np.array(128), errors=z["errors"][:128], status=z["status"][:128])
    resumed, st2 = mb.simulate_errors(*args, checkpoint=ck, checkpoint_every=128)
    np.testing.assert_array_equal(resumed, ref)
    np.testing.assert_array_equal(st2, st_ref)
    # a checkpoint of another run (different truth) must not be picked up
    np.savez(path, signature=np.array("something else"), done=np.array(300), errors=np.zeros_like(ref), status=st_ref)
    again, _ = mb.simulate_errors(*args, checkpoint=ck, checkpoint_every=128)
    np.testing.assert_array_equal(again, ref)
