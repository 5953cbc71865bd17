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
    # the same simulations, one stage at a time through the C ABI
    x = eng.sample(s, sim_id0=3)
    mu_hat, cov_hat, _ = eng.moments(x, moments)
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


def test_unknown_plugin_types_raise():
    mu, cov = block_diagonal_truth(20, n_blocks=2)
    sim = mb.MuCovObservationSimulator(mu, cov, 30)

    class MyOpt(mb.AbstractOptimizer):
        name = "mine"

        def allocate(self, mu, cov):
            return np.ones(len(cov)) / len(cov)

    with pytest.raises(TypeError):
        mb.simulate_optimizations(sim, 4, [MyOpt()], mb.VarianceErrorEstimator(), [])


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
