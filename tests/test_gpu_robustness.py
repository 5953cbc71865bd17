"""GPU: per-simulation failure handling (the reference aborts the whole run on any exception; the
engine flags the simulation in `status` and keeps the batch going), empty batches, bad arguments."""
import numpy as np
import pytest

import mcos_b200 as mb
from mcos_b200 import _lib
from mcos_b200.engine import Engine
from mcos_b200.synthetic import block_diagonal_truth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def test_nan_in_one_simulation_does_not_poison_or_hang_the_batch():
    n, t = 40, 100
    mu, cov = block_diagonal_truth(n, n_blocks=4)
    rng = np.random.RandomState(0)
    covs = np.stack([np.cov(rng.multivariate_normal(mu, cov, size=t), rowvar=False) for _ in range(4)])
    mus = np.tile(mu, (4, 1))
    bad = covs.copy()
    bad[2, 3, 5] = bad[2, 5, 3] = np.nan
    eng = Engine(n, t)
    good_d, _ = eng.denoise(covs, t)
    d, info = eng.denoise(bad, t)
    assert info["status"][2] != 0 and (info["status"][[0, 1, 3]] == 0).all()
    assert np.array_equal(d[[0, 1, 3]], good_d[[0, 1, 3]])
    for fn in (lambda c: eng.hrp(c), lambda c: eng.markowitz(mus, c), lambda c: eng.nco(mus, c, 4, 1),
               lambda c: eng.risk_parity(c)):
        w_good, _ = fn(covs)
        w_bad, _ = fn(bad)                                # must return (no hang); the clean simulations are untouched
        assert np.array_equal(w_bad[[0, 1, 3]], w_good[[0, 1, 3]])


def test_singular_sample_covariance():
    """T < N: the sample covariance is singular. HRP and Ledoit-Wolf + DeNoiser are well defined; Markowitz on the raw
    singular matrix either solves (the free set stays non-singular) or flags NOT_PD -- it must not hang or return junk."""
    n, t = 30, 12
    mu, cov = block_diagonal_truth(n, n_blocks=3)
    sim = mb.MuCovObservationSimulator(mu, cov, t, seed=5)
    mu_hat, cov_hat = sim.simulate_batch(6)
    eng = Engine(n, t)
    w, info = eng.hrp(cov_hat)
    assert (info["status"] == 0).all() and np.allclose(w.sum(1), 1.0)
    wm, minfo = eng.markowitz(mu_hat, cov_hat, clean=False)
    for i in range(6):
        if minfo["status"][i] == 0:
            assert abs(wm[i].sum() - 1) < 1e-9 and (wm[i] >= 0).all()
        else:
            assert minfo["status"][i] & (_lib.ST_NOT_PD | _lib.ST_QP_CAP | _lib.ST_NO_EXCESS)
    lw = mb.MuCovLedoitWolfObservationSimulator(mu, cov, t, seed=5)
    errs, st = mb.simulate_errors(lw, 6, [mb.MarkowitzOptimizer(), mb.HRPOptimizer()], mb.VarianceErrorEstimator(),
                                  [mb.DeNoiserCovarianceTransformer()])
    assert np.isfinite(errs).all()


def test_empty_batches_and_bad_arguments():
    eng = Engine(10, 20)
    mu, cov = block_diagonal_truth(10, n_blocks=2)
    eng.set_truth(mu, cov)
    assert eng.sample(0).shape == (0, 20, 10)
    w, info = eng.hrp(np.zeros((0, 10, 10)))
    assert w.shape == (0, 10)
    plan = mb.build_plan(mb.MuCovObservationSimulator(mu, cov, 20), [mb.HRPOptimizer()], mb.VarianceErrorEstimator(), [])
    err, st = eng.run(plan, 0)
    assert err.shape == (0, 1)
    with pytest.raises(_lib.MCOSBError):
        eng.detone(cov[None], 17)                          # more than 16 components
    with pytest.raises(_lib.MCOSBError):
        eng.moments(np.zeros((1, 20, 10)), 7)              # unknown estimator kind
    fresh = Engine(10, 20)
    with pytest.raises(_lib.MCOSBError):
        fresh.sample(1)                                    # truth not set
    with pytest.raises(_lib.MCOSBError):
        Engine(0, 5)


def test_truth_below_the_risk_free_rate_runs_and_is_reported():
    """A truth on a daily return scale (every mu_i < rf = 0.02): the reference hands Markowitz to SLSQP and carries on;
    here every simulation gets the best-vertex allocation, the run completes, and the condition is reported."""
    mu, cov = block_diagonal_truth(20, n_blocks=2)
    sim = mb.MuCovObservationSimulator(mu / 252.0, cov / 252.0, 50)
    with pytest.warns(RuntimeWarning, match="NO_EXCESS"):
        errs, status = mb.simulate_errors(sim, 8, [mb.MarkowitzOptimizer(), mb.HRPOptimizer()], mb.VarianceErrorEstimator(), [])
    assert np.isfinite(errs).all() and (status & _lib.ST_NO_EXCESS).all()
    assert mb.status_counts(status) == {"NO_EXCESS": 8}
    tab = mb.summarize_errors(errs, [mb.MarkowitzOptimizer(), mb.HRPOptimizer()], status=status)
    assert (tab["n_flagged"] == 8).all()


@pytest.mark.parametrize("n", [20, 100, 257, 500])
def test_truth_without_a_cholesky_factor_is_refused_by_both_factorisations(n):
    """mcosb_set_truth factorises the truth in blocked form (32-column panels; one-CTA kernel when the panel does not fit
    shared memory): an indefinite truth -- non-positive pivot in the first panel, in a later one, on the last column --
    must come back as MCOSB_ERR_NOT_PD (-4) and a definite one must factorise to numpy's factor."""
    from mcos_b200._lib import MCOSBError as MCOSError
    rng = np.random.RandomState(n)
    a = rng.standard_normal((n, 2 * n))
    good = a @ a.T / (2 * n)
    mu = np.zeros(n)
    eng = Engine(n, 8, seed=1)
    for bad_at in (0, n // 2, n - 1):
        bad = good.copy()
        bad[bad_at, bad_at] = -1.0 if bad_at == 0 else good[bad_at, :bad_at] @ np.linalg.solve(good[:bad_at, :bad_at], good[:bad_at, bad_at]) - 1e-3
        with pytest.raises(MCOSError):
            eng.set_truth(mu, bad)
        eng.set_truth(mu, bad, require_pd=False)          # accepted for the estimators, sampling still refused
    eng.set_truth(mu, good)
    x = eng.sample(1)
    from philox_ref import standard_normals
    want = standard_normals([0], 8, n, 1)[0] @ np.linalg.cholesky(good).T
    np.testing.assert_allclose(x[0], want, rtol=1e-10, atol=1e-12 * np.abs(want).max())
