"""GPU parity: Markowitz (K4a) through mcosb_markowitz.
Bars (north_star / SURVEY 8d): 1e-9 relative vs the exact-QP oracle before clean_weights; after clean_weights equal up
to rounding ties; <= 2e-4 absolute vs the reference's SLSQP output (SLSQP's own accuracy)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

import oracle
import mcos_b200 as mb
from conftest import run_cases
from mcos_b200 import _lib
from mcos_b200.engine import Engine
from mcos_b200.synthetic import block_diagonal_truth

pytestmark = pytest.mark.gpu


def test_markowitz_golden(stock, literals):
    """reference tests/test_optimizer.py:11-22: the golden is SLSQP's 5-decimal answer."""
    eng = Engine(20, 50)
    w, info = eng.markowitz(stock["mu"][None], stock["cov"][None])
    assert info["status"][0] == 0
    assert np.abs(w[0] - literals["markowitz_w"]).max() <= 2e-4
    assert ((w[0] > 0) == (literals["markowitz_w"] > 0)).all()          # identical support
    raw, _ = eng.markowitz(stock["mu"][None], stock["cov"][None], clean=False)
    want = oracle.markowitz_exact_qp(stock["mu"], stock["cov"])[0]
    assert_allclose(raw[0], want, rtol=1e-9, atol=1e-13)
    assert np.array_equal(w[0], oracle.clean_weights(raw[0]))


@pytest.mark.parametrize("case", ["stock20_T50", "block100_T200"])
def test_markowitz_vs_reference_runs(runs, case):
    keys = run_cases(runs, case)
    mu = np.stack([runs[f"{k}/mucov_mu"].ravel() for k in keys])
    cov = np.stack([runs[f"{k}/denoised"] for k in keys])
    eng = Engine(cov.shape[1], 10)
    w, info = eng.markowitz(mu, cov)
    raw, _ = eng.markowitz(mu, cov, clean=False)
    assert (info["status"] == 0).all()
    for i, k in enumerate(keys):
        assert np.abs(w[i] - runs[f"{k}/markowitz_w"]).max() <= 3e-4    # reference = SLSQP, ~1e-4 accurate
        want = oracle.markowitz_exact_qp(mu[i], cov[i])[0]
        assert_allclose(raw[i], want, rtol=1e-9, atol=1e-13)
        cleaned = oracle.clean_weights(want)
        differs = w[i] != cleaned
        assert differs.sum() == 0 or np.abs(w[i] - cleaned)[differs].max() <= 1.0001e-5  # rounding ties only


@pytest.mark.parametrize("n,t,nsim", [(500, 1000, 4), (100, 200, 16), (50, 400, 8), (1, 5, 2), (2, 50, 3)])
def test_markowitz_vs_exact_oracle(n, t, nsim):
    rng = np.random.RandomState(n * 7 + t)
    if n >= 50:
        mu, cov = block_diagonal_truth(n)
    else:
        cov = np.eye(n) * 0.04 + 0.01
        mu = np.full(n, 0.1) + 0.01 * np.arange(n)
    mus, covs = [], []
    for _ in range(nsim):
        x = rng.multivariate_normal(mu, cov, size=t)
        mus.append(x.mean(0))
        c = np.cov(x, rowvar=False).reshape(n, n)
        covs.append(oracle.denoise(c, t) if n >= 50 else c)
    mus, covs = np.stack(mus), np.stack(covs)
    eng = Engine(n, t)
    raw, info = eng.markowitz(mus, covs, clean=False)
    assert (info["status"] == 0).all()
    for i in range(nsim):
        want, winfo = oracle.markowitz_exact_qp(mus[i], covs[i])
        assert_allclose(raw[i], want, rtol=1e-9, atol=1e-13)
        assert abs(raw[i].sum() - 1.0) < 1e-12 and (raw[i] >= 0).all()


def test_markowitz_no_excess_return_is_flagged():
    """All mu_i <= rf: the max-Sharpe QP does not apply.  The stated problem (PyPortfolioOpt 0.5.2 negative_sharpe on the
    simplex) then has its maximum at a vertex: the asset with the largest (least negative) stand-alone Sharpe ratio.
    The simulation is flagged, gets that allocation, and the batch is not aborted.  The reference's SLSQP run stops at a
    vertex too -- never a better one."""
    mu, cov = block_diagonal_truth(20, n_blocks=2)
    eng = Engine(20, 10)
    low = mu / 252.0
    assert (low < 0.02).all()
    mus = np.stack([mu, low, np.full(20, 0.01)])
    w, info = eng.markowitz(mus, np.stack([cov, cov, cov]))
    assert info["status"][0] == 0 and (info["status"][1:] & _lib.ST_NO_EXCESS).all()
    assert abs(w[0].sum() - 1) < 1e-4
    sd = np.sqrt(np.diag(cov))
    for k, m in ((1, low), (2, np.full(20, 0.01))):
        best = int(np.argmax((m - 0.02) / sd))
        want = np.zeros(20)
        want[best] = 1.0
        assert np.array_equal(w[k], want)
        ref, _ = oracle.markowitz_slsqp(m, cov)          # the reference's own answer: a vertex, objective no better
        assert np.sort(ref)[-1] > 1 - 1e-6
        sharpe = lambda x: (x @ m - 0.02) / np.sqrt(x @ cov @ x)
        assert sharpe(w[k]) >= sharpe(ref) - 1e-12
    # and through the class API: no exception (the reference does not raise here either)
    assert np.array_equal(mb.MarkowitzOptimizer().allocate(low, cov), w[1])


def test_markowitz_mu_as_column_vector(stock):
    """mu_hat arrives as (N, 1) from the simulators (observation_simulator.py:28,40,57)."""
    eng = Engine(20, 50)
    w1, _ = eng.markowitz(stock["mu"].reshape(1, 20, 1), stock["cov"][None])
    w2, _ = eng.markowitz(stock["mu"][None], stock["cov"][None])
    assert np.array_equal(w1, w2)
