"""GPU: RiskParityOptimizer (SURVEY 8f "next" #1) through mcosb_risk_parity.
The reference is scipy SLSQP (ftol 1e-6) started at the budget.  Bars: (i) where SLSQP's first convergence test fires the
reference returns the budget unchanged and so must the kernel (exact equality); (ii) where SLSQP converges (the
reference's O(1) golden) the kernel's exact risk-budgeting weights agree to SLSQP's accuracy (<= 1e-3, the tolerance of
the reference's own budget test); (iii) the exact solution equalises the risk contributions to 1e-10."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

import oracle
import mcos_b200 as mb
from mcos_b200 import _lib
from mcos_b200.engine import Engine
from mcos_b200.synthetic import block_diagonal_truth

pytestmark = pytest.mark.gpu


def _risk_contributions(w, cov):
    m = cov @ w
    return w * m / (w @ m)


def test_risk_parity_goldens(literals):
    """reference tests/test_optimizer.py:70-89."""
    cov = literals["rp_cov"]
    w = mb.RiskParityOptimizer().allocate(literals["rp_mu"], cov)
    assert np.abs(w - literals["rp_w"]).max() < 1e-3
    assert_allclose(_risk_contributions(w, cov), np.full(4, 0.25), atol=1e-10)
    wb = mb.RiskParityOptimizer(literals["rp_budget"]).allocate(literals["rp_mu"], cov)
    assert np.abs(wb - literals["rp_budget_w"]).max() < 1e-3          # the reference's own tolerance (decimal=3)
    assert_allclose(_risk_contributions(wb, cov), literals["rp_budget"], atol=1e-10)
    assert mb.RiskParityOptimizer().name == "Risk Parity"


def test_risk_parity_returns_budget_where_slsqp_stops_at_once():
    """on the synthetic truth the reference returns exactly 1/N (SLSQP nit = 1)."""
    mu, cov = block_diagonal_truth(100)
    rng = np.random.RandomState(0)
    covs = [cov] + [np.cov(rng.multivariate_normal(mu, cov, size=200), rowvar=False) for _ in range(3)]
    eng = Engine(100, 200)
    w, info = eng.risk_parity(np.stack(covs))
    for i, c in enumerate(covs):
        ref = oracle.risk_parity(c)
        assert np.array_equal(ref, np.full(100, 0.01))                 # what the reference does
        assert np.array_equal(w[i], ref)
        assert info["iters"][i] == 0


@pytest.mark.parametrize("n", [4, 30, 200])
def test_exact_risk_budgeting(n):
    rng = np.random.RandomState(n)
    a = rng.standard_normal((n, 2 * n))
    cov = a @ a.T / (2 * n) * 4.0 + 0.5                                 # O(1) scale: SLSQP would iterate
    budget = rng.dirichlet(np.ones(n) * 5)
    eng = Engine(n, 10)
    w, info = eng.risk_parity(np.stack([cov, cov * 3.0]), budget)
    assert (info["status"] == 0).all() and (info["iters"] > 0).all()
    for i in range(2):
        assert abs(w[i].sum() - 1) < 1e-12 and (w[i] > 0).all()
        assert_allclose(_risk_contributions(w[i], cov), budget, atol=1e-10)
    ref = oracle.risk_parity(cov, budget)                              # converged SLSQP, ~1e-3 accurate
    assert np.abs(w[0] - ref).max() < 5e-3


def test_risk_parity_in_the_fused_loop(stock):
    sim = mb.MuCovObservationSimulator(stock["mu"], stock["cov"], 60, seed=2)
    df = mb.simulate_optimizations(sim, 32, [mb.MarkowitzOptimizer(), mb.HRPOptimizer(), mb.RiskParityOptimizer()],
                                   mb.VarianceErrorEstimator(), [mb.DeNoiserCovarianceTransformer()])
    assert list(df.index) == ["markowitz", "HRP", "Risk Parity"] and np.isfinite(df.values).all()
