"""Pins the CPU oracle (oracle/) against the reference: (i) the golden vectors the reference's own tests hold
(tests/golden/reference_literals.npz, parsed from the reference's test files) and (ii) outputs of the reference's
classes on injected draws (tests/golden/reference_runs.npz, made by tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_almost_equal, assert_array_almost_equal

import oracle
from conftest import run_cases


def test_price_history_helpers(stock, literals):
    import pandas as pd
    prices = pd.DataFrame(stock["prices"], columns=[str(t) for t in stock["tickers"]],
                          index=pd.to_datetime(stock["dates"]))
    mu, cov = oracle.convert_price_history(prices)
    tick = [str(t) for t in stock["tickers"]]
    got = [mu[tick.index("GOOG")], mu[tick.index("AAPL")], cov[tick.index("GOOG"), tick.index("AAPL")],
           cov[tick.index("GE"), tick.index("FB")]]
    assert_allclose(got, literals["utils_values"], rtol=1e-12)        # reference tests/test_utils.py:9-12
    assert_allclose(mu, stock["mu"], rtol=0, atol=0)
    assert_allclose(cov, stock["cov"], rtol=0, atol=0)


def test_markowitz_golden(stock, literals):
    # reference tests/test_optimizer.py:11-22
    w = oracle.markowitz_reference(stock["mu"], stock["cov"])
    assert_array_almost_equal(w, literals["markowitz_w"])
    assert np.abs(w - literals["markowitz_w"]).max() == 0.0
    # the exact QP agrees with SLSQP's answer to SLSQP's own accuracy and has the same support
    we = oracle.markowitz_exact_qp(stock["mu"], stock["cov"])[0]
    assert np.abs(we - w).max() < 2e-4
    assert ((we > 1e-4) == (w > 0)).all()


def test_hrp_golden(stock, literals):
    # reference tests/test_optimizer.py:55-64
    assert_almost_equal(oracle.hrp(stock["cov"]), literals["hrp_w"])
    d = oracle.hrp_details(stock["cov"], use_scipy=False)
    assert_almost_equal(d["weights"], literals["hrp_w"])
    assert list(d["order"]) == [15, 12, 4, 17, 19, 9, 14, 7, 10, 11, 13, 0, 5, 16, 8, 18, 1, 6, 2, 3]


def test_nco_algebra_golden(stock, literals):
    # reference tests/test_optimizer.py:30-47, with the partition that sklearn 0.22.1 found
    labels = np.zeros(20, dtype=int)
    for cid, members in enumerate(([0, 5, 7, 8, 10, 11, 13, 16, 17, 18, 19], [1, 4, 6, 15], [2, 3], [9, 12, 14])):
        labels[members] = cid
    assert_array_almost_equal(oracle.nco(stock["mu"], stock["cov"], labels=labels), literals["nco_max_sharpe_w"])
    assert_array_almost_equal(oracle.nco(None, stock["cov"], labels=labels), literals["nco_min_var_w"])


def test_nco_clustering_runs(stock):
    # parity unpinned (sklearn version); the restated search must at least return a valid partition and weights
    w = oracle.nco(stock["mu"], stock["cov"])
    assert w.shape == (20,) and abs(w.sum() - 1.0) < 1e-9


def test_denoise_golden(stock, literals):
    # reference tests/test_covariance_transformer.py:46-52, 67-148
    d = oracle.denoise_details(stock["cov"], int(literals["denoise_n_observations"]))
    assert d["cov"].shape == (20, 20)
    assert_almost_equal(d["cov"], literals["denoised_cov"])
    assert d["n_facts"] == 3


def test_detone_golden(literals):
    # reference tests/test_covariance_transformer.py:10-43
    sd = np.diag(literals["detone_std"])
    cov = sd @ literals["detone_corr"] @ sd.T
    assert_array_almost_equal(oracle.detone(cov, 1), literals["detone_golden"])
    assert_array_almost_equal(oracle.detone(cov, 0), cov)


def test_estimators_golden(stock, literals):
    # reference tests/test_error_estimator.py:12-37
    a, o = literals["est_allocation"], literals["est_optimal_allocation"]
    got = [oracle.err_expected_outcome(stock["mu"], stock["cov"], a, o), oracle.err_variance(stock["mu"], stock["cov"], a, o),
           oracle.err_sharpe(stock["mu"], stock["cov"], a, o)]
    assert_almost_equal(got, literals["est_expected"])


def test_risk_parity_golden(literals):
    # reference tests/test_optimizer.py:70-89
    assert_almost_equal(oracle.risk_parity(literals["rp_cov"]), literals["rp_w"])
    assert_almost_equal(oracle.risk_parity(literals["rp_cov"], literals["rp_budget"]), literals["rp_budget_w"], decimal=3)


@pytest.mark.parametrize("case", ["stock20_T50", "block100_T200", "block60_T30"])
def test_against_reference_runs(runs, case):
    """Oracle functions vs the reference's classes on the same injected draws."""
    t = int(case.split("_T")[1])
    for key in run_cases(runs, case):
        x = runs[f"{key}/x"]
        mu_hat, cov_hat = oracle.moments_mucov(x)
        assert_allclose(mu_hat, runs[f"{key}/mucov_mu"], rtol=0, atol=0)
        assert_allclose(cov_hat, runs[f"{key}/mucov_cov"], rtol=0, atol=0)
        _, lw = oracle.moments_ledoit_wolf(x)
        assert_allclose(lw, runs[f"{key}/lw_cov"], rtol=0, atol=0)
        assert_allclose(oracle.ledoit_wolf_formula(x)[0], lw, rtol=1e-11, atol=1e-16)
        jm, jc = oracle.moments_jackknife(x)
        assert_allclose(jc, runs[f"{key}/jk_cov"], rtol=1e-12, atol=1e-18)
        assert_allclose(jm, runs[f"{key}/jk_mu"], rtol=1e-12, atol=1e-18)
        # the jackknife estimator is algebraically the plain one
        assert_allclose(jc, cov_hat, rtol=1e-9, atol=1e-15)
        base = runs[f"{key}/denoise_in"]
        d = oracle.denoise_details(base, t)
        assert d["n_facts"] == int(runs[f"{key}/n_facts"])
        assert_allclose(d["cov"], runs[f"{key}/denoised"], rtol=1e-12, atol=1e-16)
        assert_allclose(d["eigvals"], runs[f"{key}/eigvals"], rtol=0, atol=1e-13)
        den = runs[f"{key}/denoised"]
        h = oracle.hrp_details(den)
        assert list(h["order"]) == list(runs[f"{key}/hrp_order"])
        assert_allclose(h["weights"], runs[f"{key}/hrp_w"], rtol=1e-13)
        h2 = oracle.hrp_details(den, use_scipy=False)       # restated linkage == scipy's, bit for bit
        assert np.array_equal(h2["linkage"], runs[f"{key}/hrp_link"])
        w = oracle.markowitz_reference(runs[f"{key}/mucov_mu"], den)
        assert_allclose(w, runs[f"{key}/markowitz_w"], rtol=0, atol=0)
        we = oracle.markowitz_exact_qp(runs[f"{key}/mucov_mu"], den)[0]
        assert np.abs(we - w).max() < 3e-4


def test_exact_qp_kkt():
    """The exact-QP oracle satisfies the KKT conditions of the long-only max-Sharpe problem to 1e-10."""
    from mcos_b200.synthetic import block_diagonal_truth
    mu, cov = block_diagonal_truth(100)
    w, info = oracle.markowitz_exact_qp(mu, cov)
    assert abs(w.sum() - 1) < 1e-12 and (w >= 0).all()
    c = mu - 0.02
    grad = cov @ w
    ratio = grad[w > 0] / c[w > 0]                       # equal on the support
    assert np.ptp(ratio) < 1e-10 * abs(ratio).max()
    assert (grad[w == 0] - ratio.mean() * c[w == 0] >= -1e-10).all()
