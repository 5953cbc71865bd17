"""GPU parity: DeNoiser (K3) through mcosb_denoise vs the reference's goldens and the CPU oracle.
Tolerance (north_star): 1e-6 relative after de-noising, n_facts equal."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_almost_equal

import oracle
from conftest import run_cases
from mcos_b200.engine import Engine
from mcos_b200.synthetic import block_diagonal_truth

pytestmark = pytest.mark.gpu


def _check(cov_in, t, out, info, i, want=None):
    d = oracle.denoise_details(cov_in, t) if want is None else want
    assert int(info["n_facts"][i]) == int(d["n_facts"])
    assert_allclose(info["eigvals"][i], d["eigvals"], rtol=0, atol=1e-11 * max(1.0, abs(d["eigvals"]).max()))
    scale = np.abs(d["cov"]).max()
    assert_allclose(out[i], d["cov"], rtol=1e-6, atol=1e-9 * scale)
    assert info["status"][i] == 0


def test_denoise_golden(stock, literals):
    """reference tests/test_covariance_transformer.py:46-52 (n_observations = prices.size, fit runs into the bound)."""
    eng = Engine(20, 50)
    out, info = eng.denoise(stock["cov"][None], int(literals["denoise_n_observations"]))
    assert_almost_equal(out[0], literals["denoised_cov"])
    assert info["n_facts"][0] == 3
    assert abs(info["var"][0] - 0.99999) < 1e-12


@pytest.mark.parametrize("case", ["stock20_T50", "block100_T200", "block60_T30"])
def test_denoise_vs_reference_runs(runs, case):
    keys = run_cases(runs, case)
    t = int(case.split("_T")[1])
    cov = np.stack([runs[f"{k}/denoise_in"] for k in keys])
    eng = Engine(cov.shape[1], t)
    out, info = eng.denoise(cov, t)
    for i, k in enumerate(keys):
        assert int(info["n_facts"][i]) == int(runs[f"{k}/n_facts"])
        assert_allclose(out[i], runs[f"{k}/denoised"], rtol=1e-6, atol=1e-9 * np.abs(runs[f"{k}/denoised"]).max())
        assert_allclose(info["eigvals"][i], runs[f"{k}/eigvals"], rtol=0, atol=1e-11)
        lam_plus = info["var"][i] * (1 + (cov.shape[1] / t) ** 0.5) ** 2
        assert abs(lam_plus - float(runs[f"{k}/lambda_plus"])) < 1e-5 * lam_plus


@pytest.mark.parametrize("n,t,nsim", [(500, 1000, 3), (200, 300, 4), (33, 100, 5), (3, 10, 2), (2, 10, 2), (130, 520, 2)])
def test_denoise_vs_oracle(n, t, nsim):
    rng = np.random.RandomState(n + t)
    if n % 10 == 0:
        mu, cov = block_diagonal_truth(n)
    else:
        a = rng.standard_normal((n, 3 * n))
        cov = a @ a.T / (3 * n) * np.outer(*(2 * [rng.uniform(0.05, 0.2, n)]))
        mu = np.zeros(n)
    covs = np.stack([np.cov(rng.multivariate_normal(mu, cov, size=t), rowvar=False).reshape(n, n) for _ in range(nsim)])
    eng = Engine(n, t)
    out, info = eng.denoise(covs, t)
    for i in range(nsim):
        _check(covs[i], t, out, info, i)


def test_denoise_identity_and_constant_inputs():
    """Degenerate spectra: identity (all eigenvalues equal) and a rank-one-plus-identity matrix."""
    n, t = 24, 100
    eng = Engine(n, t)
    eye = np.eye(n) * 0.04
    one = 0.04 * (0.5 * np.ones((n, n)) + 0.5 * np.eye(n))
    out, info = eng.denoise(np.stack([eye, one]), t)
    for i, c in enumerate((eye, one)):
        d = oracle.denoise_details(c, t)
        assert int(info["n_facts"][i]) == d["n_facts"]
        assert_allclose(out[i], d["cov"], rtol=1e-6, atol=1e-12)


def test_denoise_many_signal_eigenvalues_and_mixed_batch():
    """More than 16 signal eigenvalues takes the one-CTA-per-matrix eigenvector kernel, fewer the wide kernels; a batch
    that mixes both must give every simulation the right answer."""
    n, t = 120, 1200
    rng = np.random.RandomState(4)
    load = np.zeros((n, 30))
    for f in range(30):
        load[4 * f:4 * f + 4, f] = 3.0                         # 30 blocks of 4 tightly tied assets -> n_facts = 30
    strong = load @ load.T + np.eye(n)
    mu, weak = block_diagonal_truth(120, n_blocks=4)           # 4 factors
    weak = weak / weak.max()
    covs = []
    for base in (strong, weak, strong, weak, weak):
        covs.append(np.cov(rng.multivariate_normal(np.zeros(n), base, size=t), rowvar=False))
    covs = np.stack(covs)
    eng = Engine(n, t)
    out, info = eng.denoise(covs, t)
    assert info["n_facts"].max() > 16 and info["n_facts"].min() <= 16
    for i in range(len(covs)):
        _check(covs[i], t, out, info, i)
