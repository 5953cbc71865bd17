"""GPU parity: DetoneCovarianceTransformer (SURVEY 8f "next") through mcosb_detone vs the reference's golden and the
CPU oracle."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_almost_equal

import oracle
import mcos_b200 as mb
from mcos_b200 import _lib
from mcos_b200.engine import Engine
from mcos_b200.synthetic import block_diagonal_truth

pytestmark = pytest.mark.gpu


def test_detone_golden(literals):
    """reference tests/test_covariance_transformer.py:10-43."""
    sd = np.diag(literals["detone_std"])
    cov = sd @ literals["detone_corr"] @ sd.T
    out = mb.DetoneCovarianceTransformer(n_remove=1).transform(cov, None)
    assert_array_almost_equal(out, literals["detone_golden"])
    w = np.linalg.eigvalsh(out)
    assert abs(w[np.argmin(np.abs(w))]) < 1e-12                      # the removed direction has eigenvalue 0
    assert mb.DetoneCovarianceTransformer(n_remove=0).transform(cov, None) is cov


@pytest.mark.parametrize("n,t,k", [(20, 100, 1), (100, 200, 3), (500, 1000, 2), (64, 40, 16)])
def test_detone_vs_oracle(n, t, k):
    rng = np.random.RandomState(n + k)
    mu, cov = block_diagonal_truth(n if n % 10 == 0 else 60, n_blocks=10 if n % 10 == 0 and n >= 100 else 2)
    if cov.shape[0] != n:
        a = rng.standard_normal((n, 3 * n))
        cov = a @ a.T / (3 * n) * 0.02 + 0.01
        mu = np.zeros(n)
    covs = np.stack([np.cov(rng.multivariate_normal(mu, cov, size=max(t, 2 * n)), rowvar=False) for _ in range(4)])
    eng = Engine(n, t)
    out, info = eng.detone(covs, k)
    assert (info["status"] == 0).all()
    for i in range(len(covs)):
        want = np.real(oracle.detone(covs[i], k))
        assert_allclose(out[i], want, rtol=1e-8, atol=1e-11 * np.abs(want).max())


def test_detone_in_the_fused_loop():
    """[DeNoiser, Detone] in the plan == the two stage calls in sequence."""
    mu, cov = block_diagonal_truth(40, n_blocks=4)
    sim = mb.MuCovObservationSimulator(mu, cov, 120, seed=9)
    errs, status = mb.simulate_errors(sim, 10, [mb.HRPOptimizer()], mb.VarianceErrorEstimator(),
                                      [mb.DeNoiserCovarianceTransformer(), mb.DetoneCovarianceTransformer(1)])
    eng = Engine(40, 120, seed=9)
    eng.set_truth(mu, cov)
    _, c, _ = eng.sample_moments(10, 0, _lib.MOMENTS_MUCOV)     # sampling + moments as the fused loop runs them
    c, _ = eng.denoise(c, 120)
    c, _ = eng.detone(c, 1)
    w, _ = eng.hrp(c)
    ideal, _ = eng.hrp(cov[None])
    assert np.array_equal(errs[:, 0], eng.errors(w, ideal[0], _lib.ERR_VARIANCE))
