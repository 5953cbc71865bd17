"""GPU: Markowitz with a free set larger than the first-tier capacity (64) is re-solved by the second tier."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

import oracle
from mcos_b200.engine import Engine

pytestmark = pytest.mark.gpu


def test_large_free_set_goes_through_the_second_tier():
    n = 200
    rng = np.random.RandomState(0)
    # nearly uncorrelated assets with similar Sharpe ratios: most of them end up in the portfolio
    sig = rng.uniform(0.1, 0.2, n)
    corr = np.full((n, n), 0.02) + 0.98 * np.eye(n)
    cov = corr * np.outer(sig, sig)
    mus = np.stack([np.where(np.arange(n) < 100, 0.02 + sig * rng.uniform(0.4, 0.6, n), 0.01) for _ in range(3)])
    covs = np.stack([cov] * 3)
    eng = Engine(n, 2 * n)
    w, info = eng.markowitz(mus, covs, clean=False)
    assert (info["status"] == 0).all()
    for i in range(3):
        ref, _ = oracle.markowitz_exact_qp(mus[i], cov, 0.02)
        assert 64 < (ref > 1e-9).sum() <= 128                     # the case really overflows the first tier
        assert_allclose(w[i], ref, rtol=0, atol=1e-9)
