"""GPU parity: error estimators (K5), moments (K2) and the sampler (K1) through the C ABI vs the CPU oracle."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_almost_equal

import oracle
from conftest import run_cases
from mcos_b200 import _lib
from mcos_b200.engine import Engine
from mcos_b200.synthetic import block_diagonal_truth
from philox_ref import standard_normals

pytestmark = pytest.mark.gpu


def test_errors_golden(stock, literals):
    """reference tests/test_error_estimator.py:12-37 through mcosb_errors."""
    eng = Engine(20, 50)
    eng.set_truth(stock["mu"], stock["cov"])
    a, o = literals["est_allocation"], literals["est_optimal_allocation"]
    for kind, want in zip((_lib.ERR_EXPECTED_OUTCOME, _lib.ERR_VARIANCE, _lib.ERR_SHARPE), literals["est_expected"]):
        got = eng.errors(a[None, :], o, kind)
        assert_almost_equal(got[0], want)


@pytest.mark.parametrize("n,s", [(20, 1), (100, 37), (500, 19), (333, 9)])
def test_errors_vs_oracle(n, s):
    rng = np.random.RandomState(n + s)
    mu, cov = block_diagonal_truth(n if n % 10 == 0 else 330)[0][:n], None
    a = rng.standard_normal((n, n))
    cov = a @ a.T / n
    mu = rng.standard_normal(n)
    eng = Engine(n, 10)
    eng.set_truth(mu, cov)
    w = rng.dirichlet(np.ones(n), size=s)
    ws = rng.dirichlet(np.ones(n), size=s)
    for kind, name in ((0, "expected_outcome"), (1, "variance"), (2, "sharpe")):
        want_b = np.array([oracle.estimate_error(name, mu, cov, w[i], ws[i]) for i in range(s)])
        want_1 = np.array([oracle.estimate_error(name, mu, cov, w[i], ws[0]) for i in range(s)])
        assert_allclose(eng.errors(w, ws, kind), want_b, rtol=1e-9, atol=1e-18)
        assert_allclose(eng.errors(w, ws[0], kind), want_1, rtol=1e-9, atol=1e-18)


def test_errors_zero_delta_is_nan_for_sharpe():
    """error_estimator.py:41 gives nan when the allocation equals the ideal one."""
    mu, cov = block_diagonal_truth(20, n_blocks=2)
    eng = Engine(20, 10)
    eng.set_truth(mu, cov)
    w = np.full((2, 20), 0.05)
    assert np.isnan(eng.errors(w, w[0], _lib.ERR_SHARPE)).all()
    assert (eng.errors(w, w[0], _lib.ERR_VARIANCE) == 0).all()


@pytest.mark.parametrize("case", ["stock20_T50", "block100_T200", "block60_T30"])
def test_moments_vs_reference_runs(runs, case):
    """mcosb_moments on the draws the reference's simulators saw (tests/golden/reference_runs.npz)."""
    keys = run_cases(runs, case)
    x = np.stack([runs[f"{k}/x"] for k in keys])
    s, t, n = x.shape
    eng = Engine(n, t)
    mu, cov, sh = eng.moments(x, _lib.MOMENTS_MUCOV)
    for i, k in enumerate(keys):
        assert_allclose(mu[i], runs[f"{k}/mucov_mu"].ravel(), rtol=1e-12, atol=1e-16)
        assert_allclose(cov[i], runs[f"{k}/mucov_cov"], rtol=1e-12, atol=1e-17)
        assert np.array_equal(cov[i], cov[i].T)
    assert (sh == 0).all()
    mu, cov, sh = eng.moments(x, _lib.MOMENTS_LEDOIT_WOLF)
    for i, k in enumerate(keys):
        assert_allclose(mu[i], runs[f"{k}/lw_mu"].ravel(), rtol=1e-12, atol=1e-16)
        assert_allclose(cov[i], runs[f"{k}/lw_cov"], rtol=1e-10, atol=1e-16)
        assert_allclose(sh[i], oracle.ledoit_wolf_formula(x[i])[1], rtol=1e-9)
    mu, cov, sh = eng.moments(x, _lib.MOMENTS_JACKKNIFE)
    for i, k in enumerate(keys):
        assert_allclose(mu[i], runs[f"{k}/jk_mu"].ravel(), rtol=1e-10, atol=1e-15)
        assert_allclose(cov[i], runs[f"{k}/jk_cov"], rtol=1e-10, atol=1e-15)


@pytest.mark.parametrize("n,t,s", [(1, 2, 3), (7, 3, 2), (129, 17, 3), (500, 1000, 2), (257, 300, 2)])
def test_moments_ragged_shapes(n, t, s):
    """Tile-edge cases: N not a multiple of the 128 tile, T not a multiple of the k-slab, large mean offset."""
    rng = np.random.RandomState(n * 1000 + t)
    x = rng.standard_normal((s, t, n)) * rng.uniform(0.05, 2.0, n) + rng.uniform(-50, 50, n)
    eng = Engine(n, t)
    mu, cov, _ = eng.moments(x, _lib.MOMENTS_MUCOV)
    for i in range(s):
        assert_allclose(mu[i], x[i].mean(0), rtol=1e-13, atol=1e-13)
        want = np.cov(x[i], rowvar=False).reshape(n, n)
        assert_allclose(cov[i], want, rtol=1e-11, atol=1e-13 * np.abs(want).max())
    if t >= 2:
        _, lw, sh = eng.moments(x, _lib.MOMENTS_LEDOIT_WOLF)
        for i in range(s):
            want, want_sh = oracle.ledoit_wolf_formula(x[i])
            assert_allclose(lw[i], want, rtol=1e-9, atol=1e-12 * np.abs(want).max())


def test_moments_device_pointers():
    """Device-resident (torch) inputs give bit-identical results to host staging."""
    import torch
    rng = np.random.RandomState(5)
    x = rng.standard_normal((3, 64, 40))
    eng = Engine(40, 64)
    mu_h, cov_h, _ = eng.moments(x, _lib.MOMENTS_LEDOIT_WOLF)
    eng.use_torch_stream()
    xd = torch.from_numpy(x).cuda()
    mu_d, cov_d, _ = eng.moments(xd, _lib.MOMENTS_LEDOIT_WOLF)
    torch.cuda.synchronize()
    assert np.array_equal(mu_d.cpu().numpy(), mu_h) and np.array_equal(cov_d.cpu().numpy(), cov_h)


@pytest.mark.parametrize("n,t", [(20, 50), (101, 33), (500, 40)])
def test_sampler_matches_spec(stock, n, t):
    """mcosb_sample == mu + Z L' with Z from the numpy restatement of the Philox/Box-Muller spec."""
    if n == 20:
        mu, cov = stock["mu"], stock["cov"]
    else:
        rng = np.random.RandomState(n)
        a = rng.standard_normal((n, 2 * n))
        cov = a @ a.T / (2 * n) * 0.01
        mu = rng.standard_normal(n) * 0.1
    eng = Engine(n, t, seed=1234567890123)
    eng.set_truth(mu, cov)
    x = eng.sample(3, sim_id0=5)
    z = standard_normals([5, 6, 7], t, n, 1234567890123)
    want = mu + z @ np.linalg.cholesky(cov).T
    assert_allclose(x, want, rtol=1e-11, atol=1e-12 * np.abs(want).max())
    # counter-based: a different batch split gives the same simulations
    x2 = eng.sample(2, sim_id0=6)
    assert np.array_equal(x2, x[1:])


@pytest.mark.parametrize("n,blocks,t", [(500, 10, 24), (200, 4, 40), (130, 13, 17), (96, 1, 20)])
def test_sampler_sparse_cholesky_factor(n, blocks, t):
    """The TRMM visits only the 16-column slabs of L that hold a non-zero for its 64 rows (mcosb_set_truth lists them):
    block-diagonal truths -- the BASELINE configurations -- skip most of the triangle.  The draws must still be
    mu + Z L' for the spec's Z; an arrow-shaped truth (dense last rows, diagonal elsewhere) and a permuted
    block structure exercise slab lists that are not contiguous."""
    mu, cov = block_diagonal_truth(n, n_blocks=blocks)
    rng = np.random.RandomState(n)
    cases = [(mu, cov)]
    arrow = np.diag(rng.uniform(0.5, 1.5, n))
    arrow[:, -3:] = arrow[-3:, :] = 0.02
    arrow[np.arange(n), np.arange(n)] = rng.uniform(1.0, 2.0, n)
    cases.append((mu, arrow))
    perm = rng.permutation(n)
    cases.append((mu[perm], cov[np.ix_(perm, perm)]))
    for m, c in cases:
        eng = Engine(n, t, seed=99)
        eng.set_truth(m, c)
        x = eng.sample(2, sim_id0=3)
        z = standard_normals([3, 4], t, n, 99)
        want = m + z @ np.linalg.cholesky(c).T
        assert_allclose(x, want, rtol=1e-11, atol=1e-12 * np.abs(want).max())


def test_sampler_distribution():
    """Sample mean / covariance of many draws converge to the truth (distributional check of the GPU RNG)."""
    mu, cov = block_diagonal_truth(20, n_blocks=2)
    eng = Engine(20, 2000, seed=7)
    eng.set_truth(mu, cov)
    x = eng.sample(16).reshape(-1, 20)                 # 32000 draws
    se = np.sqrt(np.diag(cov) / x.shape[0])
    assert (np.abs(x.mean(0) - mu) < 5 * se).all()
    emp = np.cov(x, rowvar=False)
    assert np.abs(emp - cov).max() < 0.05 * np.abs(cov).max()
    from scipy import stats
    zs = (x - mu) / np.sqrt(np.diag(cov))
    for j in (0, 7, 19):
        assert stats.kstest(zs[:, j], "norm").pvalue > 1e-4


def test_not_pd_truth_is_reported():
    eng = Engine(3, 5)
    bad = np.array([[1.0, 2.0, 0.0], [2.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    with pytest.raises(_lib.MCOSBError):
        eng.set_truth(np.zeros(3), bad)
