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


@pytest.mark.parametrize("n", [40, 200, 500])
def test_eigenvalues_on_clustered_and_graded_spectra(n):
    """The eigenvalue kernel brackets by Sturm counts and finishes with Newton steps once a bracket isolates one eigenvalue;
    clusters never isolate and stay on bisection.  Spectra with exact multiplicities, pairs 1e-9 and 1e-13 apart, a geometric
    grading over 10 decades and the sample correlation of the block truth must all come back within a few ulp of the
    spectrum's scale, in descending order."""
    rng = np.random.RandomState(n)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    spectra = []
    lam = np.ones(n); lam[: n // 4] = 3.0; lam[n // 4: n // 2] = 0.25                     # three exact clusters
    spectra.append(lam)
    lam = np.linspace(0.1, 2.0, n); lam[1::2] = lam[0::2][: len(lam[1::2])] + 1e-9         # close pairs
    lam[5] = lam[4] + 1e-13
    spectra.append(lam)
    spectra.append(np.geomspace(1e-10, 10.0, n))                                            # graded
    mats = []
    for lam in spectra:
        a = (q * lam) @ q.T
        a = 0.5 * (a + a.T)
        dg = np.sqrt(np.diag(a))
        mats.append(a / np.outer(dg, dg))         # unit diagonal: the kernel sees exactly this matrix as its correlation
    mu, cov = block_diagonal_truth(n)
    mats.append(np.corrcoef(rng.multivariate_normal(mu, cov, size=2 * n), rowvar=False))
    mats = np.stack(mats)
    eng = Engine(n, 2 * n)
    for method in (1, 2) if n >= 128 else (1,):
        eng.set_eig_method(method)
        _, info = eng.denoise(mats, 2 * n)
        for i in range(mats.shape[0]):
            ref = np.linalg.eigvalsh(mats[i])[::-1]
            got = info["eigvals"][i]
            assert np.all(np.diff(got) <= 0)
            assert_allclose(got, ref, rtol=0, atol=4e-14 * n * max(1.0, abs(ref[0])))


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
