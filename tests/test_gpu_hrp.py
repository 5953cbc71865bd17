"""GPU parity: HRP (K4b-d) through mcosb_hrp.  Bars: cluster order and linkage bit-exact vs scipy, weights 1e-9."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_almost_equal

import oracle
from conftest import run_cases
from mcos_b200 import _lib
from mcos_b200.engine import Engine
from mcos_b200.synthetic import block_diagonal_truth

pytestmark = pytest.mark.gpu


def _check(cov, w, info, i):
    d = oracle.hrp_details(cov)
    assert np.array_equal(info["order"][i], d["order"])
    assert np.array_equal(info["linkage"][i], d["linkage"])            # bit-exact, distances included
    assert_allclose(w[i], d["weights"], rtol=1e-9, atol=1e-15)
    assert info["status"][i] == 0


def test_hrp_golden(stock, literals):
    """reference tests/test_optimizer.py:55-64 (weights in quasi-diagonal position order)."""
    eng = Engine(20, 50)
    w, info = eng.hrp(stock["cov"][None])
    assert_almost_equal(w[0], literals["hrp_w"])
    assert list(info["order"][0]) == [15, 12, 4, 17, 19, 9, 14, 7, 10, 11, 13, 0, 5, 16, 8, 18, 1, 6, 2, 3]
    _check(stock["cov"], w, info, 0)


@pytest.mark.parametrize("case", ["stock20_T50", "block100_T200", "block60_T30"])
def test_hrp_vs_reference_runs(runs, case):
    keys = run_cases(runs, case)
    cov = np.stack([runs[f"{k}/denoised"] for k in keys])
    eng = Engine(cov.shape[1], 10)
    w, info = eng.hrp(cov)
    for i, k in enumerate(keys):
        assert np.array_equal(info["order"][i], runs[f"{k}/hrp_order"])
        assert np.array_equal(info["linkage"][i], runs[f"{k}/hrp_link"])
        assert_allclose(w[i], runs[f"{k}/hrp_w"], rtol=1e-9, atol=1e-15)


@pytest.mark.parametrize("n,t,nsim", [(500, 1000, 3), (100, 200, 8), (65, 70, 4), (1, 5, 1), (2, 9, 2), (3, 9, 2)])
def test_hrp_vs_oracle(n, t, nsim):
    rng = np.random.RandomState(n * 3 + t)
    if n % 10 == 0:
        mu, cov = block_diagonal_truth(n)
    else:
        a = rng.standard_normal((n, 2 * n + 1))
        cov = a @ a.T / (2 * n + 1) * 0.01
        mu = np.zeros(n)
    covs = np.stack([np.cov(rng.multivariate_normal(mu, cov, size=t), rowvar=False).reshape(n, n) for _ in range(nsim)])
    eng = Engine(n, t)
    w, info = eng.hrp(covs)
    for i in range(nsim):
        if n == 1:
            assert w[i, 0] == 1.0
        else:
            _check(covs[i], w, info, i)


def test_hrp_exact_ties():
    """The true block-diagonal covariance has massive exact ties in every distance: the merge order then depends on
    Prim's lowest-index rule and the stable sort, which must match scipy's."""
    for n, blocks in ((20, 2), (60, 6), (100, 10)):
        mu, cov = block_diagonal_truth(n, n_blocks=blocks)
        eng = Engine(n, 10)
        w, info = eng.hrp(cov[None])
        _check(cov, w, info, 0)
    # rounded inputs -> ties in the sample covariance too
    rng = np.random.RandomState(0)
    x = np.round(rng.standard_normal((50, 30)), 1)
    cov = np.cov(x, rowvar=False)
    eng = Engine(30, 50)
    w, info = eng.hrp(cov[None])
    _check(cov, w, info, 0)


def test_hrp_fast_path_equals_exact_path_bit_for_bit():
    """Fast path (tensor-core distances + guarded Prim + exact merge heights, mode 0) against exact distances for every
    simulation (mode 1) and against the fast pass with every simulation forced through the fallback (mode 2): identical
    linkage, order and weights, and on sampled covariances almost no simulation needs the fallback."""
    for n, t, nsim in ((500, 1000, 96), (100, 200, 300), (37, 80, 200), (129, 300, 64)):
        rng = np.random.RandomState(n + 7)
        if n % 10 == 0:
            mu, cov = block_diagonal_truth(n)
        else:
            a = rng.standard_normal((n, 2 * n + 1))
            cov = a @ a.T / (2 * n + 1) * 0.01
            mu = np.zeros(n)
        eng = Engine(n, t, seed=5)
        eng.set_truth(mu, cov)
        x = eng.sample(nsim)
        _, covs, _ = eng.moments(x, _lib.MOMENTS_MUCOV)
        res = {}
        for mode in (0, 1, 2, 3):
            eng.set_hrp_mode(mode)
            eng.hrp_fallback_count()
            w, info = eng.hrp(covs)
            res[mode] = (w, info, eng.hrp_fallback_count())
        eng.set_hrp_mode(0)
        assert res[1][2] == 0 and res[2][2] == nsim
        assert res[0][2] <= max(1, nsim // 20), f"{res[0][2]} of {nsim} simulations took the fallback at n={n}"
        assert res[3][2] == res[0][2]          # mode 3 = the monolithic fast kernel (n_assets > 1024): same decisions
        for mode in (0, 2, 3):
            assert np.array_equal(res[mode][1]["order"], res[1][1]["order"])
            assert np.array_equal(res[mode][1]["linkage"], res[1][1]["linkage"])
            assert np.array_equal(res[mode][0], res[1][0])
            assert (res[mode][1]["status"] == 0).all()
        _check(covs[0], res[0][0], res[0][1], 0)


def test_hrp_ties_take_the_fallback():
    """Exact ties (the true block-diagonal covariance) are inside the error bound by definition: the fast pass must hand
    them to the exact path rather than decide."""
    mu, cov = block_diagonal_truth(100)
    eng = Engine(100, 10)
    eng.hrp_fallback_count()
    w, info = eng.hrp(np.stack([cov, cov]))
    assert eng.hrp_fallback_count() == 2
    _check(cov, w, info, 1)


def test_hrp_near_ties_on_the_kernel_take_the_fallback_and_match_scipy():
    """Near-ties, not exact ties: one covariance entry of the tie-ridden block-diagonal truth is moved by 1..4 ulp, so
    distances that were equal now differ in the last bits -- far inside the guard's error bound.  The fast pass must
    flag every such simulation (it cannot know which way scipy's rounded distances fall) and the exact path must then
    reproduce scipy's linkage bit for bit.  (tests/test_hrp_guard_cpu.py checks the same property on the numpy
    restatement of the guard; this is the kernel.)"""
    n = 100
    mu, cov = block_diagonal_truth(n)
    corr = cov / np.sqrt(np.outer(np.diag(cov), np.diag(cov)))
    rng = np.random.RandomState(5)
    covs = []
    for k in range(12):
        c = cov.copy()
        while True:
            i, j = rng.randint(0, n, 2)
            if i != j and corr[i, j] > 0.25:          # same block
                break
        v = c[i, j]
        for _ in range(k % 4 + 1):
            v = np.nextafter(v, np.inf if k % 2 else -np.inf)
        c[i, j] = c[j, i] = v
        covs.append(c)
    covs = np.stack(covs)
    eng = Engine(n, 10)
    eng.hrp_fallback_count()
    w, info = eng.hrp(covs)
    assert eng.hrp_fallback_count() == len(covs)
    for i in range(len(covs)):
        _check(covs[i], w, info, i)
    # sampled covariances with ONE near-tie planted: two assets made exchangeable up to a few ulp
    t = 200
    x = rng.multivariate_normal(mu, cov, size=t)
    x[:, 7] = x[:, 3]                                     # asset 7 = asset 3: exact ties in every distance to them
    base = np.cov(x, rowvar=False)
    planted = []
    for k in range(4):
        c = base.copy()
        v = c[7, 11]
        for _ in range(k + 1):
            v = np.nextafter(v, np.inf)
        c[7, 11] = c[11, 7] = v                           # d(7, .) and d(3, .) now differ by ulps
        planted.append(c)
    planted = np.stack(planted)
    w2, info2 = eng.hrp(planted)
    assert eng.hrp_fallback_count() == len(planted)
    for i in range(len(planted)):
        _check(planted[i], w2, info2, i)


def test_device_tensors_are_ordered_with_torch_work():
    """Device tensors in and out: nothing synchronises with the host, so the kernels must run on the stream the caller's
    torch work is on.  (torch's default stream has handle 0, which mcosb_set_stream reads as "own stream"; the engine
    passes cudaStreamLegacy instead.)  Exact distances for 256 matrices take milliseconds: a copy that does not wait
    for them returns uninitialised memory."""
    import torch
    n, t, nsim = 300, 600, 256
    mu, cov = block_diagonal_truth(n)
    eng = Engine(n, t, seed=9)
    eng.set_truth(mu, cov)
    x = eng.sample(nsim)
    _, covs, _ = eng.moments(x, _lib.MOMENTS_MUCOV)
    eng.set_hrp_mode(1)
    w_ref, info_ref = eng.hrp(covs)                                   # host buffers: staged and synchronised
    w, info = eng.hrp(torch.from_numpy(covs).cuda())                  # device tensors on torch's current (default) stream
    order, link, w = info["order"].cpu().numpy(), info["linkage"].cpu().numpy(), w.cpu().numpy()
    eng.set_hrp_mode(0)
    assert np.array_equal(order, info_ref["order"])
    assert np.array_equal(link, info_ref["linkage"])
    assert np.array_equal(w, w_ref)
