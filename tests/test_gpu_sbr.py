"""GPU: the two-stage tridiagonalisation (sy2sb + sb2st) behind np.linalg.eigh (covariance_transformer.py:105).
Stage 1 alone (band matrix similar to the input), both stages (tridiagonal similar to the input), and the whole DeNoiser /
Detone with the one-stage and the two-stage method against each other and the oracle."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

import oracle
from mcos_b200.engine import Engine

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def _corr_batch(n, s, seed):
    rng = np.random.RandomState(seed)
    out = []
    for k in range(s):
        t = 2 * n if k % 2 == 0 else max(2, n // 2)        # full rank and rank deficient
        out.append(np.corrcoef(rng.standard_normal((t, n)) + 0.3 * rng.standard_normal((t, 1)), rowvar=False))
    return np.stack(out)


def _tri_eigs(d, e):
    n = len(d)
    t = np.diag(d) + np.diag(e[:n - 1], 1) + np.diag(e[:n - 1], -1)
    return np.linalg.eigvalsh(t)


@pytest.mark.parametrize("n", [10, 11, 16, 17, 18, 25, 33, 64, 65, 96, 100, 129, 200, 255, 256, 257, 333, 500, 511, 512, 513, 520, 690, 696,
                               768, 777, 1000, 1023, 1024])
def test_two_stage_tridiagonal_is_similar(n):
    a = _corr_batch(n, 3, n)
    eng = Engine(n, 2 * n)
    eng.set_eig_method(2)
    d, e, work = eng.tridiagonalize(a, want_work=True)
    for i in range(a.shape[0]):
        ref = np.linalg.eigvalsh(a[i])
        band = np.tril(work[i]) - np.tril(work[i], -9)       # stage 1 leaves the band in the lower 9 diagonals
        band = band + np.tril(band, -1).T
        assert_allclose(np.linalg.eigvalsh(band), ref, rtol=0, atol=2e-13 * n)
        assert_allclose(_tri_eigs(d[i], e[i]), ref, rtol=0, atol=2e-13 * n)
        assert e[i, n - 1] == 0.0


def test_one_stage_still_available():
    n = 130
    a = _corr_batch(n, 2, 5)
    eng = Engine(n, 2 * n)
    eng.set_eig_method(1)
    d, e = eng.tridiagonalize(a)
    for i in range(2):
        assert_allclose(_tri_eigs(d[i], e[i]), np.linalg.eigvalsh(a[i]), rtol=0, atol=1e-12 * n)


@pytest.mark.parametrize("n,t", [(100, 200), (130, 520), (500, 1000), (60, 30), (333, 700), (1000, 2000)])
def test_denoise_two_stage_matches_one_stage_and_oracle(n, t):
    rng = np.random.RandomState(n)
    a = rng.standard_normal((n, 8)) * (np.arange(8) + 1.0)
    cov_true = a @ a.T + np.diag(rng.uniform(0.5, 2.0, n))
    covs = np.stack([np.cov(rng.multivariate_normal(np.zeros(n), cov_true, size=t), rowvar=False) for _ in range(3)])
    eng = Engine(n, t)
    eng.set_eig_method(1)
    out1, info1 = eng.denoise(covs, t)
    eng.set_eig_method(2)
    out2, info2 = eng.denoise(covs, t)
    assert (info1["status"] == 0).all() and (info2["status"] == 0).all()
    assert np.array_equal(info1["n_facts"], info2["n_facts"])
    assert_allclose(info2["eigvals"], info1["eigvals"], rtol=0, atol=1e-11 * np.abs(info1["eigvals"]).max())
    assert_allclose(out2, out1, rtol=1e-8, atol=1e-11 * np.abs(out1).max())
    d = oracle.denoise_details(covs[0], t)
    assert int(info2["n_facts"][0]) == int(d["n_facts"])
    assert_allclose(out2[0], d["cov"], rtol=1e-6, atol=1e-9 * np.abs(d["cov"]).max())


def test_many_signal_eigenvalues_take_the_one_cta_path():
    """More than 16 signal eigenvalues: eigvec_kernel applies Q2 and the panel reflectors itself."""
    n, t, f = 120, 600, 24
    rng = np.random.RandomState(3)
    load = rng.standard_normal((n, f)) * 3.0
    cov_true = load @ load.T + np.eye(n)
    covs = np.stack([np.cov(rng.multivariate_normal(np.zeros(n), cov_true, size=t), rowvar=False) for _ in range(2)])
    eng = Engine(n, t)
    eng.set_eig_method(2)
    out, info = eng.denoise(covs, t)
    assert (info["n_facts"] > 16).all()
    for i in range(2):
        d = oracle.denoise_details(covs[i], t)
        assert int(info["n_facts"][i]) == int(d["n_facts"])
        assert_allclose(out[i], d["cov"], rtol=1e-6, atol=1e-9 * np.abs(d["cov"]).max())


def test_detone_two_stage():
    n = 150
    a = _corr_batch(n, 2, 9) * 0.04
    eng = Engine(n, 2 * n)
    eng.set_eig_method(1)
    o1, _ = eng.detone(a, 2)
    eng.set_eig_method(2)
    o2, _ = eng.detone(a, 2)
    assert_allclose(o2, o1, rtol=1e-8, atol=1e-12)


def test_rank_deficient_and_nan_inputs_on_the_two_stage_path():
    """T < N (hundreds of exactly repeated zero eigenvalues) and a NaN-poisoned matrix in the same batch: no hang, the clean
    simulations are untouched, the poisoned one is flagged."""
    n, t = 160, 40
    rng = np.random.RandomState(11)
    covs = np.stack([np.cov(rng.standard_normal((t, n)) * rng.uniform(0.05, 0.2, n), rowvar=False) for _ in range(4)])
    eng = Engine(n, t)
    eng.set_eig_method(2)
    corr = np.stack([c / np.sqrt(np.outer(np.diag(c), np.diag(c))) for c in covs])
    d, e = eng.tridiagonalize(corr)
    for i in range(4):
        assert_allclose(_tri_eigs(d[i], e[i]), np.linalg.eigvalsh(corr[i]), rtol=0, atol=1e-12 * n)
    good, info_good = eng.denoise(covs, t)
    assert (info_good["status"] == 0).all()
    for i in range(4):
        ref = oracle.denoise_details(covs[i], t)
        assert int(info_good["n_facts"][i]) == int(ref["n_facts"])
        assert_allclose(good[i], ref["cov"], rtol=1e-6, atol=1e-9 * np.abs(ref["cov"]).max())
    bad = covs.copy()
    bad[1, 3, 5] = bad[1, 5, 3] = np.nan
    out, info = eng.denoise(bad, t)
    assert info["status"][1] != 0 and (info["status"][[0, 2, 3]] == 0).all()
    assert np.array_equal(out[[0, 2, 3]], good[[0, 2, 3]])


def test_tensor_core_bulge_chasing_variant():
    """sb2st_mma_kernel (MCOSB_SB2ST_MMA=1): the reflectors applied as 8 x 8 GEMMs on the FP64 tensor core."""
    import subprocess, sys, os, textwrap
    code = textwrap.dedent("""
        import numpy as np
        from mcos_b200.engine import Engine
        n = 300
        rng = np.random.RandomState(4)
        a = np.stack([np.corrcoef(rng.standard_normal((2 * n, n)), rowvar=False) for _ in range(3)])
        eng = Engine(n, 2 * n)
        eng.set_eig_method(2)
        d, e = eng.tridiagonalize(a)
        for i in range(3):
            t = np.diag(d[i]) + np.diag(e[i][:n - 1], 1) + np.diag(e[i][:n - 1], -1)
            err = np.abs(np.linalg.eigvalsh(t) - np.linalg.eigvalsh(a[i])).max()
            assert err < 1e-12 * n, err
        print("ok")
    """)
    env = dict(os.environ, MCOSB_SB2ST_MMA="1", PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_automatic_method_outside_the_two_stage_range():
    """N = 1030 > 1024 (shared memory of stage 1): automatic selection falls back to the one-stage kernel, forcing the
    two-stage method is an argument error -- never a silent wrong answer."""
    from mcos_b200._lib import MCOSBError
    n = 1030
    a = _corr_batch(n, 2, 3)
    eng = Engine(n, 2 * n)
    d, e = eng.tridiagonalize(a)
    for i in range(2):
        assert_allclose(_tri_eigs(d[i], e[i]), np.linalg.eigvalsh(a[i]), rtol=0, atol=2e-13 * n)
    eng.set_eig_method(2)
    with pytest.raises(MCOSBError):
        eng.tridiagonalize(a)
    with pytest.raises(MCOSBError):
        eng.set_eig_method(3)


@pytest.mark.parametrize("kind", ["1", "2", "3"])
def test_all_stage1_kernels_over_their_size_ranges(kind):
    """The automatic choice is the batched dense -> band reduction (sy2sb3.cu: panel kernel + tile pass over all matrices);
    MCOSB_STAGE1 forces the one-CTA-per-matrix kernels of round 1 (sy2sb.cu, N <= 696) and of round 2's first session
    (sy2sb2.cu) wherever they apply.  All must give a band similar to the input over their whole range (odd sizes take the
    8-byte copy path of the batched pass; tile rows that overhang N and the diagonal tiles take its predicated steps)."""
    import subprocess, sys, os, textwrap
    sizes = [10, 11, 17, 64, 65, 73, 129, 255, 256, 257, 333, 440, 441, 500, 512, 513, 600, 640, 641, 696] + ([] if kind == "1" else [777, 1000, 1023])
    code = textwrap.dedent(f"""
        import numpy as np
        from mcos_b200.engine import Engine
        for n in {sizes}:
            rng = np.random.RandomState(n)
            a = np.stack([np.corrcoef(rng.standard_normal((tt, n)), rowvar=False) for tt in (2 * n, max(2, n // 2))])
            eng = Engine(n, 2 * n)
            eng.set_eig_method(2)
            d, e, work = eng.tridiagonalize(a, want_work=True)
            for i in range(2):
                ref = np.linalg.eigvalsh(a[i])
                band = np.tril(work[i]) - np.tril(work[i], -9)
                band = band + np.tril(band, -1).T
                t = np.diag(d[i]) + np.diag(e[i][:n - 1], 1) + np.diag(e[i][:n - 1], -1)
                assert np.abs(np.linalg.eigvalsh(band) - ref).max() < 2e-13 * n, (n, "band")
                assert np.abs(np.linalg.eigvalsh(t) - ref).max() < 2e-13 * n, (n, "tridiagonal")
            eng.close()
        print("ok")
    """)
    env = dict(os.environ, MCOSB_STAGE1=kind, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_batched_stage1_is_reproducible_and_batch_independent():
    """sy2sb3 (panel kernel + tile pass over all matrices): the partial sums of Z have one writer per slot and are added in
    slot order, so a matrix gives the same bits whatever else is in the batch and however often it is run."""
    n = 200
    a = _corr_batch(n, 4, 21)
    eng = Engine(n, 2 * n)
    eng.set_eig_method(2)
    d0, e0, w0 = eng.tridiagonalize(a, want_work=True)
    d1, e1, w1 = eng.tridiagonalize(a, want_work=True)
    assert np.array_equal(d0, d1) and np.array_equal(e0, e1) and np.array_equal(np.tril(w0), np.tril(w1))
    big = np.concatenate([a[2:3]] * 37 + [a[0:1]] * 3)          # other neighbours, another batch size
    d2, e2 = eng.tridiagonalize(big)
    assert np.array_equal(d2[0], d0[2]) and np.array_equal(e2[0], e0[2])
    assert np.array_equal(d2[36], d0[2]) and np.array_equal(d2[39], d0[0]) and np.array_equal(e2[39], e0[0])


def test_batched_stage1_more_matrices_than_one_grid_dimension():
    """More than 32768 matrices in one call: the pass indexes the matrix by blockIdx.y and is launched in chunks."""
    n, s = 12, 33000
    rng = np.random.RandomState(5)
    base = np.stack([np.corrcoef(rng.standard_normal((3 * n, n)), rowvar=False) for _ in range(3)])
    a = base[np.arange(s) % 3]
    eng = Engine(n, 2 * n)
    eng.set_eig_method(2)
    d, e = eng.tridiagonalize(a)
    for i in (0, 1, 2, 32767, 32768, 32999):
        assert_allclose(_tri_eigs(d[i], e[i]), np.linalg.eigvalsh(a[i]), rtol=0, atol=1e-12 * n)
        assert np.array_equal(d[i], d[i % 3]) and np.array_equal(e[i], e[i % 3])


def test_compact_wy_back_transformation_matches_the_reflectorwise_one():
    """The signal eigenvectors through backtransform_wy_kernel (one DMMA reduction per panel of 8 reflectors) against the
    reflector-by-reflector kernel (MCOSB_BACKTRANSFORM_WY=0), at a size per launch configuration and with the shapes of the
    bench; both against the oracle."""
    import subprocess, sys, os, textwrap
    code = textwrap.dedent("""
        import numpy as np, sys
        from mcos_b200.engine import Engine
        outs = []
        for n, t in ((60, 150), (130, 520), (257, 600), (500, 1000), (777, 1600)):
            rng = np.random.RandomState(n)
            a = rng.standard_normal((n, 6)) * (np.arange(6) + 1.0)
            cov_true = a @ a.T + np.diag(rng.uniform(0.5, 2.0, n))
            covs = np.stack([np.cov(rng.multivariate_normal(np.zeros(n), cov_true, size=t), rowvar=False) for _ in range(2)])
            eng = Engine(n, t)
            eng.set_eig_method(2)
            out, info = eng.denoise(covs, t)
            assert (info["status"] == 0).all() and (info["n_facts"] > 0).all()
            outs.append(out)
            eng.close()
        np.savez(sys.argv[1], *outs)
        print("ok")
    """)
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    with tempfile.TemporaryDirectory() as tmp:
        for wy in ("1", "0"):
            path = os.path.join(tmp, f"wy{wy}.npz")
            env = dict(os.environ, MCOSB_BACKTRANSFORM_WY=wy, PYTHONPATH=root)
            out = subprocess.run([sys.executable, "-c", code, path], env=env, capture_output=True, text=True, timeout=600)
            assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]
            with np.load(path) as z:
                res[wy] = [z[k] for k in z.files]
    for a, b in zip(res["1"], res["0"]):
        assert_allclose(a, b, rtol=1e-11, atol=1e-13 * np.abs(b).max())
