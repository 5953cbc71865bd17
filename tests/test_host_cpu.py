"""CPU-only tests: the C ABI library loads and exports every symbol include/mcosb.h declares, host-side logic
(plan compilation, simulation-id sharding, the world_size-2 gather over gloo), and the numpy restatement of the
one-variable L-BFGS-B that mpfit_kernel implements, checked against scipy iteration for iteration."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from mcos_b200 import _lib
    header = open(os.path.join(ROOT, "include", "mcosb.h")).read()
    declared = set(re.findall(r"\b(mcosb_[a-z_0-9]+)\s*\(", header))
    declared -= {"mcosb_ctx", "mcosb_plan"}
    assert {"mcosb_create", "mcosb_run", "mcosb_denoise", "mcosb_markowitz", "mcosb_hrp", "mcosb_nco", "mcosb_errors",
            "mcosb_sample", "mcosb_moments"} <= declared
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.run(["make", "-C", os.path.join(ROOT, "mcos_b200", "csrc")], check=True)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/mcosb.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    _lib.load()
    assert lib.mcosb_version() >= 100


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mcos_b200._lib import MCOSBError
    from mcos_b200.engine import Engine
    with pytest.raises(MCOSBError):
        Engine(10, 20)
    import mcos_b200 as mb
    with pytest.raises(MCOSBError):
        mb.HRPOptimizer().allocate(None, np.eye(4))


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mcos_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_shard_range_partitions_ids():
    from mcos_b200.mcos import shard_range
    for n, world in ((100000, 8), (10, 3), (7, 8), (0, 2), (5, 1)):
        spans = [shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_build_plan_maps_plugins_and_rejects_unknown_types():
    import ctypes
    import mcos_b200 as mb
    from mcos_b200 import _lib
    mu, cov = np.zeros(4), np.eye(4)
    sim = mb.MuCovLedoitWolfObservationSimulator(mu, cov, 10)
    plan = mb.build_plan(sim, [mb.MarkowitzOptimizer(), mb.HRPOptimizer(), mb.NCOOptimizer(5, 3)],
                         mb.SharpeRatioErrorEstimator(), [mb.DeNoiserCovarianceTransformer(0.3), mb.DetoneCovarianceTransformer(0)])
    assert plan.moments_kind == _lib.MOMENTS_LEDOIT_WOLF
    assert plan.n_transformers == 1 and plan.transformer_kind[0] == _lib.TR_DENOISE and plan.transformer_arg[0] == 0.3
    assert list(plan.optimizers)[:3] == [_lib.OPT_MARKOWITZ, _lib.OPT_HRP, _lib.OPT_NCO] and plan.n_optimizers == 3
    assert plan.error_kind == _lib.ERR_SHARPE and plan.n_error_kinds == 0
    assert plan.nco_max_clusters == 5 and plan.nco_trials == 3 and plan.risk_free_rate == 0.02 and not plan.risk_budget
    assert list(mb.build_plan(sim, [mb.RiskParityOptimizer()], mb.VarianceErrorEstimator(), []).optimizers)[0] == \
        _lib.OPT_RISK_PARITY
    # RiskParityOptimizer(target_risk) (optimizer.py:295-309): the budget travels with the plan
    budget = np.array([.1, .2, .3, .4])
    prp = mb.build_plan(sim, [mb.RiskParityOptimizer(budget)], mb.VarianceErrorEstimator(), [])
    got = np.ctypeslib.as_array(ctypes.cast(prp.risk_budget, ctypes.POINTER(ctypes.c_double)), shape=(4,))
    assert np.array_equal(got, budget)
    # any list of the built-in transformers, in order (mcos.py:25-26)
    plan2 = mb.build_plan(sim, [mb.HRPOptimizer()], mb.VarianceErrorEstimator(),
                          [mb.DetoneCovarianceTransformer(1), mb.DeNoiserCovarianceTransformer(), mb.DetoneCovarianceTransformer(2)])
    assert plan2.n_transformers == 3 and list(plan2.transformer_kind)[:3] == [_lib.TR_DETONE, _lib.TR_DENOISE, _lib.TR_DETONE]
    assert list(plan2.transformer_arg)[:3] == [1.0, 0.25, 2.0]
    # several estimators in one pass
    plan3 = mb.build_plan(sim, [mb.HRPOptimizer()],
                          [mb.ExpectedOutcomeErrorEstimator(), mb.VarianceErrorEstimator(), mb.SharpeRatioErrorEstimator()], [])
    assert plan3.n_error_kinds == 3 and list(plan3.error_kinds) == [_lib.ERR_EXPECTED_OUTCOME, _lib.ERR_VARIANCE, _lib.ERR_SHARPE]
    # the struct the header declares: 144 bytes, the pointer last
    assert ctypes.sizeof(_lib.Plan) == 144 and _lib.Plan.risk_budget.offset == 136

    # subclasses that keep the behaviour method stay on the GPU; user plugins and overriding subclasses have no GPU stage
    class Tuned(mb.HRPOptimizer):
        pass

    class Mine(mb.HRPOptimizer):
        def allocate(self, mu, cov):
            return np.full(len(cov), 1.0 / len(cov))

    assert mb.build_plan(sim, [Tuned()], mb.VarianceErrorEstimator(), []).optimizers[0] == _lib.OPT_HRP
    with pytest.raises(TypeError):
        mb.build_plan(sim, [Mine()], mb.VarianceErrorEstimator(), [])
    with pytest.raises(TypeError):
        mb.build_plan(object(), [mb.HRPOptimizer()], mb.VarianceErrorEstimator(), [])


def test_plan_struct_matches_the_header():
    """The ctypes Plan mirrors `struct mcosb_plan` of include/mcosb.h field for field (names and order)."""
    from mcos_b200 import _lib
    text = open(os.path.join(ROOT, "include", "mcosb.h")).read()
    body = text[text.index("typedef struct mcosb_plan {"):text.index("} mcosb_plan;")]
    import re
    names = re.findall(r"^\s*(?:const\s+)?(?:int|double)\*?\s+(\w+)(?:\[[^\]]+\])?;", body, flags=re.M)
    assert names == [f[0] for f in _lib.Plan._fields_], names


def test_reference_api_surface():
    """Same names and constructor signatures as the reference's plugin classes."""
    import inspect
    import mcos_b200 as mb
    for cls in (mb.MuCovObservationSimulator, mb.MuCovLedoitWolfObservationSimulator, mb.MuCovJackknifeObservationSimulator):
        assert list(inspect.signature(cls).parameters)[:3] == ["mu", "cov", "n_observations"]
        assert issubclass(cls, mb.AbstractObservationSimulator)
    assert list(inspect.signature(mb.DeNoiserCovarianceTransformer).parameters)[0] == "bandwidth"
    assert inspect.signature(mb.DeNoiserCovarianceTransformer).parameters["bandwidth"].default == .25
    assert list(inspect.signature(mb.NCOOptimizer).parameters)[:2] == ["max_num_clusters", "num_clustering_trials"]
    assert inspect.signature(mb.NCOOptimizer).parameters["num_clustering_trials"].default == 10
    assert list(inspect.signature(mb.simulate_optimizations).parameters)[:5] == \
        ["obs_simulator", "n_sims", "optimizers", "error_estimator", "covariance_transformers"]
    assert list(inspect.signature(mb.simulate_optimizations_from_price_history).parameters)[:7] == \
        ["price_history", "simulator_name", "n_observations", "n_sims", "optimizers", "error_estimator",
         "covariance_transformers"]
    assert [mb.MarkowitzOptimizer().name, mb.HRPOptimizer().name, mb.NCOOptimizer().name, mb.RiskParityOptimizer().name] == \
        ["markowitz", "HRP", "NCO", "Risk Parity"]
    sim = mb.MuCovObservationSimulator(np.zeros(3), np.eye(3), 7)
    assert sim.n_observations == 7 and sim.mu.shape == (3,) and sim.cov.shape == (3, 3)


def test_convert_price_history_matches_reference_values(stock, literals):
    import pandas as pd
    from mcos_b200.utils import convert_price_history
    prices = pd.DataFrame(stock["prices"], columns=[str(t) for t in stock["tickers"]], index=pd.to_datetime(stock["dates"]))
    mu, cov = convert_price_history(prices)
    assert np.array_equal(mu, stock["mu"]) and np.array_equal(cov, stock["cov"])
    tick = [str(t) for t in stock["tickers"]]
    assert abs(mu[tick.index("GOOG")] - literals["utils_values"][0]) < 1e-14       # reference tests/test_utils.py:9


_GLOO_WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from mcos_b200.mcos import gather_errors, shard_range
rank, world, n = int(sys.argv[2]), 2, 11
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=sys.argv[3], RANK=str(rank), WORLD_SIZE=str(world))
dist.init_process_group("gloo", rank=rank, world_size=world)
lo, hi = shard_range(n, rank, world)
local = np.stack([np.arange(lo, hi, dtype=float), -np.arange(lo, hi, dtype=float)], axis=1)
full = gather_errors(local, n)
want = np.stack([np.arange(n, dtype=float), -np.arange(n, dtype=float)], axis=1)
assert np.array_equal(full, want), full
# 3-D blocks (several estimators) travel the same way
full3 = gather_errors(local.reshape(hi - lo, 1, 2), n)
assert np.array_equal(full3, want.reshape(n, 1, 2))
dist.destroy_process_group()
print("ok", rank)
"""


def test_world_size_2_gather_over_gloo(tmp_path):
    """The one exchange step of the multi-GPU path (all_gather of per-simulation errors in simulation-id order),
    exercised with two CPU processes over gloo."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), port], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_lbfgsb_1d_restatement_matches_scipy():
    """mpfit_kernel's minimiser (one-variable L-BFGS-B + dcsrch/dcstep) in numpy vs scipy.optimize.minimize on the
    reference's objective: same iteration count, same minimiser to 1e-6."""
    import oracle
    from lbfgsb1d_ref import lbfgsb_1d, make_obj
    from mcos_b200.synthetic import block_diagonal_truth
    from scipy.optimize import minimize
    mu, cov = block_diagonal_truth(100)
    for seed in range(3):
        np.random.seed(seed)
        x = np.random.multivariate_normal(mu, cov, size=200)
        lam = np.sort(np.linalg.eigvalsh(oracle.cov_to_corr(np.cov(x, rowvar=False))))[::-1]
        out = minimize(lambda v: oracle.mp_fit_objective(v, lam, 2.0), .5, bounds=((1e-5, 1 - 1e-5),))
        xm, ok, it, nfev = lbfgsb_1d(make_obj(lam, 2.0))
        assert ok and out.success and it == out.nit and abs(xm - out.x[0]) < 1e-6


def test_summarize_errors_nan_accounting_and_interval():
    """`mean` / `stdev` stay the reference's (a NaN propagates as in mcos/mcos.py:36-40); the NaN count and
    NaN-excluded statistics sit next to them."""
    import mcos_b200 as mb
    rng = np.random.RandomState(3)
    e = rng.normal(1.0, 2.0, size=(4000, 2))
    e[5, 1] = np.nan
    e[77, 1] = np.nan
    df = mb.summarize_errors(e, [mb.MarkowitzOptimizer(), mb.HRPOptimizer()])
    assert list(df.index) == ["markowitz", "HRP"]
    assert df.loc["markowitz", "mean"] == np.mean(e[:, 0]) and df.loc["markowitz", "stdev"] == np.std(e[:, 0])
    assert df.loc["markowitz", "n_nan"] == 0 and df.loc["HRP", "n_nan"] == 2 and df.loc["HRP", "n_sims"] == 4000
    assert np.isnan(df.loc["HRP", "mean"]) and np.isnan(df.loc["HRP", "stdev"])
    ok = e[~np.isnan(e[:, 1]), 1]
    assert df.loc["HRP", "mean_valid"] == pytest.approx(ok.mean(), rel=1e-15)
    se = ok.std(ddof=1) / np.sqrt(ok.size)
    assert df.loc["HRP", "stderr"] == pytest.approx(se, rel=1e-14)
    assert df.loc["HRP", "ci_high"] - df.loc["HRP", "ci_low"] == pytest.approx(2 * 1.959963984540054 * se, rel=1e-12)
    assert df.loc["markowitz", "ci_low"] < 1.0 < df.loc["markowitz", "ci_high"]
