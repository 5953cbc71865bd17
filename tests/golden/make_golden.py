"""Generates the committed parity fixtures under tests/golden/ from the REFERENCE ITSELF.

Run in the build container only (it needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

What it writes
  stock.npz              the reference's price fixture (tests/stock_prices.csv) as arrays, plus (mu, cov) derived from it
  reference_literals.npz the golden vectors the reference's own tests hold, parsed out of its test files with `ast`
  reference_runs.npz     outputs of the reference's classes (imported from /root/reference) on injected draws

PyPortfolioOpt 0.5.2 cannot be installed here, so `mcos.optimizer` / `mcos.utils` are imported against an in-memory
stand-in module built from this repo's restatement of the 0.5.2 behaviour (oracle.mcos_port.markowitz_slsqp etc.); the
stand-in is itself pinned by the Markowitz golden at tests/test_optimizer.py:17-22 (max diff 0.0).
The reference's NCOOptimizer does not run on this stack (sklearn >= 1.0 / pandas 3) and is not called.
"""
import ast
import os
import sys
import types
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import mcos_port as port  # noqa: E402
from mcos_b200.synthetic import block_diagonal_truth  # noqa: E402


def install_pypfopt_stand_in():
    pkg = types.ModuleType("pypfopt")
    ef_mod = types.ModuleType("pypfopt.efficient_frontier")
    rm = types.ModuleType("pypfopt.risk_models")
    er = types.ModuleType("pypfopt.expected_returns")

    class EfficientFrontier:
        def __init__(self, expected_returns, cov_matrix, weight_bounds=(0, 1), gamma=0):
            self.expected_returns, self.cov_matrix, self.gamma = expected_returns, cov_matrix, gamma
            self.tickers = list(range(len(expected_returns)))
            self.weights = None

        def max_sharpe(self, risk_free_rate=0.02):
            self.weights = port.markowitz_slsqp(self.expected_returns, self.cov_matrix, risk_free_rate, self.gamma)[0]
            return dict(zip(self.tickers, self.weights))

        def clean_weights(self, cutoff=1e-4, rounding=5):
            return dict(zip(self.tickers, port.clean_weights(self.weights, cutoff, rounding)))

    ef_mod.EfficientFrontier = EfficientFrontier
    rm.sample_cov = port.sample_cov
    er.mean_historical_return = port.mean_historical_return
    pkg.efficient_frontier, pkg.risk_models, pkg.expected_returns = ef_mod, rm, er
    sys.modules.update({"pypfopt": pkg, "pypfopt.efficient_frontier": ef_mod, "pypfopt.risk_models": rm,
                        "pypfopt.expected_returns": er})


def literal_arrays(path):
    """All list/np.array literals with >= 3 numeric entries found in a reference test file, in source order."""
    tree = ast.parse(open(path).read())
    found = []
    for node in ast.walk(tree):
        if isinstance(node, ast.List):
            try:
                val = np.array(ast.literal_eval(node), dtype=float)
            except Exception:
                continue
            if val.size >= 3:
                found.append((node.lineno, val))
    found.sort(key=lambda t: t[0])
    return found


def main():
    warnings.simplefilter("ignore")
    install_pypfopt_stand_in()
    from mcos import covariance_transformer as ref_ct, error_estimator as ref_ee, observation_simulator as ref_os, \
        optimizer as ref_opt

    # ---- stock fixture -------------------------------------------------------------------------------------------
    prices = pd.read_csv(f"{REF}/tests/stock_prices.csv", parse_dates=True, index_col="date")
    mu, cov = port.convert_price_history(prices)
    np.savez_compressed(f"{HERE}/stock.npz", prices=prices.to_numpy(), tickers=np.array(prices.columns, dtype="U8"),
                        dates=np.array(prices.index.strftime("%Y-%m-%d"), dtype="U10"), mu=mu, cov=cov)

    # ---- literals from the reference's tests ---------------------------------------------------------------------
    opt_lits = literal_arrays(f"{REF}/tests/test_optimizer.py")
    ct_lits = literal_arrays(f"{REF}/tests/test_covariance_transformer.py")
    ee_lits = literal_arrays(f"{REF}/tests/test_error_estimator.py")
    by_line = {}
    for ln, v in opt_lits:
        if ln not in by_line or v.size > by_line[ln].size:
            by_line[ln] = v
    lit = {
        "markowitz_w": by_line[17], "nco_max_sharpe_w": by_line[35], "nco_min_var_w": by_line[44],
        "hrp_w": by_line[61], "rp_cov": by_line[73], "rp_w": by_line[82], "rp_budget": by_line[85],
        "rp_budget_w": by_line[89], "rp_mu": by_line[71],
        "denoised_cov": next(v for ln, v in ct_lits if v.shape == (20, 20)),
        "detone_golden": next(v for ln, v in ct_lits if v.shape == (3, 3) and ln >= 24),
        "detone_corr": next(v for ln, v in ct_lits if v.shape == (3, 3) and ln < 24),
        "detone_std": next(v for ln, v in ct_lits if v.shape == (3,)),
        "est_allocation": next(v for ln, v in ee_lits if v.shape == (20,) and ln < 28),
        "est_optimal_allocation": next(v for ln, v in ee_lits if v.shape == (20,) and ln >= 28),
        "est_expected": np.array([0.00277877, 0.0044184, 0.04180305]),      # tests/test_error_estimator.py:13-15
        "utils_values": np.array([0.26770283812412754, 0.36378640487631986, 0.04620227917174632,
                                  0.014789747995647461]),                    # tests/test_utils.py:9-12
        "denoise_n_observations": np.array(prices.size),                     # tests/test_covariance_transformer.py:49
    }
    np.savez_compressed(f"{HERE}/reference_literals.npz", **lit)

    # ---- the reference's classes on injected draws ---------------------------------------------------------------
    runs = {}
    mu100, cov100 = block_diagonal_truth(100)
    mu60, cov60 = block_diagonal_truth(60, n_blocks=6, seed=3)
    cases = {"stock20_T50": (mu, cov, 50, 4), "block100_T200": (mu100, cov100, 200, 3),
             "block60_T30": (mu60, cov60, 30, 2)}
    denoiser = ref_ct.DeNoiserCovarianceTransformer()
    for name, (m, c, t, n_seeds) in cases.items():
        for seed in range(n_seeds):
            np.random.seed(seed)
            x = np.random.multivariate_normal(m.flatten(), c, size=t)
            key = f"{name}/s{seed}"
            runs[f"{key}/x"] = x
            for sim_name, cls in (("mucov", ref_os.MuCovObservationSimulator),
                                  ("lw", ref_os.MuCovLedoitWolfObservationSimulator),
                                  ("jk", ref_os.MuCovJackknifeObservationSimulator)):
                # replay the identical draw through the reference class
                np.random.seed(seed)
                mh, ch = cls(m, c, t).simulate()
                runs[f"{key}/{sim_name}_mu"], runs[f"{key}/{sim_name}_cov"] = mh, ch
            base = runs[f"{key}/lw_cov"] if t < len(m) else runs[f"{key}/mucov_cov"]
            mu_hat = runs[f"{key}/mucov_mu"]
            den = denoiser.transform(base, t)
            lam, vec = denoiser._get_PCA(ref_ct.cov_to_corr(base))
            lam_plus = denoiser._find_max_eigenvalue(np.diag(lam), t / base.shape[1])
            runs[f"{key}/denoise_in"] = base
            runs[f"{key}/denoised"] = den
            runs[f"{key}/eigvals"] = np.diag(lam).copy()
            runs[f"{key}/lambda_plus"] = np.array(lam_plus)
            runs[f"{key}/n_facts"] = np.array(lam.shape[0] - np.diag(lam)[::-1].searchsorted(lam_plus))
            runs[f"{key}/markowitz_w"] = ref_opt.MarkowitzOptimizer().allocate(mu_hat, den)
            hrp = ref_opt.HRPOptimizer()
            runs[f"{key}/hrp_w"] = hrp.allocate(mu_hat, den)
            import scipy.cluster.hierarchy as sch
            link = sch.linkage(hrp._correlation_distance(ref_ct.cov_to_corr(den)), "single")
            runs[f"{key}/hrp_link"] = link
            runs[f"{key}/hrp_order"] = np.array(hrp._quasi_diagonal_cluster_sequence(link), dtype=np.int64)
            ideal = ref_opt.HRPOptimizer().allocate(m, c)
            for est_name, est in (("eo", ref_ee.ExpectedOutcomeErrorEstimator()),
                                  ("var", ref_ee.VarianceErrorEstimator()),
                                  ("sharpe", ref_ee.SharpeRatioErrorEstimator())):
                runs[f"{key}/err_{est_name}"] = np.array(est.estimate(m, c, runs[f"{key}/hrp_w"], ideal)).reshape(-1)
    # reference optimizers on the stock truth itself (what the reference's unit tests do)
    runs["stock20/markowitz_w"] = ref_opt.MarkowitzOptimizer().allocate(mu, cov)
    runs["stock20/hrp_w"] = ref_opt.HRPOptimizer().allocate(mu, cov)
    runs["stock20/denoised"] = denoiser.transform(cov, prices.size)
    runs["stock20/risk_parity_w"] = ref_opt.RiskParityOptimizer().allocate(mu, cov)
    np.savez_compressed(f"{HERE}/reference_runs.npz", **runs)
    for f in ("stock.npz", "reference_literals.npz", "reference_runs.npz"):
        print(f, os.path.getsize(f"{HERE}/{f}"), "bytes")


if __name__ == "__main__":
    main()
