"""Small end-to-end pass over every kernel for compute-sanitizer (memcheck / racecheck)."""
import sys, numpy as np
sys.path.insert(0, '.')
import mcos_b200 as mb
from mcos_b200 import _lib
from mcos_b200.engine import Engine
from mcos_b200.synthetic import block_diagonal_truth
for n, t in ((20, 30), (200, 260)):
    mu, cov = block_diagonal_truth(n, n_blocks=2 if n == 20 else 10)
    sim = mb.MuCovLedoitWolfObservationSimulator(mu, cov, t, seed=1)
    opts = [mb.MarkowitzOptimizer(), mb.HRPOptimizer(), mb.NCOOptimizer(4, 2), mb.RiskParityOptimizer()]
    errs, st = mb.simulate_errors(sim, 3, opts, mb.SharpeRatioErrorEstimator(),
                                  [mb.DeNoiserCovarianceTransformer(), mb.DetoneCovarianceTransformer(1)])
    print(n, t, errs.shape, st)
eng = Engine(20, 30)
print(eng.risk_parity(np.stack([np.eye(20) * 2 + 1.0]))[0].sum())
