"""Per-kernel device time of BASELINE config 5's shape (N=1000, T=2000, LW + DeNoiser, Markowitz+HRP+NCO(10)+RiskParity,
three estimators in one pass) on one GPU: python scratch/config5_profile.py [n_sims] [N T]"""
import sys
import numpy as np
import mcos_b200 as mb
from mcos_b200.engine import get_engine
from mcos_b200.synthetic import block_diagonal_truth
s = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n, t = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1000, 2000)
mu, cov = block_diagonal_truth(n)
sim = mb.MuCovLedoitWolfObservationSimulator(mu, cov, t, seed=0)
opts = [mb.MarkowitzOptimizer(), mb.HRPOptimizer(), mb.NCOOptimizer(10, 10), mb.RiskParityOptimizer()]
ests = [mb.ExpectedOutcomeErrorEstimator(), mb.VarianceErrorEstimator(), mb.SharpeRatioErrorEstimator()]
plan = mb.build_plan(sim, opts, ests, [mb.DeNoiserCovarianceTransformer()])
eng = get_engine(n, t, 0, 0)
eng.ensure_truth(mu, cov)
eng.run(plan, s, 0)
eng.set_profiling(True)
err, st = eng.run(plan, s, s)
kt = eng.kernel_times()
eng.set_profiling(False)
tot = sum(ms for ms, c in kt.values())
print(f"N={n} T={t} sims={s}: {tot / s * 1e3:.1f} us per simulation ({s / tot * 1e3:.0f} sims/s), status nonzero {(st != 0).sum()}")
for k, (ms, c) in sorted(kt.items(), key=lambda kv: -kv[1][0]):
    if c:
        print(f"  {k:14s} {ms / s * 1e3:8.2f} us/sim  {100 * ms / tot:5.1f} %  ({c} launches)")
