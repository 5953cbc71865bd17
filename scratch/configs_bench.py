"""Throughput of the other BASELINE configs (parity-test cases, not bench lines) -> profiles/configs_r01.md"""
import sys, time, json, numpy as np
sys.path.insert(0, '.')
import mcos_b200 as mb
from mcos_b200.engine import get_engine
from mcos_b200.synthetic import block_diagonal_truth

def run(name, sim, opts, est, trs, n_sims):
    mb.simulate_errors(sim, min(n_sims, 512), opts, est, trs)          # warm-up (allocations)
    eng = get_engine(np.asarray(sim.cov).shape[0], sim.n_observations, 0, sim.seed)
    eng.set_profiling(True)
    t0 = time.time()
    errs, st = mb.simulate_errors(sim, n_sims, opts, est, trs)
    dt = time.time() - t0
    kt = {k: round(v[0], 2) for k, v in eng.kernel_times().items() if v[1]}
    eng.set_profiling(False)
    top = sorted(kt.items(), key=lambda kv: -kv[1])[:4]
    print(json.dumps({"config": name, "n_sims": n_sims, "seconds": round(dt, 3), "sims_per_s": round(n_sims / dt, 1),
                      "status_nonzero": int((st != 0).sum()), "mean_err": [float(x) for x in np.nanmean(errs, 0)], "top_kernels_ms": top}))

which = sys.argv[1:] or ["1", "2", "4", "5"]
if "1" in which:
    st = np.load('tests/golden/stock.npz')
    run("1: README sample, 20 stocks, MuCov T=50, DeNoiser, Markowitz+HRP, Variance, 50 sims",
        mb.MuCovObservationSimulator(st['mu'], st['cov'], 50), [mb.MarkowitzOptimizer(), mb.HRPOptimizer()], mb.VarianceErrorEstimator(), [mb.DeNoiserCovarianceTransformer()], 50)
if "2" in which:
    mu, cov = block_diagonal_truth(100)
    run("2: N=100 T=200 MuCov + DeNoiser + Markowitz, Variance, 10k sims",
        mb.MuCovObservationSimulator(mu, cov, 200), [mb.MarkowitzOptimizer()], mb.VarianceErrorEstimator(), [mb.DeNoiserCovarianceTransformer()], 10000)
if "4" in which:
    mu, cov = block_diagonal_truth(100)
    run("4: N=100 T=200 Jackknife + NCO(max_num_clusters=10, trials=10), Variance, 50k sims",
        mb.MuCovJackknifeObservationSimulator(mu, cov, 200), [mb.NCOOptimizer(10, 10)], mb.VarianceErrorEstimator(), [], 50000)
if "5" in which:
    mu, cov = block_diagonal_truth(1000)
    run("5 (1 GPU share): N=1000 T=2000 LedoitWolf + DeNoiser, Markowitz+HRP+NCO(10), Sharpe, 2048 sims",
        mb.MuCovLedoitWolfObservationSimulator(mu, cov, 2000), [mb.MarkowitzOptimizer(), mb.HRPOptimizer(), mb.NCOOptimizer(10, 10)], mb.SharpeRatioErrorEstimator(), [mb.DeNoiserCovarianceTransformer()], 2048)
