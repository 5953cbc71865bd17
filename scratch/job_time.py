"""Wall time of the 100 000-simulation job (BASELINE config 3) from a cold and from a warm engine: python scratch/job_time.py"""
import time, numpy as np, torch
import mcos_b200 as mb
from mcos_b200.synthetic import block_diagonal_truth
mu, cov = block_diagonal_truth(500)
opts = [mb.MarkowitzOptimizer(), mb.HRPOptimizer()]
for rep in range(3):
    sim = mb.MuCovLedoitWolfObservationSimulator(mu, cov, 1000, seed=rep)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    df = mb.simulate_optimizations(sim, 100000, opts, mb.SharpeRatioErrorEstimator(), [mb.DeNoiserCovarianceTransformer()])
    torch.cuda.synchronize()
    print(f"run {rep}: {time.perf_counter() - t0:.3f} s")
