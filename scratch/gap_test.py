"""Where does the non-kernel time of a step go? python scratch/gap_test.py"""
import time
import numpy as np, torch
import mcos_b200 as mb
from mcos_b200.engine import Engine
from mcos_b200.synthetic import block_diagonal_truth
n, t, s = 500, 1000, 2048
mu, cov = block_diagonal_truth(n)
sim = mb.MuCovLedoitWolfObservationSimulator(mu, cov, t)
plan = mb.build_plan(sim, [mb.MarkowitzOptimizer(), mb.HRPOptimizer()], mb.SharpeRatioErrorEstimator(), [mb.DeNoiserCovarianceTransformer()])
eng = Engine(n, t)
eng.set_truth(mu, cov)
eng.use_torch_stream()
for _ in range(2):
    eng.run(plan, s, 0, out_torch=True)
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    eng.run(plan, s, 0, out_torch=True)
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    st = eng.last_stage_ms()
    print(f"wall {1e3*(t1-t0):.2f} ms  events {e0.elapsed_time(e1):.2f} ms  stages {sum(st.values()):.2f} ms", {k: round(v, 1) for k, v in st.items()})
