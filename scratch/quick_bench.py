"""Quick per-kernel timing of the DeNoiser path for both tridiagonalisation methods. python scratch/quick_bench.py [N T S]"""
import sys, time
import numpy as np, torch
from mcos_b200.engine import Engine
from mcos_b200.synthetic import block_diagonal_truth
import mcos_b200 as mb
n, t, s = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (500, 1000, 1184)))
mu, cov = block_diagonal_truth(n)
sim = mb.MuCovLedoitWolfObservationSimulator(mu, cov, t)
eng = Engine(n, t)
eng.set_truth(mu, cov)
x = eng.sample(s, out_torch=True)
_, c, _ = eng.moments(x, 1)
del x
for method in (1, 2):
    eng.set_eig_method(method)
    for rep in range(2):
        eng.set_profiling(True)
        out, info = eng.denoise(c, t)
        kt = eng.kernel_times()
    tot = sum(v[0] for v in kt.values())
    print(f"method {method}: total {tot:.2f} ms for {s} matrices = {1e3*tot/s:.2f} us/matrix")
    for k, (ms, cnt) in kt.items():
        if cnt:
            print(f"   {k:14s} {ms:8.3f} ms  {1e3*ms/s:7.2f} us/matrix")
