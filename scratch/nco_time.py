"""NCO stage alone on sampled covariances: python scratch/nco_time.py N T S  (per-kernel us per simulation)"""
import sys
import numpy as np
from mcos_b200.engine import Engine
from mcos_b200.synthetic import block_diagonal_truth
n, t, s = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mu, cov = block_diagonal_truth(n)
eng = Engine(n, t, seed=1)
eng.set_truth(mu, cov)
import torch
x = eng.sample(s, 0, out_torch=True)
mu_hat, covs, _ = eng.moments(x, 0)
del x
w, info = eng.nco(mu_hat, covs, 10, 10)
torch.cuda.synchronize()
eng.set_profiling(True)
w, info = eng.nco(mu_hat, covs, 10, 10)
kt = eng.kernel_times()
print(f"N={n} S={s}: " + "  ".join(f"{k} {1e3 * ms / cnt / s:.2f}" for k, (ms, cnt) in kt.items() if cnt) + " us/sim")
