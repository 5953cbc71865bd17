"""What partitions does NCO find at config 5's shape? python scratch/nco_partitions.py [N T S]"""
import sys, collections
import numpy as np, torch
from mcos_b200.engine import Engine
from mcos_b200.synthetic import block_diagonal_truth
n, t, s = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (1000, 2000, 64)
mu, cov = block_diagonal_truth(n)
eng = Engine(n, t, seed=0)
eng.set_truth(mu, cov)
x = eng.sample(s, 0, out_torch=True)
mu_hat, covs, _ = eng.moments(x, 1)
del x
covs, _ = eng.denoise(covs, t)
w, info = eng.nco(mu_hat, covs, 10, 10)
lab = info["labels"].cpu().numpy()
ks = collections.Counter(len(set(l.tolist())) for l in lab)
mx = collections.Counter(int(np.bincount(l).max()) for l in lab)
print("clusters per simulation:", dict(ks))
print("largest cluster size:", dict(sorted(mx.items())))
eng.set_profiling(True)
w, info = eng.nco(mu_hat, covs, 10, 10)
kt = eng.kernel_times()
print("  ".join(f"{k} {1e3 * ms / cnt / s:.2f}" for k, (ms, cnt) in kt.items() if cnt), "us/sim")
