"""nco_alloc_kernel alone with injected partitions: python scratch/nco_alloc_time.py"""
import numpy as np, torch
from mcos_b200.engine import Engine
from mcos_b200.synthetic import block_diagonal_truth
n, t, s = 1000, 2000, 296
mu, cov = block_diagonal_truth(n)
eng = Engine(n, t, seed=0)
eng.set_truth(mu, cov)
x = eng.sample(s, 0, out_torch=True)
mu_hat, covs, _ = eng.moments(x, 1)
del x
def run(labels, tag):
    lab = torch.from_numpy(labels.astype(np.int32)).cuda()
    eng.nco(mu_hat, covs, 10, 10, labels=lab)
    eng.set_profiling(True)
    eng.nco(mu_hat, covs, 10, 10, labels=lab)
    kt = eng.kernel_times()
    eng.set_profiling(False)
    ms, cnt = kt["nco_alloc"]
    print(f"{tag}: nco_alloc {ms / cnt:.2f} ms for {s} simulations")
base = np.tile(np.arange(n) // 100, (s, 1))
run(base, "10 clusters of 100 everywhere")
one = base.copy(); one[7] = np.arange(n) // 500
run(one, "one simulation with 2 clusters of 500")
allbig = np.tile(np.arange(n) // 500, (s, 1))
run(allbig, "2 clusters of 500 everywhere")
