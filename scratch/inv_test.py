import sys, time, numpy as np
sys.path.insert(0, '.')
from mcos_b200 import _lib
from mcos_b200.engine import Engine
from mcos_b200.synthetic import block_diagonal_truth
for n, t, s in ((1000, 2000, 148), (1000, 2000, 1184), (750, 1500, 592), (500, 1000, 592)):
    mu, cov = block_diagonal_truth(n)
    eng = Engine(n, t, seed=1)
    eng.set_truth(mu, cov)
    p = _lib.Plan(); p.moments_kind = 1; p.denoise = 1; p.bandwidth = .25; p.n_optimizers = 1; p.optimizers[0] = 1
    p.error_kind = 2; p.risk_free_rate = .02; p.wave = 0
    eng.run(p, s)
    eng.set_profiling(True)
    eng.run(p, s)
    kt = eng.kernel_times()
    print(n, s, {k: round(v[0] / s * 1e3, 1) for k, v in kt.items() if k in ('inviter', 'backtransform', 'reconstruct', 'tridiag2', 'eigvals', 'mpfit')})
    eng.close()
