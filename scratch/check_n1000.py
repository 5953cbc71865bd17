import sys, time, numpy as np
sys.path.insert(0, '.')
import oracle
from mcos_b200 import _lib
from mcos_b200.engine import Engine
from mcos_b200.synthetic import block_diagonal_truth
n, t, s = 1000, 2000, 3
mu, cov = block_diagonal_truth(n)
eng = Engine(n, t, seed=1)
eng.set_truth(mu, cov)
x = eng.sample(s)
mu_hat, cov_hat, sh = eng.moments(x, _lib.MOMENTS_LEDOIT_WOLF)
for i in range(s):
    want, wsh = oracle.ledoit_wolf_formula(x[i])
    print('LW', i, np.abs(cov_hat[i] - want).max() / np.abs(want).max(), abs(sh[i] - wsh))
den, info = eng.denoise(cov_hat, t)
for i in range(s):
    d = oracle.denoise_details(cov_hat[i], t)
    print('denoise', i, info['n_facts'][i], d['n_facts'], np.abs(den[i] - d['cov']).max() / np.abs(d['cov']).max(), np.abs(info['eigvals'][i] - d['eigvals']).max(), info['status'][i])
w, minfo = eng.markowitz(mu_hat, den, clean=False)
for i in range(s):
    want, _ = oracle.markowitz_exact_qp(mu_hat[i], den[i])
    print('markowitz', i, np.abs(w[i] - want).max(), minfo['iters'][i], minfo['status'][i], (want > 0).sum())
wh, hinfo = eng.hrp(den)
for i in range(s):
    d = oracle.hrp_details(den[i])
    print('hrp', i, np.array_equal(hinfo['order'][i], d['order']), np.array_equal(hinfo['linkage'][i], d['linkage']), np.abs(wh[i] - d['weights']).max() / d['weights'].max())
p = _lib.Plan(); p.moments_kind = 1; p.denoise = 1; p.bandwidth = .25; p.n_optimizers = 2; p.optimizers[0] = 0; p.optimizers[1] = 1
p.error_kind = 2; p.risk_free_rate = .02; p.wave = 0
eng.run(p, 296)
eng.set_profiling(True)
t0 = time.time(); err, st = eng.run(p, 592); dt = time.time() - t0
print(f"N={n} T={t} S=592: {dt:.3f}s -> {592/dt:.1f} sims/s; status nonzero {(st != 0).sum()}")
for k, (ms, c) in sorted(eng.kernel_times().items(), key=lambda kv: -kv[1][0])[:8]:
    print(f"  {k:16s} {ms:10.2f} ms {ms/592*1e3:8.2f} us/sim")
