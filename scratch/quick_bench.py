import sys, time, numpy as np
sys.path.insert(0, '.')
from mcos_b200 import _lib
from mcos_b200.engine import Engine
from mcos_b200.synthetic import block_diagonal_truth
n, t, s = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mu, cov = block_diagonal_truth(n)
eng = Engine(n, t)
eng.set_truth(mu, cov)
p = _lib.Plan(); p.moments_kind = 1; p.denoise = 1; p.bandwidth = .25; p.n_optimizers = 2; p.optimizers[0] = 0; p.optimizers[1] = 1
p.error_kind = 2; p.risk_free_rate = .02; p.wave = 0
eng.run(p, s)
eng.set_profiling(True)
t0 = time.time(); err, st = eng.run(p, s); dt = time.time() - t0
print(f"N={n} T={t} S={s}: {dt:.3f}s -> {s/dt:.1f} sims/s; status nonzero: {(st!=0).sum()}; mean err {np.nanmean(err,0)}")
print({k: round(v, 2) for k, v in eng.last_stage_ms().items()})
kt = eng.kernel_times()
for k, (ms, c) in sorted(kt.items(), key=lambda kv: -kv[1][0]):
    if c: print(f"  {k:16s} {ms:10.2f} ms  {c:5d} launches  {ms/s*1e3:8.2f} us/sim")
