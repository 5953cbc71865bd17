"""Times the two-stage tridiagonalisation kernels alone: python scratch/sbr_time.py N S [reps]; prints per-kernel us per matrix
(CUDA events per launch, mcosb_set_profiling).  MCOSB_SY2SB_OLD=1 selects the round-1 stage-1 kernel."""
import sys
import numpy as np, torch
from mcos_b200.engine import Engine
n, s = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rng = np.random.RandomState(0)
x = rng.standard_normal((2 * n, n)) + 0.5 * rng.standard_normal((2 * n, 1))
c = np.corrcoef(x, rowvar=False)
a = torch.from_numpy(np.broadcast_to(c, (s, n, n)).copy()).cuda()
eng = Engine(n, 2 * n)
eng.set_eig_method(2)
d, e = eng.tridiagonalize(a)
torch.cuda.synchronize()
eng.set_profiling(True)
for _ in range(reps):
    d, e = eng.tridiagonalize(a)
kt = eng.kernel_times()
eng.set_profiling(False)
ref = np.linalg.eigvalsh(c)
dd, ee = d[0].cpu().numpy(), e[0].cpu().numpy()
t = np.diag(dd) + np.diag(ee[:n - 1], 1) + np.diag(ee[:n - 1], -1)
err = np.abs(np.linalg.eigvalsh(t) - ref).max()
print(f"N={n} S={s} eig err {err:.2e} :: " + "  ".join(f"{k} {1e3 * ms / cnt / s:.2f} us/matrix" for k, (ms, cnt) in kt.items() if cnt))
