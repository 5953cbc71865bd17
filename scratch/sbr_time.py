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
import os
if os.environ.get("BENCHDATA"):
    # the matrices the bench feeds the eigensolver: correlation of Ledoit-Wolf covariances of the block-diagonal truth
    from mcos_b200.synthetic import block_diagonal_truth
    mu, cov = block_diagonal_truth(n)
    e2 = Engine(n, 2 * n, seed=3)
    e2.set_truth(mu, cov)
    a = torch.empty((s, n, n), dtype=torch.float64, device="cuda")
    for i0 in range(0, s, 512):
        m = min(512, s - i0)
        xx = e2.sample(m, i0, out_torch=True)
        _, ch, _ = e2.moments(xx, 1)
        dd = torch.diagonal(ch, dim1=1, dim2=2).rsqrt()
        a[i0:i0 + m] = ch * dd[:, :, None] * dd[:, None, :]
        del xx, ch
    e2.close()
    torch.cuda.empty_cache()
    c = a[0].cpu().numpy()
elif os.environ.get("DISTINCT"):
    g = torch.Generator(device="cuda").manual_seed(1)
    xs = torch.randn((s, 2 * n, n), dtype=torch.float64, device="cuda", generator=g) if s * n * n * 16 < 40e9 else None
    if xs is None:
        a = torch.empty((s, n, n), dtype=torch.float64, device="cuda")
        for i0 in range(0, s, 256):
            xx = torch.randn((min(256, s - i0), 2 * n, n), dtype=torch.float64, device="cuda", generator=g)
            xx = xx - xx.mean(1, keepdim=True)
            cc = xx.transpose(1, 2) @ xx
            dd = torch.diagonal(cc, dim1=1, dim2=2).rsqrt()
            a[i0:i0 + xx.shape[0]] = cc * dd[:, :, None] * dd[:, None, :]
    else:
        xs = xs - xs.mean(1, keepdim=True)
        cc = xs.transpose(1, 2) @ xs
        dd = torch.diagonal(cc, dim1=1, dim2=2).rsqrt()
        a = (cc * dd[:, :, None] * dd[:, None, :]).contiguous()
        del xs, cc
    c = a[0].cpu().numpy()
else:
    a = torch.from_numpy(np.broadcast_to(c, (s, n, n)).copy()).cuda()
eng = Engine(n, 2 * n)
eng.set_eig_method(2)
d, e = eng.tridiagonalize(a)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(reps):
    d, e = eng.tridiagonalize(a)
t1.record()
torch.cuda.synchronize()
print(f"whole tridiagonalize call (copy + stage 1 + stage 2), torch events: {1e3 * t0.elapsed_time(t1) / reps / s:.2f} us/matrix")
eng.set_profiling(True)
for _ in range(reps):
    d, e = eng.tridiagonalize(a)
kt = eng.kernel_times()
eng.set_profiling(False)
ref = np.linalg.eigvalsh(c)
dd, ee = d[0].cpu().numpy(), e[0].cpu().numpy()
t = np.diag(dd) + np.diag(ee[:n - 1], 1) + np.diag(ee[:n - 1], -1)
err = np.abs(np.linalg.eigvalsh(t) - ref).max()
print(f"N={n} S={s} eig err {err:.2e} :: " + "  ".join(f"{k} {1e3 * ms / reps / s:.2f} us/matrix ({cnt // reps} launches)" for k, (ms, cnt) in kt.items() if cnt))
