import sys
import numpy as np, torch
sys.path.insert(0, '.')
from mcos_b200 import _lib
from mcos_b200.engine import Engine
from mcos_b200.synthetic import block_diagonal_truth
n, t = 500, 1000
mu, cov = block_diagonal_truth(n)
eng = Engine(n, t, seed=77)
eng.set_truth(mu, cov)
eng.use_torch_stream()   # outputs are torch tensors: run the library on torch's current stream
x = eng.sample(512, sim_id0=0, out_torch=True)
_, c, _ = eng.moments(x, _lib.MOMENTS_LEDOIT_WOLF)
del x
c, _ = eng.denoise(c, t)
res = {}
for mode in (0, 1, 3):
    eng.set_hrp_mode(mode)
    eng.hrp_fallback_count()
    w, info = eng.hrp(c)
    res[mode] = (w.cpu().numpy(), info["order"].cpu().numpy(), info["linkage"].cpu().numpy(), eng.hrp_fallback_count())
for mode in (0, 3):
    for k, name in enumerate(("weights", "order", "linkage")):
        a, b = res[mode][k], res[1][k]
        diff = [i for i in range(512) if not np.array_equal(a[i], b[i])]
        print("mode", mode, name, "sims differing:", len(diff), diff[:5])
        if diff and name == "linkage":
            i = diff[0]
            rows = np.where((a[i] != b[i]).any(axis=1))[0]
            print("  rows", rows[:6], "of", a[i].shape[0])
            for r in rows[:4]:
                print("   fast", a[i][r], "exact", b[i][r])
cc = c.cpu().numpy()[0]
print("symmetric cov?", np.array_equal(cc, cc.T), np.abs(cc - cc.T).max())
