"""Fast HRP path against exact distances on the bench workload's de-noised covariances: python scratch/hrp_stress.py [sims]"""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from mcos_b200 import _lib
from mcos_b200.engine import Engine
from mcos_b200.synthetic import block_diagonal_truth
s = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n, t = 500, 1000
mu, cov = block_diagonal_truth(n)
eng = Engine(n, t, seed=77)
eng.set_truth(mu, cov)
eng.use_torch_stream()   # outputs are torch tensors: run the library on torch's current stream
bad = flagged = 0
for rep in range(s // 512):
    x = eng.sample(512, sim_id0=rep * 512, out_torch=True)
    _, c, _ = eng.moments(x, _lib.MOMENTS_LEDOIT_WOLF)
    del x
    c, _ = eng.denoise(c, t)
    res = {}
    for mode in (0, 1):
        eng.set_hrp_mode(mode)
        eng.hrp_fallback_count()
        w, info = eng.hrp(c)
        res[mode] = (w.cpu().numpy(), info["order"].cpu().numpy(), info["linkage"].cpu().numpy(), eng.hrp_fallback_count())
    flagged += res[0][3]
    for k in range(3):
        if not np.array_equal(res[0][k], res[1][k]):
            bad += 1
print(f"{s} simulations: mismatching outputs {bad}, fast-path fallbacks {flagged}")
