"""Runs only the two-stage tridiagonalisation on S correlation matrices (for ncu). python scratch/sbr_only.py [N S method]"""
import sys
import numpy as np, torch
from mcos_b200.engine import Engine
n, s, method = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (500, 296, 2)))
rng = np.random.RandomState(0)
x = rng.standard_normal((2 * n, n)) + 0.5 * rng.standard_normal((2 * n, 1))
c = np.corrcoef(x, rowvar=False)
a = torch.from_numpy(np.broadcast_to(c, (s, n, n)).copy()).cuda()
eng = Engine(n, 2 * n)
eng.set_eig_method(method)
for _ in range(2):
    d, e = eng.tridiagonalize(a)
torch.cuda.synchronize()
print("ok", float(d.sum()))
