"""Stage-1 check for whichever dense -> band kernel MCOSB_STAGE1 selects: band and tridiagonal eigenvalues against numpy over
ragged sizes.  python scratch/s3_check.py [N ...]"""
import sys
import numpy as np
from mcos_b200.engine import Engine

sizes = [int(a) for a in sys.argv[1:]] or [10, 11, 16, 17, 18, 25, 33, 64, 65, 72, 73, 96, 100, 129, 136, 137, 200, 255, 256, 257, 333, 500,
                                           511, 512, 513, 520, 696, 777, 1000, 1023, 1024]
worst = 0.0
for n in sizes:
    rng = np.random.RandomState(n)
    a = []
    for k in range(3):
        t = 2 * n if k % 2 == 0 else max(2, n // 2)
        a.append(np.corrcoef(rng.standard_normal((t, n)) + 0.3 * rng.standard_normal((t, 1)), rowvar=False))
    a = np.stack(a)
    eng = Engine(n, 2 * n)
    eng.set_eig_method(2)
    d, e, work = eng.tridiagonalize(a, want_work=True)
    for i in range(a.shape[0]):
        ref = np.linalg.eigvalsh(a[i])
        band = np.tril(work[i]) - np.tril(work[i], -9)
        band = band + np.tril(band, -1).T
        eb = np.abs(np.linalg.eigvalsh(band) - ref).max() if np.isfinite(band).all() else np.inf
        tt = np.diag(d[i]) + np.diag(e[i][:n - 1], 1) + np.diag(e[i][:n - 1], -1)
        et = np.abs(np.linalg.eigvalsh(tt) - ref).max() if np.isfinite(tt).all() else np.inf
        worst = max(worst, eb / n, et / n)
        flag = "" if max(eb, et) < 2e-13 * n else "   <-- FAIL"
        print(f"N={n} matrix {i}: band err {eb:.2e} tri err {et:.2e}{flag}", flush=True)
    eng.close()
print("worst err / N:", worst)
