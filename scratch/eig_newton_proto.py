"""Prototype: count-bracketed Newton on the characteristic polynomial (p, p' from the three-term recurrence)."""
import numpy as np, scipy.linalg as sl, sys, math, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from mcos_b200.synthetic import block_diagonal_truth

def sturm(d, e2, x):
    pm, pc, qm, qc, cnt, sprev = 0.0, 1.0, 0.0, 0.0, 0, 1
    for i in range(len(d)):
        pn = (d[i] - x) * pc - e2[i] * pm
        qn = (d[i] - x) * qc - pc - e2[i] * qm
        s = 1 if pn > 0 else (-1 if pn < 0 else -sprev)
        cnt += s != sprev; sprev = s
        pm, pc, qm, qc = pc, pn, qc, qn
        if i % 8 == 7:
            m, ex = math.frexp(max(abs(pc), abs(pm), abs(qc), abs(qm)))
            if abs(ex) > 300:
                pc = math.ldexp(pc, -ex); pm = math.ldexp(pm, -ex); qc = math.ldexp(qc, -ex); qm = math.ldexp(qm, -ex)
    return cnt, pc, qc

def solve(d, e, mode):
    n = len(d)
    e2 = np.concatenate([[0.0], e[:n-1] ** 2])
    r = np.abs(np.concatenate([e[:n-1], [0]])) + np.abs(np.concatenate([[0], e[:n-1]]))
    lo, hi = (d - r).min(), (d + r).max()
    span = max(abs(lo), abs(hi)); pad = 2.2e-16 * span * 2 * n + 1e-300
    lo -= pad; hi += pad
    h = (hi - lo) / n
    cnts = np.array([sturm(d, e2, lo + h * i)[0] for i in range(n + 1)])
    tol = 8.9e-16 * span
    lam = np.empty(n); nev = np.zeros(n, int); cheap = np.zeros(n, int); probes = {}
    for j in range(n):
        l0 = int(np.searchsorted(cnts, j, side='right'))
        a, b = lo + h * (l0 - 1), lo + h * l0
        ca, cb = cnts[l0 - 1], cnts[l0]
        if mode and cb - ca > 1:      # second multisection step: the cell's m eigenvalues probe m interior points (count only)
            c0, c1 = ca, cb
            if (l0, 'p') not in probes:
                probes[(l0, 'p')] = [(a + h * (r + 0.5) / (c1 - c0), sturm(d, e2, a + h * (r + 0.5) / (c1 - c0))[0]) for r in range(c1 - c0)]
            cheap[j] += 1
            for xx, cc in probes[(l0, 'p')]:
                if cc > j:
                    if xx < b: b, cb = xx, cc
                elif xx > a: a, ca = xx, cc
        x = 0.5 * (a + b); dxp = b - a; nw = False
        while True:
            c, p, dp = sturm(d, e2, x); nev[j] += 1
            if c > j: b, cb = x, c
            else: a, ca = x, c
            mid = 0.5 * (a + b)
            if not (a < mid < b) or b - a <= tol:
                lam[j] = mid; break
            xn = mid; dxn = b - a; nwn = False
            if mode and ca == j and cb == j + 1 and dp != 0.0:
                dx = -p / dp
                t = min(max(x + dx, a), b)
                adx = abs(dx)
                if adx <= 32 * tol or (nw and adx * adx * adx <= 0.1 * tol * dxp * dxp and adx < 0.01 * dxp):
                    lam[j] = t; break
                if a < t < b and adx < 0.5 * dxp:
                    xn = t; dxn = adx; nwn = True
            nw = nwn
            x = xn; dxp = dxn
    return lam, nev

if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    t = 2 * n
    mu, cov = block_diagonal_truth(n)
    rng = np.random.RandomState(seed)
    x = rng.multivariate_normal(mu, cov, size=t)
    c = np.corrcoef(x, rowvar=False)
    hh = sl.hessenberg(c)
    d = np.diag(hh).copy(); e = np.concatenate([np.diag(hh, -1), [0.0]])
    ref = sl.eigvalsh_tridiagonal(d, e[:n-1])
    for mode in (0, 1):
        lam, nev = solve(d, e, mode)
        err = np.abs(lam - ref).max() / np.abs(ref).max()
        wmax = [nev[i:i + 32].max() for i in range(0, n, 32)]
        print(f"newton={mode}: rel err {err:.2e}  evals mean {nev.mean():.1f} max {nev.max()}  mean of warp max {np.mean(wmax):.1f}")
        print(np.bincount(nev))
