"""Prototype of the eigvals hybrid (count-driven bracket + Illinois interpolation on p_N): evaluation counts and accuracy."""
import numpy as np, scipy.linalg as sl, sys, math
sys.path.insert(0, '.')
from mcos_b200.synthetic import block_diagonal_truth

def sturm(d, e2, x):
    """count of eigenvalues < x, and p_N as (mantissa, exponent)"""
    pm, pc, es, cnt = 0.0, 1.0, 0, 0
    sprev = 1
    for i in range(len(d)):
        pn = (d[i] - x) * pc - e2[i] * pm
        s = 1 if pn > 0 else (-1 if pn < 0 else -sprev)
        cnt += s != sprev
        sprev = s
        pm, pc = pc, pn
        if i % 8 == 7:
            m, ex = math.frexp(max(abs(pc), abs(pm)))
            if abs(ex) > 300:
                pc = math.ldexp(pc, -ex); pm = math.ldexp(pm, -ex); es += ex
    m, ex = math.frexp(pc) if pc != 0 else (0.0, -10**6)
    return cnt, abs(m), es + ex

def solve(d, e, hybrid=True):
    n = len(d)
    e2 = np.concatenate([[0.0], e[:n-1] ** 2])
    r = np.abs(np.concatenate([e[:n-1], [0]])) + np.abs(np.concatenate([[0], e[:n-1]]))
    lo, hi = (d - r).min(), (d + r).max()
    span = max(abs(lo), abs(hi)); pad = 2.2e-16 * span * 2 * n + 1e-300
    lo -= pad; hi += pad
    h = (hi - lo) / n
    grid = [sturm(d, e2, lo + h * i) for i in range(n + 1)]
    cnts = np.array([g[0] for g in grid])
    tol = 8.9e-16 * span
    lam = np.empty(n); nev = np.zeros(n, int)
    for j in range(n):
        l0 = int(np.searchsorted(cnts, j, side='right'))  # smallest i with cnts[i] > j
        a, b = lo + h * (l0 - 1), lo + h * l0
        ca, ma, ea = grid[l0 - 1]; cb, mb, eb = grid[l0]
        side = 0; w2 = b - a; w1 = b - a; force = False
        while True:
            mid = 0.5 * (a + b)
            if not (a < mid < b) or b - a <= tol: break
            x = mid
            if hybrid and ca == j and cb == j + 1 and not force and b - a > 1.2 * tol:
                de = max(-1000, min(1000, eb - ea))
                rr = (mb / ma) * 2.0 ** de if ma > 0 else 0.0
                t = 1.0 / (1.0 + rr)
                x = a + (b - a) * t
                x = min(max(x, a + 0.6 * tol), b - 0.6 * tol)
                if not (a < x < b): x = mid
            c, m, ex = sturm(d, e2, x); nev[j] += 1
            if c > j:
                b, cb, mb, eb = x, c, m, ex
                if side == 1: ea -= 1
                side = 1
            else:
                a, ca, ma, ea = x, c, m, ex
                if side == -1: eb -= 1
                side = -1
            force = (b - a) > 0.5 * w2
            w2, w1 = w1, b - a
        lam[j] = 0.5 * (a + b)
    return lam, nev

if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    t = 2 * n
    mu, cov = block_diagonal_truth(n)
    rng = np.random.RandomState(0)
    x = rng.multivariate_normal(mu, cov, size=t)
    c = np.corrcoef(x, rowvar=False)
    hh = sl.hessenberg(c)
    d = np.diag(hh).copy(); e = np.diag(hh, -1).copy(); e = np.concatenate([e, [0.0]])
    ref = sl.eigvalsh_tridiagonal(d, e[:n-1])
    for hyb in (False, True):
        lam, nev = solve(d, e, hyb)
        err = np.abs(lam - ref).max() / np.abs(ref).max()
        wmax = [nev[i:i + 32].max() for i in range(0, n, 32)]
        print(f"hybrid={hyb}: rel err {err:.2e}  evals mean {nev.mean():.1f} max {nev.max()}  mean of warp max {np.mean(wmax):.1f}")
    print(np.bincount(nev))
