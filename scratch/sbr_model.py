"""numpy model of the two-stage tridiagonalisation (sy2sb.cu / sb2st.cu): data layout and update order as on the GPU.
Run: python scratch/sbr_model.py [N] [b]"""
import sys

import numpy as np


def house(x):
    """(v, tau, beta): (I - tau v v') x = beta e1, v[0] = 1 (LAPACK dlarfg convention)."""
    x = np.asarray(x, dtype=float)
    alpha = x[0]
    xn2 = float(x[1:] @ x[1:])
    v = np.zeros_like(x)
    v[0] = 1.0
    if xn2 == 0.0:
        return v, 0.0, alpha
    beta = -np.copysign(np.sqrt(alpha * alpha + xn2), alpha)
    tau = (beta - alpha) / beta
    v[1:] = x[1:] / (alpha - beta)
    return v, tau, beta


def sy2sb(A, b):
    """Lower triangle of A -> band (half bandwidth b). Returns (band matrix (full symmetric for checking), Vrows, taus).
    Vrows[k] holds reflector k = j0 + c: u[k + b] = 1, u[j] = Vrows[k][j] for j > k + b."""
    n = A.shape[0]
    A = np.tril(A).copy()
    Vrows = np.zeros((n, n))
    taus = np.zeros(n)
    Vp = np.zeros((n, b))
    Wp = np.zeros((n, b))
    j1 = 0
    while True:
        # a. bring the next panel's columns up to date (rows >= j1)
        nb = min(b, n - j1)
        for c in range(nb):
            col = j1 + c
            rows = np.arange(col, n)
            A[rows, col] -= Vp[rows] @ Wp[col] + Wp[rows] @ Vp[col]
        r0 = j1 + b                      # first row of the block to factorise
        if n - r0 < 2:                   # nothing left to eliminate: update the remaining columns and stop
            for col in range(j1 + nb, n):
                rows = np.arange(col, n)
                A[rows, col] -= Vp[rows] @ Wp[col] + Wp[rows] @ Vp[col]
            break
        m = n - r0
        P = A[r0:, j1:j1 + b].copy()     # m x b
        Vn = np.zeros((n, b))
        tau = np.zeros(b)
        ncol = min(b, m - 1)             # columns that still have something below the diagonal
        for c in range(ncol):
            v, t, beta = house(P[c:, c])
            tau[c] = t
            P[c, c] = beta
            P[c + 1:, c] = 0.0
            if c + 1 < b:
                w = t * (v @ P[c:, c + 1:])
                P[c:, c + 1:] -= np.outer(v, w)
            Vn[r0 + c:, c] = v
            Vrows[j1 + c, r0 + c:] = v
            taus[j1 + c] = t
        A[r0:, j1:j1 + b] = P
        # T (compact WY): Q = H_0 ... H_{b-1} = I - V T V'
        T = np.zeros((b, b))
        G = Vn.T @ Vn
        for c in range(b):
            T[c, c] = tau[c]
            if c:
                T[:c, c] = -tau[c] * (T[:c, :c] @ G[:c, c])
        # b. fused pass over the trailing lower triangle: update with (Vp, Wp), then Z = A22' Vn
        Z = np.zeros((n, b))
        for i in range(r0, n):
            for j in range(r0, i + 1):
                a = A[i, j] - (Vp[i] @ Wp[j] + Wp[i] @ Vp[j])
                A[i, j] = a
                Z[i] += a * Vn[j]
                if i != j:
                    Z[j] += a * Vn[i]
        # c. W = Z T - 1/2 V (T' V' Z T)
        Y = Z @ T
        M = T.T @ (Vn.T @ Y)
        Wn = Y - 0.5 * Vn @ M
        Vp, Wp = Vn, Wn
        j1 += b
    B = np.tril(A)
    B = B + np.tril(B, -1).T
    return B, Vrows, taus


def sb2st(B, b):
    """Band -> tridiagonal by bulge chasing (one column per sweep). Q2[s][i - s - 1]: reflector entries aligned with row i
    (slot of the leading 1 holds tau)."""
    n = B.shape[0]
    B = B.copy()
    Q2 = np.zeros((n, n))
    for s in range(n - 2):
        k = 0
        while True:
            lo = s + 1 + k * b
            hi = min(lo + b, n)          # rows [lo, hi)
            if hi - lo < 2:
                break
            col = s if k == 0 else lo - b
            v, tau, beta = house(B[lo:hi, col])
            Q2[s, lo - s - 1] = tau
            Q2[s, lo - s:hi - s - 1] = v[1:]
            H = np.eye(hi - lo) - tau * np.outer(v, v)
            B[lo:hi, :] = H @ B[lo:hi, :]
            B[:, lo:hi] = B[:, lo:hi] @ H
            k += 1
            if lo + b >= n:
                break
    return np.diag(B).copy(), np.diag(B, -1).copy(), Q2, B


def apply_q2(Q2, y, b):
    n = len(y)
    y = y.copy()
    for s in range(n - 3, -1, -1):
        k = 0
        while True:
            lo = s + 1 + k * b
            hi = min(lo + b, n)
            if hi - lo < 2:
                break
            tau = Q2[s, lo - s - 1]
            v = np.concatenate([[1.0], Q2[s, lo - s:hi - s - 1]])
            y[lo:hi] -= tau * v * (v @ y[lo:hi])
            k += 1
    return y


def apply_q1(Vrows, taus, y, b):
    n = len(y)
    y = y.copy()
    for k in range(n - 1, -1, -1):
        if taus[k] == 0.0:
            continue
        u = Vrows[k].copy()
        u[:k + b] = 0.0
        u[k + b] = 1.0
        y -= taus[k] * u * (u @ y)
    return y


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 61
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    rng = np.random.RandomState(1)
    X = rng.standard_normal((2 * n, n))
    A = np.corrcoef(X, rowvar=False)
    B, Vrows, taus = sy2sb(A, b)
    assert np.abs(np.tril(B, -b - 1)).max() == 0.0
    lam_ref = np.linalg.eigvalsh(A)
    print("band eig err", np.abs(np.linalg.eigvalsh(B) - lam_ref).max())
    d, e, Q2, Tm = sb2st(B, b)
    print("offband after chase", np.abs(np.tril(Tm, -2)).max())
    Tri = np.diag(d) + np.diag(e, -1) + np.diag(e, 1)
    lam, Zt = np.linalg.eigh(Tri)
    print("tri eig err", np.abs(lam - lam_ref).max())
    worst = 0.0
    for idx in (-1, -2, 0):
        x = apply_q1(Vrows, taus, apply_q2(Q2, Zt[:, idx], b), b)
        worst = max(worst, np.abs(A @ x - lam[idx] * x).max())
    print("eigvec residual", worst)


def sb2st_lanes(B, b=8):
    """Lane-level emulation of sb2st_kernel: band storage Bs[j][d] = B[j+d][j], d < 16, zero padded to N + 16 columns;
    per step three groups of 8 lanes hold 8 values each, addressed lo * 16 + off[m]."""
    n = B.shape[0]
    Bs = np.zeros((n + 16) * 16 + 256)
    for j in range(n):
        for d in range(min(9, n - j)):
            Bs[j * 16 + d] = B[j + d, j]
    Q2 = np.zeros((n, n))
    off = np.zeros((24, 8), dtype=int)
    for l in range(8):
        for m in range(8):
            off[l, m] = (l - 8) * 16 + 8 - l + m          # G0: column l of C_k, row m
            off[8 + l, m] = min(l, m) * 16 + abs(l - m)    # G1: row l of D_k (symmetric)
            off[16 + l, m] = m * 16 + 8 + l - m            # G2: row l of C_{k+1}
    for s in range(n - 2):
        k = 0
        while True:
            lo = s + 1 + 8 * k
            if n - lo < 2:
                break
            z = np.zeros((24, 8))
            for l in range(24):
                for m in range(8):
                    if l < 8 and k == 0:
                        z[l, m] = Bs[lo * 16 - 15 + m] if l == 0 else 0.0
                    else:
                        z[l, m] = Bs[lo * 16 + off[l, m]]
            x = z[0].copy()
            alpha, ss = x[0], float(x[1:] @ x[1:])
            v = np.zeros(8); v[0] = 1.0
            if ss > 0:
                beta = -np.copysign(np.sqrt(alpha * alpha + ss), alpha)
                tau = (beta - alpha) / beta
                v[1:] = x[1:] / (alpha - beta)
            else:
                beta, tau = alpha, 0.0
            for m in range(8):
                if lo + m < n:
                    Q2[s, lo - s - 1 + m] = tau if m == 0 else v[m]
            dot = z @ v
            # G0 / G2
            for l in list(range(8)) + list(range(16, 24)):
                z[l] -= tau * dot[l] * v
            z[0] = 0.0; z[0, 0] = beta
            # G1
            pvec = tau * dot[8:16]
            K = 0.5 * tau * (v @ pvec)
            wv = pvec - K * v
            for r in range(8):
                z[8 + r] -= v[r] * wv + wv[r] * v
            # stores
            for l in range(24):
                for m in range(8):
                    if l < 8:
                        if k == 0:
                            if l == 0:
                                Bs[lo * 16 - 15 + m] = z[l, m]
                        else:
                            Bs[lo * 16 + off[l, m]] = z[l, m]
                    elif l < 16:
                        if m <= l - 8:
                            Bs[lo * 16 + off[l, m]] = z[l, m]
                    else:
                        Bs[lo * 16 + off[l, m]] = z[l, m]
            k += 1
    assert np.abs(Bs[n * 16:]).max() == 0.0, "padding must stay zero"
    d = np.array([Bs[i * 16] for i in range(n)])
    e = np.array([Bs[i * 16 + 1] for i in range(n - 1)])
    return d, e, Q2


if __name__ == "__main__":
    d2, e2, Q2b = sb2st_lanes(B)
    Tri2 = np.diag(d2) + np.diag(e2, -1) + np.diag(e2, 1)
    lam2, Zt2 = np.linalg.eigh(Tri2)
    print("lane model: tri eig err", np.abs(lam2 - lam_ref).max(), "d/e diff vs dense model",
          np.abs(d2 - d).max(), np.abs(np.abs(e2) - np.abs(e)).max())
    x = apply_q1(Vrows, taus, apply_q2(Q2b, Zt2[:, -1], b), b)
    print("lane model eigvec residual", np.abs(A @ x - lam2[-1] * x).max())
