"""Prototype: the Gaussian KDE sums of the MP fit (s0 = sum E_i, kd = sum E_i (x - lam_i)) from Taylor expansions about a
uniform set of nodes (spacing h / 8, degree DEG) instead of exp-recurrences over the 1000-point grid of every objective
evaluation.  Prints the worst error relative to max |s0|."""
import numpy as np, sys
h = 0.25; DEG = int(sys.argv[1]) if len(sys.argv) > 1 else 10; DIV = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rng = np.random.RandomState(0)
n, q = 500, 2.0
lam = np.concatenate([rng.uniform(0.05, 2.9, n - 10), rng.uniform(5, 26, 10)])
top = (1 + np.sqrt(1 / q)) ** 2
delta = h / DIV
nodes = np.arange(0, int(np.ceil(top / delta)) + 1) * delta
u = (nodes[:, None] - lam[None, :]) / h
E = np.exp(-0.5 * u * u)
C = np.zeros((len(nodes), DEG + 1))
cm1 = np.zeros_like(u); c = np.ones_like(u)
for m in range(DEG + 1):
    C[:, m] = (c * E).sum(1)
    cm1, c = c, (-u * c - cm1) / (m + 1)
def direct(x):
    d = x[:, None] - lam[None, :]
    e = np.exp(-d * d / (2 * h * h))
    return e.sum(1), (e * d).sum(1)
worst0 = worst1 = 0
for v in (1e-5, 0.1, 0.37, 0.5, 0.77, 1 - 1e-5):
    s = np.sqrt(1 / q)
    x = np.linspace(v * (1 - s) ** 2, v * (1 + s) ** 2, 1000)
    g = np.rint(x / delta).astype(int)
    t = (x - nodes[g]) / h
    s0 = np.zeros_like(x); d1 = np.zeros_like(x)
    for m in range(DEG, -1, -1):
        d1 = d1 * t + s0          # derivative Horner (w.r.t. t)
        s0 = s0 * t + C[g, m]
    kd = -h * d1                  # sum E (x - lam) = -h^2 f'(x) = -h * df/dt
    r0, r1 = direct(x)
    worst0 = max(worst0, np.abs(s0 - r0).max() / np.abs(r0).max()); worst1 = max(worst1, np.abs(kd - r1).max() / np.abs(r1).max())
print(f"nodes {len(nodes)} degree {DEG}: s0 rel err {worst0:.2e}, kd rel err {worst1:.2e}")
