"""CPU property test of the guarded Prim behind the HRP fast path, restated in numpy.

hrp_gram_kernel / hrp_prim_kernel decide scipy's single-linkage merge order from APPROXIMATE squared distances
(n_i + n_j - 2 D_i . D_j) and flag a simulation whenever two compared values are closer than 3 delta,
delta = 16 N u max_i |D_i|^2.  Claims checked here on random, clustered and tie-ridden inputs:
  1. |approximate - scipy's sequentially summed square| <= delta / 2 for every pair;
  2. an unflagged run takes exactly the decisions of scipy's mst_single_linkage on the exact distances (same edges, same
     tree node realising each minimum), so the exact heights of those N - 1 pairs reproduce the linkage;
  3. exact ties are always flagged.
"""
import numpy as np

U = 2.0 ** -53


def scipy_squares(d):
    """pdist's arithmetic: sequential over k, separate subtract, multiply, add (numpy elementwise ops do not fuse)."""
    n = d.shape[0]
    s = np.zeros((n, n))
    for k in range(d.shape[1]):
        diff = d[:, k][:, None] - d[:, k][None, :]
        s = s + diff * diff
    return s


def exact_prim(e):
    """scipy _hierarchy.mst_single_linkage on the square matrix of exact distances: (x, y, tree node realising D[y])."""
    n = e.shape[0]
    merged = np.zeros(n, bool)
    best = np.full(n, np.inf)
    bestx = np.zeros(n, int)
    x, edges = 0, []
    for _ in range(n - 1):
        merged[x] = True
        cur, y = np.inf, -1
        for i in range(n):
            if merged[i]:
                continue
            if best[i] > e[x, i]:
                best[i], bestx[i] = e[x, i], x
            if best[i] < cur:
                cur, y = best[i], i
        edges.append((x, y, bestx[y]))
        x = y
    return edges


def guarded_prim(a, delta3):
    """The fast path's Prim on approximate squares: returns (edges, flagged)."""
    n = a.shape[0]
    merged = np.zeros(n, bool)
    best = np.full(n, np.inf)
    bestx = np.zeros(n, int)
    x, edges = 0, []
    for _ in range(n - 1):
        merged[x] = True
        m1, m2, y = np.inf, np.inf, -1
        for i in range(n):
            if merged[i]:
                continue
            dd, b = a[x, i], best[i]
            if b > dd:
                if b - dd <= delta3:
                    return edges, True
                best[i], bestx[i] = dd, x
                b = dd
            elif dd - b <= delta3:
                return edges, True
            if b < m1:
                m2, m1, y = m1, b, i
            elif b < m2:
                m2 = b
        if not (m2 - m1 > delta3):
            return edges, True
        edges.append((x, y, bestx[y]))
        x = y
    return edges, False


def _case(rng, n, kind):
    if kind == "random":
        d = rng.uniform(0.0, 1.0, (n, n))
    elif kind == "clustered":           # rows that are nearly equal: small distances, large norms
        base = rng.uniform(0.3, 0.9, (4, n))
        d = base[rng.randint(0, 4, n)] + 1e-3 * rng.standard_normal((n, n))
    else:                                # exact duplicates: exact ties
        d = rng.uniform(0.0, 1.0, (n, n))
        d[n // 2:] = d[: n - n // 2]
    return d


def test_guarded_prim_equals_exact_prim_or_flags():
    rng = np.random.RandomState(0)
    unflagged = 0
    for trial in range(60):
        n = int(rng.randint(5, 48))
        kind = ("random", "clustered", "ties")[trial % 3]
        d = _case(rng, n, kind)
        s = scipy_squares(d)
        norms = (d * d).sum(1)
        a = norms[:, None] + norms[None, :] - 2.0 * (d @ d.T)
        np.fill_diagonal(a, 0.0)
        a = np.maximum(a, 0.0)
        delta = 16.0 * n * U * norms.max()
        assert np.abs(a - s).max() <= delta / 2, (kind, np.abs(a - s).max(), delta)      # claim 1
        edges, flagged = guarded_prim(a, 3.0 * delta)
        if kind == "ties":
            assert flagged                                                                # claim 3
            continue
        if not flagged:
            unflagged += 1
            assert edges == exact_prim(np.sqrt(s))                                        # claim 2
    assert unflagged >= 30


def test_near_ties_inside_the_bound_are_flagged_or_decided_correctly():
    """Adversarial: two candidates whose exact squares differ by a few ulps.  The guard must flag (it cannot know the
    order); with a gap far above the bound it must decide, and decide like the exact path."""
    rng = np.random.RandomState(1)
    for gap_ulps, must_flag in ((0, True), (3, True), (10 ** 7, False)):
        n = 12
        d = rng.uniform(0.0, 1.0, (n, n))
        d[2] = d[1]                                  # nodes 1 and 2 equidistant from everything ...
        d[2, 0] = np.nextafter(d[2, 0], 2.0) if gap_ulps == 0 else d[1, 0] + gap_ulps * np.spacing(d[1, 0])
        d[1] += 10.0                                 # ... and far from node 0, so that they are compared late
        d[2] += 10.0
        s = scipy_squares(d)
        norms = (d * d).sum(1)
        a = np.maximum(norms[:, None] + norms[None, :] - 2.0 * (d @ d.T), 0.0)
        np.fill_diagonal(a, 0.0)
        delta = 16.0 * n * U * norms.max()
        edges, flagged = guarded_prim(a, 3.0 * delta)
        if must_flag:
            assert flagged
        elif not flagged:
            assert edges == exact_prim(np.sqrt(s))
