"""CPU restatement of the MCOS hot path (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Every function cites the reference lines it follows (paths relative to the reference repo root).  The reference is pure
Python on numpy / scipy / scikit-learn / PyPortfolioOpt; this port keeps the same library calls where the library *is*
the algorithm (L-BFGS-B, SLSQP, KernelDensity, KMeans, linkage) and restates the rest as plain functions that take and
return arrays, so they can be applied to injected inputs one simulation at a time.

Third-party code the reference calls that is absent from the reference tree:
  * PyPortfolioOpt==0.5.2 (requirements.txt:5): `EfficientFrontier(mu, cov).max_sharpe()` / `.clean_weights()`
    (call site mcos/optimizer.py:44-49) and `expected_returns.mean_historical_return` / `risk_models.sample_cov`
    (mcos/utils.py:13-14).  Restated below from the published 0.5.2 behaviour; pinned by the goldens at
    tests/test_optimizer.py:17-22 and tests/test_utils.py:9-12 of the reference.
"""
from __future__ import annotations

import math
import warnings

import numpy as np

__all__ = [
    "draw_observations", "moments_mucov", "moments_ledoit_wolf", "ledoit_wolf_formula", "moments_jackknife",
    "cov_to_corr", "corr_to_cov", "mp_pdf", "kde_gauss", "mp_fit_objective", "find_max_eigenvalue", "denoise",
    "denoise_details", "detone", "markowitz_slsqp", "clean_weights", "markowitz_reference", "markowitz_exact_qp",
    "markowitz_exact", "correlation_distance", "single_linkage", "quasi_diagonal_order", "hrp_bisection", "hrp",
    "hrp_details", "nco_cluster", "nco_allocate", "nco", "closed_form_portfolio", "risk_parity",
    "err_expected_outcome", "err_variance", "err_sharpe", "estimate_error", "mean_historical_return", "sample_cov",
    "convert_price_history", "simulate_optimizations", "SIMULATORS", "ESTIMATORS",
]


# ----------------------------------------------------------------------------------------------------------------
# observation simulators  (mcos/observation_simulator.py)
# ----------------------------------------------------------------------------------------------------------------

def draw_observations(mu, cov, n_observations, rng=None):
    """T x N Gaussian draws exactly as the reference makes them: the legacy global-state
    `np.random.multivariate_normal(mu.flatten(), cov, size=T)` (observation_simulator.py:27,39,51)."""
    rng = np.random if rng is None else rng
    return rng.multivariate_normal(np.asarray(mu).flatten(), cov, size=n_observations)


def moments_mucov(x):
    """MuCovObservationSimulator.simulate on a given draw (observation_simulator.py:38-40)."""
    return x.mean(axis=0).reshape(-1, 1), np.cov(x, rowvar=False)


def moments_ledoit_wolf(x):
    """MuCovLedoitWolfObservationSimulator.simulate on a given draw (observation_simulator.py:26-28)."""
    from sklearn.covariance import LedoitWolf
    return x.mean(axis=0).reshape(-1, 1), LedoitWolf().fit(x).covariance_


def ledoit_wolf_formula(x):
    """Closed form of what sklearn's LedoitWolf().fit(x) computes (sklearn/covariance/_shrunk_covariance.py,
    `ledoit_wolf_shrinkage` + `ledoit_wolf`), used to cross-check the CUDA epilogue's intermediate quantities.
    Returns (cov_hat, shrinkage)."""
    t, n = x.shape
    xc = x - x.mean(axis=0)
    s_b = xc.T @ xc / t                                  # biased sample covariance
    m = np.trace(s_b) / n
    beta_raw = float(((xc ** 2).sum(axis=1) ** 2).sum())  # == sum_ij (Xc^2' Xc^2)_ij
    delta_raw = float((s_b ** 2).sum())                   # ||S_b||_F^2
    beta = (beta_raw / t - delta_raw) / (n * t)
    delta = (delta_raw - n * m * m) / n
    beta = min(beta, delta)
    shrink = 0.0 if beta == 0 else beta / delta
    out = (1.0 - shrink) * s_b
    out[np.diag_indices(n)] += shrink * m
    return out, shrink


def moments_jackknife(x):
    """MuCovJackknifeObservationSimulator.simulate on a given draw (observation_simulator.py:50-57): the mean over i of
    the leave-one-out covariances, and the mean of the summed leave-one-out sample blocks.  O(T) calls of np.cov, as
    the reference does it."""
    t = len(x)
    keep = np.arange(t)
    cov_acc = None
    blk_acc = None
    for i in range(t):
        part = x[keep != i]
        c = np.cov(part, rowvar=False)
        cov_acc = c if cov_acc is None else cov_acc + c
        blk_acc = part.copy() if blk_acc is None else blk_acc + part
    cov_hat = cov_acc / float(t)
    x_prime = blk_acc / float(t)
    return x_prime.mean(axis=0).reshape(-1, 1), cov_hat


SIMULATORS = {"mucov": moments_mucov, "mucovledoitwolf": moments_ledoit_wolf, "jackknife": moments_jackknife}


# ----------------------------------------------------------------------------------------------------------------
# covariance transformers  (mcos/covariance_transformer.py)
# ----------------------------------------------------------------------------------------------------------------

def cov_to_corr(cov):
    """covariance_transformer.py:9-18 (entries clipped into [-1, 1])."""
    sd = np.sqrt(np.diag(cov))
    corr = np.asarray(cov) / np.outer(sd, sd)
    np.clip(corr, -1.0, 1.0, out=corr)
    return corr


def corr_to_cov(corr, std):
    """covariance_transformer.py:21-29."""
    return corr * np.outer(std, std)


def mp_pdf(var, q, pts):
    """Marcenko-Pastur density on a `pts`-point grid spanning its support (covariance_transformer.py:153-167).
    Returns (grid, pdf)."""
    var = float(np.ravel(var)[0])
    lo = var * (1.0 - (1.0 / q) ** 0.5) ** 2
    hi = var * (1.0 + (1.0 / q) ** 0.5) ** 2
    grid = np.linspace(lo, hi, pts)
    pdf = q / (2.0 * np.pi * var * grid) * ((hi - grid) * (grid - lo)) ** 0.5
    return grid, pdf


def kde_gauss(obs, bandwidth, x):
    """Gaussian KDE of `obs` evaluated at `x` through sklearn's KernelDensity, as the reference does
    (covariance_transformer.py:169-192)."""
    from sklearn.neighbors import KernelDensity
    kde = KernelDensity(kernel="gaussian", bandwidth=bandwidth).fit(np.asarray(obs).reshape(-1, 1))
    return np.exp(kde.score_samples(np.asarray(x).reshape(-1, 1)))


def mp_fit_objective(var, eigvals, q, bandwidth=0.25, pts=1000):
    """Sum of squared differences between the MP density and the KDE of the eigenvalues on the MP grid
    (covariance_transformer.py:134-151)."""
    grid, pdf = mp_pdf(var, q, pts)
    emp = kde_gauss(eigvals, bandwidth, grid)
    return float(np.sum((emp - pdf) ** 2))


def find_max_eigenvalue(eigvals, q, bandwidth=0.25):
    """covariance_transformer.py:111-132.  scipy's default method with bounds is L-BFGS-B with a 2-point
    finite-difference gradient.  Returns (lambda_plus, var, success)."""
    from scipy.optimize import minimize
    out = minimize(lambda v: mp_fit_objective(v, eigvals, q, bandwidth), 0.5, bounds=((1e-5, 1.0 - 1e-5),))
    var = float(out["x"][0]) if out["success"] else 1.0
    return var * (1.0 + (1.0 / q) ** 0.5) ** 2, var, bool(out["success"])


def denoise_details(cov, n_observations, bandwidth=0.25):
    """DeNoiserCovarianceTransformer.transform with its intermediates exposed (covariance_transformer.py:62-97,
    99-109, 194-208)."""
    cov = np.asarray(cov, dtype=float)
    n = cov.shape[1]
    q = n_observations / n
    corr = cov_to_corr(cov)
    lam, vec = np.linalg.eigh(corr)
    desc = lam.argsort()[::-1]
    lam, vec = lam[desc], vec[:, desc]
    lam_plus, var, ok = find_max_eigenvalue(lam, q, bandwidth)
    n_facts = n - int(lam[::-1].searchsorted(lam_plus))
    clipped = lam.copy()
    clipped[n_facts:] = clipped[n_facts:].sum() / float(n - n_facts)
    corr1 = cov_to_corr((vec * clipped) @ vec.T)
    out = corr_to_cov(corr1, np.diag(cov) ** 0.5)
    return {"cov": out, "eigvals": lam, "eigvecs": vec, "var": var, "lambda_plus": lam_plus, "n_facts": n_facts,
            "success": ok, "corr_denoised": corr1}


def denoise(cov, n_observations, bandwidth=0.25):
    return denoise_details(cov, n_observations, bandwidth)["cov"]


def detone(cov, n_remove):
    """DetoneCovarianceTransformer.transform (covariance_transformer.py:222-250)."""
    if n_remove == 0:
        return cov
    corr = cov_to_corr(cov)
    w, v = np.linalg.eig(corr)
    big_first = np.argsort(-np.abs(w))
    w, v = w[big_first], v[:, big_first]
    vm, wm = v[:, :n_remove], w[:n_remove]
    market = (vm * wm) @ vm.T
    c2 = corr - market
    scale = np.diag(c2) ** -0.5
    c2 = c2 * np.outer(scale, scale)
    return corr_to_cov(c2, np.diag(cov) ** 0.5)


# ----------------------------------------------------------------------------------------------------------------
# Markowitz  (mcos/optimizer.py:41-53 + PyPortfolioOpt 0.5.2)
# ----------------------------------------------------------------------------------------------------------------

def _negative_sharpe(w, mu, cov, gamma, rf):
    # PyPortfolioOpt 0.5.2 objective_functions.negative_sharpe
    ret = w.dot(mu)
    sigma = np.sqrt(np.dot(w, np.dot(cov, w.T)))
    return -(ret - rf) / sigma + gamma * (w ** 2).sum()


def markowitz_slsqp(mu, cov, rf=0.02, gamma=0.0):
    """EfficientFrontier(mu, cov, weight_bounds=(0, 1)).max_sharpe(risk_free_rate=0.02) of PyPortfolioOpt 0.5.2:
    scipy SLSQP from x0 = 1/N with bounds (0, 1) and the single equality sum(w) = 1; default ftol 1e-6, numerical
    gradients, no `success` check.  `mu` is passed through unflattened, as the reference does (it may be (N, 1))."""
    from scipy.optimize import minimize
    n = len(mu)
    res = minimize(_negative_sharpe, x0=np.array([1.0 / n] * n), args=(mu, cov, gamma, rf), method="SLSQP",
                   bounds=((0, 1),) * n, constraints=[{"type": "eq", "fun": lambda x: np.sum(x) - 1}])
    return res["x"], res


def clean_weights(w, cutoff=1e-4, rounding=5):
    """EfficientFrontier.clean_weights of 0.5.2: zero |w| < cutoff, round to 5 decimals, no renormalisation."""
    out = np.array(w, dtype=float, copy=True)
    out[np.abs(out) < cutoff] = 0.0
    return np.round(out, rounding) if rounding is not None else out


def markowitz_reference(mu, cov):
    """MarkowitzOptimizer.allocate as the reference runs it (optimizer.py:44-49)."""
    return clean_weights(markowitz_slsqp(mu, cov)[0])


def markowitz_exact_qp(mu, cov, rf=0.02, tol=1e-12, max_iter=None):
    """Exact solution of the problem SLSQP approximates.  When some mu_i > rf the long-only max-Sharpe problem
    max (w.mu - rf)/sqrt(w'Cw), sum w = 1, 0 <= w <= 1 is equivalent to the convex QP
        min y'Cy   s.t.  (mu - rf)'y = 1,  y >= 0,     w = y / sum(y)
    solved here with a primal active-set method (one index enters or leaves per iteration).  Returns (w, info) with
    info = dict(free=sorted free set, iters=...).  SURVEY.md Appendix A.5."""
    c = np.asarray(mu, dtype=float).reshape(-1) - rf
    cov = np.asarray(cov, dtype=float)
    n = c.size
    if not np.any(c > 0):
        raise ValueError("no asset has an expected return above the risk-free rate")
    max_iter = 50 * n if max_iter is None else max_iter
    # start at the single-asset vertex with the best stand-alone Sharpe ratio
    start = int(np.argmax(np.where(c > 0, c / np.sqrt(np.diag(cov)), -np.inf)))
    free = [start]
    y = np.zeros(n)
    y[start] = 1.0 / c[start]
    iters = 0
    while True:
        iters += 1
        if iters > max_iter:
            raise RuntimeError("active-set iteration cap")
        f = np.array(free)
        z = np.linalg.solve(cov[np.ix_(f, f)], c[f])
        target = z / c[f].dot(z)
        if np.all(target > 0):
            y[:] = 0.0
            y[f] = target
            grad = cov[:, f] @ target                     # (C y)_i
            nu = target.dot(grad[f])                      # y'Cy  (multiplier of the equality, c'y = 1)
            lam = grad - nu * c                           # bound multipliers; must be >= 0 off the free set
            lam[f] = 0.0
            worst = int(np.argmin(lam))
            if lam[worst] >= -tol * max(1.0, abs(nu)):
                break
            free.append(worst)
        else:
            cur = y[f]
            step = target - cur
            ratios = np.where(step < 0, cur / np.where(step < 0, -step, 1.0), np.inf)
            k = int(np.argmin(ratios))
            alpha = min(1.0, ratios[k])
            y[f] = cur + alpha * step
            y[f[k]] = 0.0
            free.pop(k)
    w = y / y.sum()
    return w, {"free": sorted(free), "iters": iters}


def markowitz_exact(mu, cov):
    """Exact-QP weights passed through the reference's clean_weights step."""
    return clean_weights(markowitz_exact_qp(mu, cov)[0])


# ----------------------------------------------------------------------------------------------------------------
# HRP  (mcos/optimizer.py:198-287)
# ----------------------------------------------------------------------------------------------------------------

def correlation_distance(corr):
    """optimizer.py:281-287: sqrt((1 - corr) / 2) with the diagonal forced to zero."""
    d = np.sqrt((1.0 - corr) / 2.0)
    np.fill_diagonal(d, 0.0)
    return d


def _row_euclid(points):
    """Pairwise Euclidean distance between the rows of `points`, accumulated sequentially over the coordinate index
    with separate multiply and add -- bit-identical to scipy.spatial.distance.pdist(..., 'euclidean').
    Returns the full symmetric matrix."""
    n = points.shape[0]
    acc = np.zeros((n, n))
    for k in range(points.shape[1]):
        col = points[:, k]
        diff = col[:, None] - col[None, :]
        acc = acc + diff * diff
    return np.sqrt(acc)


def single_linkage(euclid):
    """Restatement of scipy.cluster.hierarchy.linkage(.., 'single') on precomputed pairwise distances (scipy
    `_hierarchy.mst_single_linkage` + `label`): Prim's MST grown from node 0 with strict-< arg-min (lowest index wins
    ties), a stable sort of the n-1 edges by length, then union-find relabelling that writes (smaller root, larger
    root, length, merged size) per row and names the merged cluster n + row.  SURVEY.md Appendix A.6."""
    n = euclid.shape[0]
    in_tree = np.zeros(n, dtype=bool)
    best = np.full(n, np.inf)
    edges = np.empty((n - 1, 3))
    x = 0
    for k in range(n - 1):
        in_tree[x] = True
        cur_min, y = np.inf, -1
        row = euclid[x]
        for i in range(n):
            if in_tree[i]:
                continue
            if best[i] > row[i]:
                best[i] = row[i]
            if best[i] < cur_min:
                cur_min, y = best[i], i
        edges[k] = (x, y, cur_min)
        x = y
    edges = edges[np.argsort(edges[:, 2], kind="mergesort")]
    parent = np.arange(2 * n - 1)
    size = np.ones(2 * n - 1, dtype=int)

    def find(a):
        root = a
        while parent[root] != root:
            root = parent[root]
        while parent[a] != root:
            parent[a], a = root, parent[a]
        return root

    link = np.empty((n - 1, 4))
    for r in range(n - 1):
        a, b = find(int(edges[r, 0])), find(int(edges[r, 1]))
        lo, hi = (a, b) if a < b else (b, a)
        new = n + r
        parent[a] = parent[b] = new
        size[new] = size[a] + size[b]
        link[r] = (lo, hi, edges[r, 2], size[new])
    return link


def quasi_diagonal_order(link):
    """optimizer.py:235-249: depth-first leaf order of the dendrogram from the last merge, column-0 child first."""
    n = int(link[-1, 3])
    stack = [n + link.shape[0] - 1]
    order = []
    while stack:
        node = stack.pop()
        if node < n:
            order.append(node)
        else:
            stack.append(int(link[node - n, 1]))
            stack.append(int(link[node - n, 0]))
    return order


def _ivp_cluster_var(block):
    ivp = 1.0 / np.diag(block)
    ivp /= ivp.sum()
    w = ivp.reshape(-1, 1)
    return np.dot(np.dot(w.T, block), w)[0, 0]


def hrp_bisection(cov, order):
    """optimizer.py:256-279: recursive bisection; `np.array_split(order, 2)` gives the first half the extra item.
    The result is in *position* order: entry k belongs to asset order[k] (the reference never un-permutes)."""
    order = np.asarray(order)
    if order.size == 0:
        raise ValueError("sorted_indices is empty")
    if order.size == 1:
        return np.array([1.0])
    left, right = np.array_split(order, 2)
    lv = _ivp_cluster_var(cov[:, left][left])
    rv = _ivp_cluster_var(cov[:, right][right])
    alpha = 1.0 - lv / (lv + rv)
    return np.concatenate([hrp_bisection(cov, left) * alpha, hrp_bisection(cov, right) * (1.0 - alpha)])


def hrp_details(cov, use_scipy=True):
    """HRPOptimizer.allocate with intermediates (optimizer.py:204-223).  `linkage` is given the *square* distance
    matrix, so scipy clusters on the Euclidean distance between its rows."""
    cov = np.asarray(cov, dtype=float)
    dist = correlation_distance(cov_to_corr(cov))
    if use_scipy:
        import scipy.cluster.hierarchy as sch
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            link = sch.linkage(dist, "single")
    else:
        link = single_linkage(_row_euclid(dist))
    order = quasi_diagonal_order(link)
    w = hrp_bisection(cov, order)
    if w.sum() > 1.001 or w.sum() < 0.999:
        raise ValueError("Portfolio allocations don't sum to 1.")
    return {"weights": w, "order": np.array(order), "linkage": link, "dist": dist}


def hrp(cov):
    return hrp_details(cov)["weights"]


# ----------------------------------------------------------------------------------------------------------------
# NCO  (mcos/optimizer.py:56-195)
# ----------------------------------------------------------------------------------------------------------------

def closed_form_portfolio(cov, mu=None):
    """optimizer.py:176-195: inv(cov) mu / 1' inv(cov) mu, mu := 1 when None; pinv when singular."""
    cov = np.asarray(cov, dtype=float)
    try:
        inv = np.linalg.inv(cov)
    except np.linalg.LinAlgError:
        inv = np.linalg.pinv(cov)
    ones = np.ones((inv.shape[0], 1))
    rhs = ones if mu is None else np.asarray(mu, dtype=float).reshape(-1, 1)
    w = inv @ rhs
    w = w / (ones.T @ w)
    return w.flatten()


def nco_cluster(corr, max_num_clusters=None, num_clustering_trials=10, random_state=42):
    """optimizer.py:140-174 restated for the installed scikit-learn (the reference's `KMeans(n_jobs=1, ...)` call does
    not run on sklearn >= 1.0).  PARITY UNPINNED against the reference's sklearn 0.22.1 k-means++ stream.
    Returns (labels, silhouette quality)."""
    from sklearn.cluster import KMeans
    from sklearn.metrics import silhouette_samples
    corr = np.nan_to_num(np.asarray(corr, dtype=float), nan=0.0)
    dist = ((1.0 - corr) / 2.0) ** 0.5
    max_k = corr.shape[0] // 2 if max_num_clusters is None else max_num_clusters
    best_q, best_labels = np.nan, None
    for _ in range(num_clustering_trials):
        for k in range(2, max_k + 1):
            km = KMeans(n_clusters=k, n_init=1, random_state=random_state).fit(dist)
            sil = silhouette_samples(dist, km.labels_)
            qual = sil.mean() / sil.std()
            if np.isnan(best_q) or qual > best_q:
                best_q, best_labels = qual, km.labels_.copy()
    return best_labels, best_q


def nco_allocate(mu, cov, labels):
    """optimizer.py:119-138 given the partition: intra-cluster closed-form portfolios, reduced problem, product."""
    cov = np.asarray(cov, dtype=float)
    n = cov.shape[0]
    mu_flat = None if mu is None else np.asarray(mu, dtype=float).flatten()
    if mu_flat is not None:
        assert mu_flat.size == n, "mu and cov dimension must be the same size"
    ids = np.unique(labels)
    intra = np.zeros((n, ids.size))
    for col, cid in enumerate(ids):
        members = np.where(labels == cid)[0]
        sub_mu = None if mu_flat is None else mu_flat[members]
        intra[members, col] = closed_form_portfolio(cov[np.ix_(members, members)], sub_mu)
    red_cov = intra.T @ (cov @ intra)
    red_mu = None if mu_flat is None else intra.T @ mu_flat
    inter = closed_form_portfolio(red_cov, red_mu)
    return (intra * inter).sum(axis=1)


def nco(mu, cov, max_num_clusters=None, num_clustering_trials=10, labels=None):
    """NCOOptimizer.allocate (optimizer.py:74-138)."""
    cov = np.asarray(cov, dtype=float)
    if labels is None:
        labels, _ = nco_cluster(cov_to_corr(cov), max_num_clusters, num_clustering_trials)
    return nco_allocate(mu, cov, np.asarray(labels))


# ----------------------------------------------------------------------------------------------------------------
# Risk parity  (mcos/optimizer.py:290-359)  -- SURVEY 8(f) "next"
# ----------------------------------------------------------------------------------------------------------------

def risk_parity(cov, target_risk=None):
    from scipy.optimize import minimize
    cov = np.asarray(cov, dtype=float)
    n = len(cov[0])
    budget = np.array([1.0 / n] * n if target_risk is None else target_risk, dtype=float)

    def objective(x):
        w = np.array(x, ndmin=2)
        sig = np.sqrt((w @ cov @ w.T)[0, 0])
        rc = np.multiply(cov @ w.T, w.T) / sig
        return float(np.sum(np.square(rc - (sig * budget).reshape(-1, 1))))

    cons = ({"type": "eq", "fun": lambda x: np.sum(x) - 1.0}, {"type": "ineq", "fun": lambda x: x})
    res = minimize(objective, budget, method="SLSQP", constraints=cons, options={"disp": False})
    return np.array(res.x)


# ----------------------------------------------------------------------------------------------------------------
# error estimators  (mcos/error_estimator.py)
# ----------------------------------------------------------------------------------------------------------------

def err_expected_outcome(mu, cov, allocation, optimal_allocation):
    """error_estimator.py:24-25, 44-45."""
    return np.dot(optimal_allocation - allocation, mu)


def err_variance(mu, cov, allocation, optimal_allocation):
    """error_estimator.py:31-32, 48-49."""
    d = optimal_allocation - allocation
    return np.dot(d, np.dot(cov, d))


def err_sharpe(mu, cov, allocation, optimal_allocation):
    """error_estimator.py:38-41."""
    return err_expected_outcome(mu, cov, allocation, optimal_allocation) / \
        np.sqrt(err_variance(mu, cov, allocation, optimal_allocation))


ESTIMATORS = {"expected_outcome": err_expected_outcome, "variance": err_variance, "sharpe": err_sharpe}


def estimate_error(kind, mu, cov, allocation, optimal_allocation):
    return ESTIMATORS[kind](mu, cov, allocation, optimal_allocation)


# ----------------------------------------------------------------------------------------------------------------
# price-history adaptor  (mcos/utils.py:6-20 + PyPortfolioOpt 0.5.2 helpers)
# ----------------------------------------------------------------------------------------------------------------

def mean_historical_return(prices, frequency=252):
    """pypfopt.expected_returns.mean_historical_return (0.5.2): daily pct_change, drop all-NaN rows, mean * 252."""
    daily = prices.pct_change(fill_method=None).dropna(how="all")
    return daily.mean() * frequency


def sample_cov(prices, frequency=252):
    """pypfopt.risk_models.sample_cov (0.5.2): pairwise-complete covariance of daily returns * 252."""
    daily = prices.pct_change(fill_method=None).dropna(how="all")
    return daily.cov() * frequency


def convert_price_history(prices):
    """utils.py:6-20: (mu as a 1-D array, cov as a 2-D array)."""
    return mean_historical_return(prices).to_numpy(), sample_cov(prices).to_numpy()


# ----------------------------------------------------------------------------------------------------------------
# the Monte-Carlo loop  (mcos/mcos.py:13-41)
# ----------------------------------------------------------------------------------------------------------------

def _allocate(name, mu, cov, nco_args):
    if name == "markowitz":
        return markowitz_reference(mu, cov)
    if name == "markowitz_exact":
        return markowitz_exact(mu, cov)
    if name == "HRP":
        return hrp(cov)
    if name == "NCO":
        return nco(mu, cov, *nco_args)
    if name == "Risk Parity":
        return risk_parity(cov)
    raise ValueError(name)


def simulate_optimizations(mu, cov, n_observations, n_sims, simulator="mucov", transformers=(), optimizers=("markowitz",),
                           estimator="variance", nco_args=(None, 10), observations=None, return_samples=False):
    """The loop of mcos.py:22-33, including its recomputation of the ideal allocation every iteration, followed by the
    mean / population-stdev table of mcos.py:35-41 (returned as {name: (mean, stdev)}).

    `transformers` is a sequence of ("denoise", bandwidth) / ("detone", n_remove).  `observations`, if given, is a
    sequence of injected T x N draws used instead of the global numpy RNG."""
    moments = SIMULATORS[simulator.lower()]
    errors = {name: [] for name in optimizers}
    for i in range(n_sims):
        x = draw_observations(mu, cov, n_observations) if observations is None else observations[i]
        mu_hat, cov_hat = moments(x)
        for kind, arg in transformers:
            cov_hat = denoise(cov_hat, n_observations, arg) if kind == "denoise" else detone(cov_hat, arg)
        for name in optimizers:
            alloc = _allocate(name, mu_hat, cov_hat, nco_args)
            ideal = _allocate(name, mu, cov, nco_args)
            errors[name].append(estimate_error(estimator, mu, cov, alloc, ideal))
    table = {name: (float(np.mean(errors[name])), float(np.std(errors[name]))) for name in optimizers}
    return (table, errors) if return_samples else table
