"""Synthetic "true" (mu, cov) inputs for benchmarks and tests (host-side numpy; runs once before the loop).

The construction is the block-diagonal experiment of the paper MCOS implements (Lopez de Prado, "A Robust Estimator of
the Efficient Frontier", section 4): `n_blocks` diagonal blocks of equal size with intra-block correlation `rho`, assets
shuffled by a random permutation, sigma_i ~ U(0.05, 0.2), cov = corr * sigma sigma', mu_i ~ N(sigma_i, sigma_i).
All draws come from numpy.random.RandomState(seed) in the order (permutation, sigma, mu).  SURVEY.md section 8(d).
"""
import numpy as np


def block_diagonal_truth(n_assets: int, n_blocks: int = 10, rho: float = 0.5, seed: int = 0):
    if n_assets % n_blocks:
        raise ValueError("n_assets must be a multiple of n_blocks")
    rng = np.random.RandomState(seed)
    size = n_assets // n_blocks
    block = np.full((size, size), rho)
    np.fill_diagonal(block, 1.0)
    corr = np.kron(np.eye(n_blocks), block)
    perm = rng.permutation(n_assets)
    corr = corr[perm][:, perm]
    sigma = rng.uniform(0.05, 0.2, n_assets)
    cov = corr * np.outer(sigma, sigma)
    mu = rng.normal(sigma, sigma)
    return mu, cov
