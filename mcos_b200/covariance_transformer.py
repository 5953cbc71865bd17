"""Covariance transformers with the reference's class API (mcos/covariance_transformer.py) on the GPU."""
from abc import ABC, abstractmethod

import numpy as np

from . import _lib
from .engine import get_engine


def cov_to_corr(cov: np.array) -> np.array:
    """Correlation matrix of a covariance matrix, clipped into [-1, 1] (mcos/covariance_transformer.py:9-18).
    Host-side helper for callers; the kernels compute it on the fly."""
    std = np.sqrt(np.diag(cov))
    corr = np.asarray(cov) / np.outer(std, std)
    return np.clip(corr, -1.0, 1.0)


def corr_to_cov(corr: np.array, std: np.array) -> np.array:
    """mcos/covariance_transformer.py:21-29."""
    return corr * np.outer(std, std)


class AbstractCovarianceTransformer(ABC):

    @abstractmethod
    def transform(self, cov: np.array, n_observations: int) -> np.array:
        pass


class DeNoiserCovarianceTransformer(AbstractCovarianceTransformer):
    """Marcenko-Pastur de-noising (mcos/covariance_transformer.py:55-208), kernel family K3."""

    def __init__(self, bandwidth: float = .25, *, device: int = 0):
        self.bandwidth = bandwidth
        self.device = device

    def transform(self, cov: np.array, n_observations: int) -> np.array:
        return self.transform_batch(np.asarray(cov, dtype=np.float64)[None], n_observations)[0]

    def transform_batch(self, cov, n_observations: int, return_details: bool = False):
        """cov [S, N, N] -> de-noised [S, N, N] (optionally with eigvals / var / n_facts / status)."""
        cov = np.asarray(cov, dtype=np.float64)
        eng = get_engine(cov.shape[1], 1, self.device)
        out, info = eng.denoise(cov, n_observations, self.bandwidth)
        return (out, info) if return_details else out


class DetoneCovarianceTransformer(AbstractCovarianceTransformer):
    """Removes the n_remove largest eigen-components of the correlation matrix -- the "market" factors -- renormalises
    and rescales (mcos/covariance_transformer.py:211-250).  GPU: shared eigen front end + detone_kernel; n_remove <= 16."""

    def __init__(self, n_remove: int, *, device: int = 0):
        self.n_remove = n_remove
        self.device = device

    def transform(self, cov: np.array, n_observations: int) -> np.array:
        if self.n_remove == 0:
            return cov
        return self.transform_batch(np.asarray(cov, dtype=np.float64)[None])[0]

    def transform_batch(self, cov):
        cov = np.asarray(cov, dtype=np.float64)
        if self.n_remove == 0:
            return cov
        eng = get_engine(cov.shape[1], 1, self.device)
        out, _ = eng.detone(cov, self.n_remove)
        return out
