"""Error estimators with the reference's class API (mcos/error_estimator.py:5-49), evaluated on the GPU (kernel K5)."""
from abc import ABC, abstractmethod

import numpy as np

from . import _lib
from .engine import get_engine


class AbstractErrorEstimator(ABC):
    """Measures how far an allocation is from the ideal one (mcos/error_estimator.py:5-18)."""

    @abstractmethod
    def estimate(self, mu: np.array, cov: np.array, allocation: np.array, optimal_allocation: np.array) -> float:
        pass


class _GpuErrorEstimator(AbstractErrorEstimator):
    kind = None

    def estimate(self, mu, cov, allocation, optimal_allocation, *, device: int = 0) -> float:
        cov = np.asarray(cov, dtype=np.float64)
        eng = get_engine(cov.shape[0], 1, device)
        eng.set_truth(mu, cov, require_pd=False)
        w = np.asarray(allocation, dtype=np.float64).reshape(1, -1)
        return float(eng.errors(w, np.asarray(optimal_allocation, dtype=np.float64).reshape(-1), self.kind)[0])

    def estimate_batch(self, mu, cov, allocations, optimal_allocation, *, device: int = 0) -> np.ndarray:
        """allocations [S, N] against one ideal allocation [N] (or [S, N])."""
        cov = np.asarray(cov, dtype=np.float64)
        eng = get_engine(cov.shape[0], 1, device)
        eng.set_truth(mu, cov, require_pd=False)
        return eng.errors(np.asarray(allocations, dtype=np.float64), np.asarray(optimal_allocation, dtype=np.float64),
                          self.kind)


class ExpectedOutcomeErrorEstimator(_GpuErrorEstimator):
    """Mean difference in expected outcomes (mcos/error_estimator.py:21-25)."""
    kind = _lib.ERR_EXPECTED_OUTCOME


class VarianceErrorEstimator(_GpuErrorEstimator):
    """Mean difference in variance (mcos/error_estimator.py:28-32)."""
    kind = _lib.ERR_VARIANCE


class SharpeRatioErrorEstimator(_GpuErrorEstimator):
    """Mean difference in Sharpe ratio (mcos/error_estimator.py:35-41); nan when the allocations coincide."""
    kind = _lib.ERR_SHARPE
