"""Optimizers with the reference's class API (mcos/optimizer.py) on the GPU (kernel families K4a-f)."""
from abc import ABC, abstractmethod

import numpy as np

from . import _lib
from .engine import get_engine


class AbstractOptimizer(ABC):

    @abstractmethod
    def allocate(self, mu: np.array, cov: np.array) -> np.array:
        pass

    @property
    @abstractmethod
    def name(self) -> str:
        pass


def _batch(mu, cov):
    cov = np.asarray(cov, dtype=np.float64)
    single = cov.ndim == 2
    cov = cov[None] if single else cov
    if mu is not None:
        mu = np.asarray(mu, dtype=np.float64).reshape(cov.shape[0], cov.shape[1])
    return mu, cov, single


class MarkowitzOptimizer(AbstractOptimizer):
    """Long-only max-Sharpe weights, risk-free rate 0.02, cleaned to 5 decimals -- what PyPortfolioOpt 0.5.2's
    EfficientFrontier(mu, cov).max_sharpe() / clean_weights() define (mcos/optimizer.py:41-53) -- solved exactly by an
    active-set QP instead of SLSQP."""

    def __init__(self, *, device: int = 0):
        self.device = device

    def allocate(self, mu: np.array, cov: np.array) -> np.array:
        """When no asset beats the risk-free rate the QP reformulation does not apply; the stated problem's maximum is
        then the single asset with the best stand-alone Sharpe ratio, which is what is returned (status NO_EXCESS; the
        reference's SLSQP also stops at a vertex there, a local one)."""
        return self.allocate_batch(mu, cov)[0]

    def allocate_batch(self, mu, cov, clean: bool = True, return_details: bool = False):
        mu, cov, _ = _batch(mu, cov)
        eng = get_engine(cov.shape[1], 1, self.device)
        w, info = eng.markowitz(mu, cov, 0.02, clean)
        return (w, info) if return_details else w

    @property
    def name(self) -> str:
        return 'markowitz'


class HRPOptimizer(AbstractOptimizer):
    """Hierarchical risk parity (mcos/optimizer.py:198-287). As in the reference, the returned vector is in
    quasi-diagonal POSITION order (entry k belongs to asset order[k]) and mu is ignored."""

    def __init__(self, *, device: int = 0):
        self.device = device

    def allocate(self, mu: np.array, cov: np.array) -> np.array:
        w, info = self.allocate_batch(cov, return_details=True)
        if info["status"][0] & _lib.ST_HRP_SUM:
            raise ValueError("Portfolio allocations don't sum to 1.")
        return w[0]

    def allocate_batch(self, cov, return_details: bool = False):
        _, cov, _ = _batch(None, cov)
        eng = get_engine(cov.shape[1], 1, self.device)
        w, info = eng.hrp(cov)
        return (w, info) if return_details else w

    @property
    def name(self) -> str:
        return 'HRP'


class NCOOptimizer(AbstractOptimizer):
    """Nested clustered optimization (mcos/optimizer.py:56-195)."""

    def __init__(self, max_num_clusters: int = None, num_clustering_trials=10, *, device: int = 0):
        self.max_num_clusters = max_num_clusters
        self.num_clustering_trials = num_clustering_trials
        self.device = device

    @property
    def name(self) -> str:
        return 'NCO'

    def allocate(self, mu: np.array, cov: np.array) -> np.array:
        return self.allocate_batch(mu, cov)[0]

    def allocate_batch(self, mu, cov, labels=None, return_details: bool = False):
        mu, cov, _ = _batch(mu, cov)
        if mu is not None:
            assert mu.shape[1] == cov.shape[1], 'mu and cov dimension must be the same size'
        eng = get_engine(cov.shape[1], 1, self.device)
        w, info = eng.nco(mu, cov, self.max_num_clusters, self.num_clustering_trials, labels)
        return (w, info) if return_details else w


class RiskParityOptimizer(AbstractOptimizer):
    """Risk parity / risk budgeting (mcos/optimizer.py:290-359).  The reference runs scipy SLSQP from the budget itself;
    the GPU kernel returns the budget when SLSQP's first convergence test would fire there (what the reference does on
    covariances of realistic scale) and the exact risk-budgeting weights otherwise (riskparity.cu)."""

    def __init__(self, target_risk: np.array = None, *, device: int = 0):
        self.target_risk = target_risk
        self.device = device

    def allocate(self, mu: np.array, cov: np.array) -> np.array:
        return self.allocate_batch(cov)[0]

    def allocate_batch(self, cov, return_details: bool = False):
        _, cov, _ = _batch(None, cov)
        eng = get_engine(cov.shape[1], 1, self.device)
        budget = None if self.target_risk is None else np.asarray(self.target_risk, dtype=np.float64).reshape(-1)
        w, info = eng.risk_parity(cov, budget)
        return (w, info) if return_details else w

    @property
    def name(self) -> str:
        return 'Risk Parity'
