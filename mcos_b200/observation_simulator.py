"""Observation simulators with the reference's class API (mcos/observation_simulator.py:6-57) on the GPU.

The reference draws from numpy's global MT19937 stream; here each simulator owns a counter-based Philox4x32-10 stream
keyed by `seed` (keyword-only extra), and simulation i always uses counter block i, so results do not depend on batch
size or GPU count.  `moments(x)` applies the estimator to an injected T x N draw (what the parity tests use).
"""
from abc import ABC, abstractmethod

import numpy as np

from . import _lib
from .engine import get_engine


class AbstractObservationSimulator(ABC):

    @abstractmethod
    def simulate(self) -> (np.array, np.array):
        """Draws an empirical mean vector (N, 1) and covariance matrix (N, N)."""
        pass


class _GpuObservationSimulator(AbstractObservationSimulator):
    kind = None

    def __init__(self, mu: np.array, cov: np.array, n_observations: int, *, seed: int = 0, device: int = 0):
        self.mu = mu
        self.cov = cov
        self.n_observations = n_observations
        self.seed, self.device = seed, device
        self._next_sim_id = 0

    def _engine(self):
        n = np.asarray(self.cov).shape[0]
        eng = get_engine(n, self.n_observations, self.device, self.seed)
        eng.ensure_truth(self.mu, self.cov)
        return eng

    def simulate(self) -> (np.array, np.array):
        mu_hat, cov_hat = self.simulate_batch(1)
        return mu_hat[0].reshape(-1, 1), cov_hat[0]

    def simulate_batch(self, n_sims: int, sim_id0: int = None):
        """(mu_hat [S, N], cov_hat [S, N, N]) for the next n_sims simulations."""
        eng = self._engine()
        if sim_id0 is None:
            sim_id0 = self._next_sim_id
            self._next_sim_id += n_sims
        mu_hat, cov_hat, _ = eng.sample_moments(n_sims, sim_id0, self.kind)   # the draws never leave the device
        return mu_hat, cov_hat

    def moments(self, x):
        """The estimator applied to injected draws x [T, N] or [S, T, N]."""
        x = np.asarray(x, dtype=np.float64)
        single = x.ndim == 2
        xb = x[None] if single else x
        eng = get_engine(xb.shape[2], xb.shape[1], self.device, self.seed)
        mu_hat, cov_hat, _ = eng.moments(xb, self.kind)
        return (mu_hat[0].reshape(-1, 1), cov_hat[0]) if single else (mu_hat, cov_hat)


class MuCovLedoitWolfObservationSimulator(_GpuObservationSimulator):
    """Sample mean + Ledoit-Wolf shrunk covariance (mcos/observation_simulator.py:19-28)."""
    kind = _lib.MOMENTS_LEDOIT_WOLF


class MuCovObservationSimulator(_GpuObservationSimulator):
    """Sample mean + sample covariance (mcos/observation_simulator.py:31-40)."""
    kind = _lib.MOMENTS_MUCOV


class MuCovJackknifeObservationSimulator(_GpuObservationSimulator):
    """Jackknifed mean and covariance (mcos/observation_simulator.py:43-57); algebraically the sample moments."""
    kind = _lib.MOMENTS_JACKKNIFE
