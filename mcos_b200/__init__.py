"""mcos_b200: B200-native engine for the MCOS Monte-Carlo loop, behind the reference's plugin API.

    from mcos_b200 import simulate_optimizations, MuCovLedoitWolfObservationSimulator, DeNoiserCovarianceTransformer, \
        MarkowitzOptimizer, HRPOptimizer, SharpeRatioErrorEstimator

Everything numerical runs in hand-written sm_100a CUDA behind the C ABI of include/mcosb.h (libmcosb200.so); there is no
CPU fallback.
"""
from .covariance_transformer import (AbstractCovarianceTransformer, DeNoiserCovarianceTransformer,
                                     DetoneCovarianceTransformer, corr_to_cov, cov_to_corr)
from .error_estimator import (AbstractErrorEstimator, ExpectedOutcomeErrorEstimator, SharpeRatioErrorEstimator,
                              VarianceErrorEstimator)
from .mcos import (build_plan, shard_range, simulate_errors, simulate_optimizations,
                   simulate_optimizations_from_price_history, status_counts, summarize_errors)
from .observation_simulator import (AbstractObservationSimulator, MuCovJackknifeObservationSimulator,
                                    MuCovLedoitWolfObservationSimulator, MuCovObservationSimulator)
from .optimizer import AbstractOptimizer, HRPOptimizer, MarkowitzOptimizer, NCOOptimizer, RiskParityOptimizer
from .utils import convert_price_history

__all__ = [n for n in dir() if not n.startswith("_")]
