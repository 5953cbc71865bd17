"""The Monte-Carlo driver with the reference's entry points (mcos/mcos.py:13-64), batched over all simulations.

The reference loops `for i in range(n_sims)` in Python and calls the plugins one simulation at a time.  Here the plugin
list is compiled into a stage plan and the whole loop body runs on the GPU through mcosb_run (sampling -> moments ->
transformers -> optimizers -> error estimators), in waves.  With torch.distributed initialised (one process per GPU)
the simulation ids are split into contiguous ranges per rank -- there is no data-path collective -- and the per-simulation
errors are gathered once at the end so the mean / stdev table is computed in simulation-id order on every rank.
"""
from typing import List

import numpy as np
import pandas as pd

from . import _lib
from .covariance_transformer import AbstractCovarianceTransformer, DeNoiserCovarianceTransformer, \
    DetoneCovarianceTransformer
from .engine import get_engine
from .error_estimator import AbstractErrorEstimator, ExpectedOutcomeErrorEstimator, SharpeRatioErrorEstimator, \
    VarianceErrorEstimator
from .observation_simulator import AbstractObservationSimulator, MuCovJackknifeObservationSimulator, \
    MuCovLedoitWolfObservationSimulator, MuCovObservationSimulator
from .optimizer import AbstractOptimizer, HRPOptimizer, MarkowitzOptimizer, NCOOptimizer, RiskParityOptimizer
from .utils import convert_price_history

_SIM_KIND = {MuCovObservationSimulator: _lib.MOMENTS_MUCOV, MuCovLedoitWolfObservationSimulator: _lib.MOMENTS_LEDOIT_WOLF,
             MuCovJackknifeObservationSimulator: _lib.MOMENTS_JACKKNIFE}
_OPT_KIND = {MarkowitzOptimizer: _lib.OPT_MARKOWITZ, HRPOptimizer: _lib.OPT_HRP, NCOOptimizer: _lib.OPT_NCO,
             RiskParityOptimizer: _lib.OPT_RISK_PARITY}
_ERR_KIND = {ExpectedOutcomeErrorEstimator: _lib.ERR_EXPECTED_OUTCOME, VarianceErrorEstimator: _lib.ERR_VARIANCE,
             SharpeRatioErrorEstimator: _lib.ERR_SHARPE}


def shard_range(n_sims: int, rank: int, world_size: int):
    """Contiguous simulation-id range [lo, hi) of `rank`: the first n_sims % world_size ranks get one extra."""
    base, extra = divmod(n_sims, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def build_plan(obs_simulator, optimizers, error_estimator, covariance_transformers, wave: int = 0) -> _lib.Plan:
    """Compiles the plugin objects into the stage plan mcosb_run executes. Unknown plugin types raise TypeError: there
    is no per-simulation Python fallback."""
    if type(obs_simulator) not in _SIM_KIND:
        raise TypeError(f"no GPU stage for observation simulator {type(obs_simulator).__name__}")
    if type(error_estimator) not in _ERR_KIND:
        raise TypeError(f"no GPU stage for error estimator {type(error_estimator).__name__}")
    plan = _lib.Plan()
    plan.moments_kind = _SIM_KIND[type(obs_simulator)]
    plan.denoise, plan.bandwidth, plan.detone_n_remove = 0, 0.25, 0
    for tr in covariance_transformers:
        if type(tr) is DeNoiserCovarianceTransformer:
            if plan.denoise or plan.detone_n_remove:
                raise TypeError("the plan supports [DeNoiser], [Detone] or [DeNoiser, Detone], in that order")
            plan.denoise, plan.bandwidth = 1, float(tr.bandwidth)
        elif type(tr) is DetoneCovarianceTransformer:
            if tr.n_remove == 0:
                continue  # the identity (covariance_transformer.py:223-224)
            if plan.detone_n_remove:
                raise TypeError("at most one DetoneCovarianceTransformer is supported in a plan")
            plan.detone_n_remove = int(tr.n_remove)
        else:
            raise TypeError(f"no GPU stage for covariance transformer {type(tr).__name__}")
    if not 1 <= len(optimizers) <= 4:
        raise ValueError("between 1 and 4 optimizers are supported")
    plan.n_optimizers = len(optimizers)
    plan.nco_max_clusters, plan.nco_trials = 0, 10
    for i, opt in enumerate(optimizers):
        if type(opt) not in _OPT_KIND:
            raise TypeError(f"no GPU stage for optimizer {type(opt).__name__}")
        plan.optimizers[i] = _OPT_KIND[type(opt)]
        if type(opt) is RiskParityOptimizer and opt.target_risk is not None:
            raise TypeError("the fused loop supports RiskParityOptimizer with the default equal risk budget only")
        if type(opt) is NCOOptimizer:
            plan.nco_max_clusters = 0 if opt.max_num_clusters is None else int(opt.max_num_clusters)
            plan.nco_trials = int(opt.num_clustering_trials)
    plan.error_kind = _ERR_KIND[type(error_estimator)]
    plan.risk_free_rate = 0.02
    plan.wave = int(wave)
    return plan


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


def gather_errors(local: np.ndarray, n_sims: int, device=None) -> np.ndarray:
    """All ranks' [n_local, n_opt] error blocks concatenated in simulation-id order (the one exchange step)."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return local
    import torch
    world = dist.get_world_size()
    n_opt = local.shape[1]
    max_local = max(shard_range(n_sims, r, world)[1] - shard_range(n_sims, r, world)[0] for r in range(world))
    dev = torch.device("cuda", device) if dist.get_backend() == "nccl" else torch.device("cpu")
    buf = torch.zeros((max_local, n_opt), dtype=torch.float64, device=dev)
    buf[:local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_sims, r, world)
        parts.append(out[r][:hi - lo].cpu().numpy())
    return np.concatenate(parts, axis=0)


def _plan_signature(plan, obs_simulator, n_sims, lo, hi, sim_id0) -> str:
    """What a checkpoint must agree on to be resumed: the stage plan, the truth, the seed and the id range."""
    import hashlib
    h = hashlib.sha256()
    h.update(bytes(plan))
    h.update(np.ascontiguousarray(np.asarray(obs_simulator.mu, dtype=np.float64)).tobytes())
    h.update(np.ascontiguousarray(np.asarray(obs_simulator.cov, dtype=np.float64)).tobytes())
    h.update(repr((int(obs_simulator.n_observations), int(getattr(obs_simulator, "seed", 0)), int(n_sims), int(lo),
                   int(hi), int(sim_id0))).encode())
    return h.hexdigest()


def simulate_errors(obs_simulator, n_sims, optimizers, error_estimator, covariance_transformers, *, device: int = None,
                    wave: int = 0, sim_id0: int = 0, checkpoint: str = None, checkpoint_every: int = 8192):
    """Per-simulation errors [n_sims, n_optimizers] (all ranks' shards gathered) and status words of the local shard.

    `checkpoint` (for the million-simulation configurations): path prefix of a per-rank file
    `<checkpoint>.rank<r>.npz` rewritten after every `checkpoint_every` simulations of the local shard.  A run started
    with the same plan, truth, seed and id range picks up after the last completed chunk; the simulations are keyed by
    id (Philox counter), so a resumed run returns exactly what an uninterrupted one would."""
    dist = _dist()
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
    if device is None:
        device = getattr(obs_simulator, "device", 0)
    plan = build_plan(obs_simulator, optimizers, error_estimator, covariance_transformers, wave)
    n = np.asarray(obs_simulator.cov).shape[0]
    eng = get_engine(n, obs_simulator.n_observations, device, getattr(obs_simulator, "seed", 0))
    eng.ensure_truth(obs_simulator.mu, obs_simulator.cov)
    lo, hi = shard_range(n_sims, rank, world)
    if checkpoint is None:
        err, status = eng.run(plan, hi - lo, sim_id0 + lo)
    else:
        import os
        path = f"{checkpoint}.rank{rank}.npz"
        sig = _plan_signature(plan, obs_simulator, n_sims, lo, hi, sim_id0)
        err = np.full((hi - lo, len(optimizers)), np.nan)
        status = np.zeros(hi - lo, dtype=np.int32)
        done = 0
        if os.path.exists(path):
            with np.load(path, allow_pickle=False) as z:
                if str(z["signature"]) == sig:
                    done = int(z["done"])
                    err[:done], status[:done] = z["errors"][:done], z["status"][:done]
        step = max(1, int(checkpoint_every))
        while done < hi - lo:
            m = min(step, hi - lo - done)
            err[done:done + m], status[done:done + m] = eng.run(plan, m, sim_id0 + lo + done)
            done += m
            tmp = path + ".tmp.npz"
            np.savez(tmp, signature=np.array(sig), done=np.array(done), errors=err[:done], status=status[:done])
            os.replace(tmp, path)
    bad = status & (_lib.ST_HRP_SUM | _lib.ST_NO_EXCESS)
    if bad.any():  # the reference aborts the whole run on these (optimizer.py:220-221)
        if (bad & _lib.ST_HRP_SUM).any():
            raise ValueError("Portfolio allocations don't sum to 1.")
        raise ValueError("no asset has an expected return above the risk-free rate")
    return gather_errors(err, n_sims, device), status


def summarize_errors(errors: np.ndarray, optimizers: List[AbstractOptimizer], confidence: float = 0.95) -> pd.DataFrame:
    """Per-optimizer summary of a per-simulation error dump: `mean` and `stdev` are the reference's
    columns (np.mean / np.std with ddof = 0 over ALL simulations, so a NaN propagates exactly as it does in
    mcos/mcos.py:36-40); next to them the NaN accounting (the Sharpe estimator yields NaN when the allocation equals the
    ideal one, error_estimator.py:41) and NaN-excluded statistics with a normal-approximation confidence interval of
    the mean."""
    from statistics import NormalDist
    z = NormalDist().inv_cdf(0.5 + confidence / 2.0)
    rows = []
    for i, opt in enumerate(optimizers):
        e = np.asarray(errors[:, i], dtype=np.float64)
        ok = e[~np.isnan(e)]
        m = float(ok.mean()) if ok.size else np.nan
        sd = float(ok.std(ddof=1)) if ok.size > 1 else np.nan
        se = sd / np.sqrt(ok.size) if ok.size > 1 else np.nan
        rows.append({'optimizer': opt.name, 'mean': np.mean(e), 'stdev': np.std(e), 'n_sims': int(e.size),
                     'n_nan': int(e.size - ok.size), 'mean_valid': m, 'stderr': se, 'ci_low': m - z * se,
                     'ci_high': m + z * se})
    return pd.DataFrame(rows).set_index('optimizer')


def simulate_optimizations(
        obs_simulator: AbstractObservationSimulator,
        n_sims: int,
        optimizers: List[AbstractOptimizer],
        error_estimator: AbstractErrorEstimator,
        covariance_transformers: List[AbstractCovarianceTransformer],
        *, device: int = None, wave: int = 0
) -> pd.DataFrame:
    """mcos/mcos.py:13-41: DataFrame indexed by optimizer name with the mean and population stdev of the error."""
    errors, _ = simulate_errors(obs_simulator, n_sims, optimizers, error_estimator, covariance_transformers,
                                device=device, wave=wave)
    return pd.DataFrame([
        {'optimizer': opt.name, 'mean': np.mean(errors[:, i]), 'stdev': np.std(errors[:, i])}
        for i, opt in enumerate(optimizers)
    ]).set_index('optimizer')


def simulate_optimizations_from_price_history(
        price_history: pd.DataFrame,
        simulator_name: str,
        n_observations: int,
        n_sims: int,
        optimizers: List[AbstractOptimizer],
        error_estimator: AbstractErrorEstimator,
        covariance_transformers: List[AbstractCovarianceTransformer],
        **kwargs):
    """mcos/mcos.py:44-64."""
    mu, cov = convert_price_history(price_history)
    name = simulator_name.lower()
    if name == "mucovledoitwolf":
        sim = MuCovLedoitWolfObservationSimulator(mu, cov, n_observations)
    elif name == "mucov":
        sim = MuCovObservationSimulator(mu, cov, n_observations)
    elif name == "jackknife":
        sim = MuCovJackknifeObservationSimulator(mu, cov, n_observations)
    else:
        raise ValueError("Invalid observation simulator name")
    return simulate_optimizations(sim, n_sims, optimizers, error_estimator, covariance_transformers, **kwargs)
