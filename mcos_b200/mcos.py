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
_TR_KIND = {DeNoiserCovarianceTransformer: _lib.TR_DENOISE, DetoneCovarianceTransformer: _lib.TR_DETONE}
_MAX_FUSED_OPTIMIZERS = 4


def _gpu_kind(obj, table, method):
    """GPU stage id of a plugin object: one of the built-in classes, or a subclass of one that does not override the
    behaviour method (`simulate` / `transform` / `allocate` / `estimate`).  None for everything else -- user plugins run
    on the host, one simulation at a time, exactly as the reference calls them (mcos/mcos.py:22-33)."""
    for cls, kind in table.items():
        if isinstance(obj, cls) and getattr(type(obj), method) is getattr(cls, method):
            return kind
    return None


def shard_range(n_sims: int, rank: int, world_size: int):
    """Contiguous simulation-id range [lo, hi) of `rank`: the first n_sims % world_size ranks get one extra."""
    base, extra = divmod(n_sims, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _as_list(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


def build_plan(obs_simulator, optimizers, error_estimator, covariance_transformers, wave: int = 0) -> _lib.Plan:
    """Compiles the plugin objects into the stage plan mcosb_run executes: any list of DeNoiser / Detone transformers in
    order, up to four of the built-in optimizers (RiskParity with its `target_risk`), and one estimator or a list of
    them (all evaluated in ONE pass; the error array then gets a trailing estimator axis).  Plugin types without a GPU
    stage raise TypeError here; `simulate_errors` runs those on the host around the GPU stages instead."""
    sim_kind = _gpu_kind(obs_simulator, _SIM_KIND, "simulate")
    if sim_kind is None:
        raise TypeError(f"no GPU stage for observation simulator {type(obs_simulator).__name__}")
    estimators = _as_list(error_estimator)
    if not 1 <= len(estimators) <= 3:
        raise ValueError("between 1 and 3 error estimators per pass")
    err_kinds = [_gpu_kind(e, _ERR_KIND, "estimate") for e in estimators]
    if None in err_kinds:
        raise TypeError(f"no GPU stage for error estimator {type(estimators[err_kinds.index(None)]).__name__}")
    plan = _lib.Plan()
    plan.moments_kind = sim_kind
    plan.denoise, plan.bandwidth, plan.detone_n_remove = 0, 0.25, 0
    trs = []
    for tr in covariance_transformers:
        kind = _gpu_kind(tr, _TR_KIND, "transform")
        if kind is None:
            raise TypeError(f"no GPU stage for covariance transformer {type(tr).__name__}")
        if kind == _lib.TR_DETONE and int(tr.n_remove) == 0:
            continue  # the identity (covariance_transformer.py:223-224)
        trs.append((kind, float(tr.bandwidth) if kind == _lib.TR_DENOISE else float(int(tr.n_remove))))
    if len(trs) > _lib.MAX_TRANSFORMERS:
        raise ValueError(f"at most {_lib.MAX_TRANSFORMERS} covariance transformers per plan")
    plan.n_transformers = len(trs)
    for i, (kind, arg) in enumerate(trs):
        plan.transformer_kind[i], plan.transformer_arg[i] = kind, arg
    if not 1 <= len(optimizers) <= _MAX_FUSED_OPTIMIZERS:
        raise ValueError(f"between 1 and {_MAX_FUSED_OPTIMIZERS} optimizers per plan")
    plan.n_optimizers = len(optimizers)
    plan.nco_max_clusters, plan.nco_trials = 0, 10
    plan.risk_budget = None
    for i, opt in enumerate(optimizers):
        kind = _gpu_kind(opt, _OPT_KIND, "allocate")
        if kind is None:
            raise TypeError(f"no GPU stage for optimizer {type(opt).__name__}")
        plan.optimizers[i] = kind
        if kind == _lib.OPT_RISK_PARITY and opt.target_risk is not None:
            budget = np.ascontiguousarray(np.asarray(opt.target_risk, dtype=np.float64).reshape(-1))
            if getattr(plan, "_budget", None) is not None and not np.array_equal(plan._budget, budget):
                raise ValueError("one risk budget per plan")
            plan._budget = budget                       # keeps the buffer alive as long as the plan
            plan.risk_budget = budget.ctypes.data
        if kind == _lib.OPT_NCO:
            plan.nco_max_clusters = 0 if opt.max_num_clusters is None else int(opt.max_num_clusters)
            plan.nco_trials = int(opt.num_clustering_trials)
    plan.error_kind = err_kinds[0]
    if isinstance(error_estimator, (list, tuple)):
        plan.n_error_kinds = len(err_kinds)
        for k, kind in enumerate(err_kinds):
            plan.error_kinds[k] = kind
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
    """All ranks' [n_local, ...] blocks concatenated in simulation-id order (the one exchange step)."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return local
    import torch
    world = dist.get_world_size()
    tail = local.shape[1:]
    max_local = max(shard_range(n_sims, r, world)[1] - shard_range(n_sims, r, world)[0] for r in range(world))
    dev = torch.device("cuda", device) if dist.get_backend() == "nccl" else torch.device("cpu")
    buf = torch.zeros((max_local,) + tuple(tail), dtype=torch.float64, device=dev)
    buf[:local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64)).to(dev)
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
    raw = bytearray(bytes(plan))
    off = type(plan).risk_budget.offset                 # the pointer value itself is not part of the identity
    raw[off:off + 8] = b"\0" * 8
    h.update(bytes(raw))
    if getattr(plan, "_budget", None) is not None:
        h.update(plan._budget.tobytes())
    h.update(np.ascontiguousarray(np.asarray(obs_simulator.mu, dtype=np.float64)).tobytes())
    h.update(np.ascontiguousarray(np.asarray(obs_simulator.cov, dtype=np.float64)).tobytes())
    h.update(repr((int(obs_simulator.n_observations), int(getattr(obs_simulator, "seed", 0)), int(n_sims), int(lo),
                   int(hi), int(sim_id0))).encode())
    return h.hexdigest()


def status_counts(status: np.ndarray) -> dict:
    """{flag name: number of simulations with that MCOSB_ST_* bit set} for the bits that occur."""
    status = np.asarray(status)
    return {name: int(((status & bit) != 0).sum()) for bit, name in _lib.ST_NAMES.items() if ((status & bit) != 0).any()}


def _run_fused(eng, obs_simulator, optimizers, error_estimator, covariance_transformers, wave, n_local, id0):
    """Errors [n_local, n_opt(, n_kinds)] and status of ids id0 .. id0 + n_local - 1 through mcosb_run; more than four
    optimizers take several passes over the same simulations (Philox draws are keyed by id, so they see the same
    observations)."""
    errs, status = [], np.zeros(n_local, dtype=np.int32)
    for g0 in range(0, len(optimizers), _MAX_FUSED_OPTIMIZERS):
        plan = build_plan(obs_simulator, optimizers[g0:g0 + _MAX_FUSED_OPTIMIZERS], error_estimator,
                          covariance_transformers, wave)
        e, st = eng.run(plan, n_local, id0)
        errs.append(e)
        status |= st
    return (errs[0] if len(errs) == 1 else np.concatenate(errs, axis=1)), status


def _run_staged(eng, obs_simulator, optimizers, estimators, covariance_transformers, n_local, id0, multi):
    """The loop body of mcos/mcos.py:22-33 with user plugins in it: every built-in plugin runs batched on the GPU through
    its `*_batch` method, a plugin type the engine does not know is called on the host once per simulation with the
    reference's arguments (`simulate()`, `transform(cov_hat, n_observations)`, `allocate(mu_hat (N, 1), cov_hat)`,
    `estimate(mu, cov, allocation, ideal)`).  The ideal allocation is loop-invariant and computed once."""
    mu, cov = obs_simulator.mu, obs_simulator.cov
    n = np.asarray(cov).shape[0]
    n_obs = obs_simulator.n_observations
    sim_kind = _gpu_kind(obs_simulator, _SIM_KIND, "simulate")
    ideal = [np.asarray(opt.allocate(mu, cov), dtype=np.float64).reshape(-1) for opt in optimizers]   # mcos.py:30
    errors = np.empty((n_local, len(optimizers), len(estimators)))
    status = np.zeros(n_local, dtype=np.int32)
    chunk = max(1, min(n_local, (1 << 28) // (n * n * 8)))      # <= 256 MiB of covariances on the host at a time
    for c0 in range(0, n_local, chunk):
        m = min(chunk, n_local - c0)
        if sim_kind is not None:
            mu_hat, cov_hat = obs_simulator.simulate_batch(m, id0 + c0)
        else:
            draws = [obs_simulator.simulate() for _ in range(m)]
            mu_hat = np.stack([np.asarray(d[0], dtype=np.float64).reshape(-1) for d in draws])
            cov_hat = np.stack([np.asarray(d[1], dtype=np.float64) for d in draws])
        for tr in covariance_transformers:
            kind = _gpu_kind(tr, _TR_KIND, "transform")
            if kind == _lib.TR_DENOISE:
                cov_hat, info = tr.transform_batch(cov_hat, n_obs, return_details=True)
                status[c0:c0 + m] |= info["status"]
            elif kind == _lib.TR_DETONE:
                cov_hat = tr.transform_batch(cov_hat)
            else:
                cov_hat = np.stack([np.asarray(tr.transform(c, n_obs), dtype=np.float64) for c in cov_hat])
        for o, opt in enumerate(optimizers):
            kind = _gpu_kind(opt, _OPT_KIND, "allocate")
            if kind == _lib.OPT_MARKOWITZ or kind == _lib.OPT_NCO:
                w, info = opt.allocate_batch(mu_hat, cov_hat, return_details=True)
            elif kind == _lib.OPT_HRP or kind == _lib.OPT_RISK_PARITY:
                w, info = opt.allocate_batch(cov_hat, return_details=True)
            else:
                w = np.stack([np.asarray(opt.allocate(mu_hat[i].reshape(-1, 1), cov_hat[i]), dtype=np.float64).reshape(-1)
                              for i in range(m)])
                info = None
            if info is not None:
                status[c0:c0 + m] |= info["status"]
            gpu_kinds = [_gpu_kind(e, _ERR_KIND, "estimate") for e in estimators]
            if None not in gpu_kinds:
                eng.set_truth(mu, cov, require_pd=False)
                errors[c0:c0 + m, o, :] = eng.errors_multi(w, ideal[o], gpu_kinds)
            else:
                for k, est in enumerate(estimators):
                    if gpu_kinds[k] is not None:
                        errors[c0:c0 + m, o, k] = est.estimate_batch(mu, cov, w, ideal[o])
                    else:
                        errors[c0:c0 + m, o, k] = [np.ravel(est.estimate(mu, cov, w[i], ideal[o]))[0] for i in range(m)]
    return (errors if multi else errors[:, :, 0]), status


def simulate_errors(obs_simulator, n_sims, optimizers, error_estimator, covariance_transformers, *, device: int = None,
                    wave: int = 0, sim_id0: int = 0, checkpoint: str = None, checkpoint_every: int = 8192,
                    warn: bool = True):
    """Per-simulation errors [n_sims, n_optimizers] and status words [n_sims], both for ALL simulations (every rank's
    shard gathered in simulation-id order).  `error_estimator` may be a list of up to three estimators: they are then
    evaluated in one pass over the simulations and the result is [n_sims, n_optimizers, n_estimators]
    (error_estimator.py:21-49 are an O(N^2) epilogue on the same weights).

    Built-in plugin types run fused on the GPU (mcosb_run).  If any plugin is a user-defined type (or overrides the
    behaviour method of a built-in), the loop runs stage by stage with that plugin called on the host per simulation.

    Per-simulation numerical conditions do not abort the run (MCOSB_ST_* bits in `status`; a `RuntimeWarning` with the
    counts is issued when `warn`), with one exception the reference shares: HRP weights that do not sum to 1 raise
    ValueError (optimizer.py:220-221) -- consistently on every rank.

    `checkpoint` (for the million-simulation configurations): path prefix of a per-rank file
    `<checkpoint>.rank<r>.npz` rewritten after every `checkpoint_every` simulations of the local shard.  A run started
    with the same plan, truth, seed and id range picks up after the last completed chunk; the simulations are keyed by
    id (Philox counter), so a resumed run returns exactly what an uninterrupted one would."""
    dist = _dist()
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
    if device is None:
        device = getattr(obs_simulator, "device", 0)
    multi = isinstance(error_estimator, (list, tuple))
    estimators = _as_list(error_estimator)
    optimizers = list(optimizers)
    fused = (_gpu_kind(obs_simulator, _SIM_KIND, "simulate") is not None
             and all(_gpu_kind(e, _ERR_KIND, "estimate") is not None for e in estimators) and 1 <= len(estimators) <= 3
             and all(_gpu_kind(t, _TR_KIND, "transform") is not None for t in covariance_transformers)
             and len(covariance_transformers) <= _lib.MAX_TRANSFORMERS
             and all(_gpu_kind(o, _OPT_KIND, "allocate") is not None for o in optimizers) and len(optimizers) >= 1)
    n = np.asarray(obs_simulator.cov).shape[0]
    eng = get_engine(n, getattr(obs_simulator, "n_observations", 1), device, getattr(obs_simulator, "seed", 0))
    lo, hi = shard_range(n_sims, rank, world)

    def run(m, id0):
        if fused:
            return _run_fused(eng, obs_simulator, optimizers, error_estimator, covariance_transformers, wave, m, id0)
        return _run_staged(eng, obs_simulator, optimizers, estimators, covariance_transformers, m, id0, multi)

    if fused:
        eng.ensure_truth(obs_simulator.mu, obs_simulator.cov)
    if checkpoint is None:
        err, status = run(hi - lo, sim_id0 + lo)
    else:
        if not fused:
            raise TypeError("checkpointing needs simulations keyed by id: built-in (GPU) plugins only")
        import os
        path = f"{checkpoint}.rank{rank}.npz"
        plan0 = build_plan(obs_simulator, optimizers[:_MAX_FUSED_OPTIMIZERS], error_estimator, covariance_transformers, wave)
        sig = _plan_signature(plan0, obs_simulator, n_sims, lo, hi, sim_id0) + f"|{len(optimizers)}"
        shape = (hi - lo, len(optimizers)) + ((len(estimators),) if multi else ())
        err = np.full(shape, np.nan)
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
            err[done:done + m], status[done:done + m] = run(m, sim_id0 + lo + done)
            done += m
            tmp = path + ".tmp.npz"
            np.savez(tmp, signature=np.array(sig), done=np.array(done), errors=err[:done], status=status[:done])
            os.replace(tmp, path)
    # one exchange: errors and status words together, so every rank sees every simulation's status and takes the same
    # decision below (a rank raising alone would leave the others waiting in the collective)
    flat = np.concatenate([err.reshape(err.shape[0], -1), status.astype(np.float64)[:, None]], axis=1)
    flat = gather_errors(flat, n_sims, device)
    status = flat[:, -1].astype(np.int32)
    err = flat[:, :-1].reshape((flat.shape[0],) + err.shape[1:])
    if (status & _lib.ST_HRP_SUM).any():  # the reference aborts the whole run on this (optimizer.py:220-221)
        raise ValueError("Portfolio allocations don't sum to 1.")
    if warn and status.any():
        import warnings
        warnings.warn(f"{int((status != 0).sum())} of {n_sims} simulations carry status flags {status_counts(status)} "
                      "(include/mcosb.h MCOSB_ST_*); their errors are included in the statistics", RuntimeWarning,
                      stacklevel=2)
    return err, status


def summarize_errors(errors: np.ndarray, optimizers: List[AbstractOptimizer], confidence: float = 0.95,
                     status: np.ndarray = None) -> pd.DataFrame:
    """Per-optimizer summary of a per-simulation error dump: `mean` and `stdev` are the reference's
    columns (np.mean / np.std with ddof = 0 over ALL simulations, so a NaN propagates exactly as it does in
    mcos/mcos.py:36-40); next to them the NaN accounting (the Sharpe estimator yields NaN when the allocation equals the
    ideal one, error_estimator.py:41) and NaN-excluded statistics with a normal-approximation confidence interval of
    the mean.  With `status` (the per-simulation words simulate_errors returns) the table carries `n_flagged`, the number of
    simulations with any MCOSB_ST_* bit set."""
    from statistics import NormalDist
    n_flagged = None if status is None else int((np.asarray(status) != 0).sum())
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
        if n_flagged is not None:
            rows[-1]['n_flagged'] = n_flagged
    return pd.DataFrame(rows).set_index('optimizer')


def simulate_optimizations(
        obs_simulator: AbstractObservationSimulator,
        n_sims: int,
        optimizers: List[AbstractOptimizer],
        error_estimator: AbstractErrorEstimator,
        covariance_transformers: List[AbstractCovarianceTransformer],
        *, device: int = None, wave: int = 0
) -> pd.DataFrame:
    """mcos/mcos.py:13-41: DataFrame indexed by optimizer name with the mean and population stdev of the error.
    Extension: a LIST of error estimators returns {estimator class name: DataFrame}, all from one pass."""
    errors, _ = simulate_errors(obs_simulator, n_sims, optimizers, error_estimator, covariance_transformers,
                                device=device, wave=wave)

    def table(e):
        return pd.DataFrame([
            {'optimizer': opt.name, 'mean': np.mean(e[:, i]), 'stdev': np.std(e[:, i])}
            for i, opt in enumerate(optimizers)
        ]).set_index('optimizer')

    if isinstance(error_estimator, (list, tuple)):
        return {type(est).__name__: table(errors[:, :, k]) for k, est in enumerate(error_estimator)}
    return table(errors)


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
