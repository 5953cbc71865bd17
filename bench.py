#!/usr/bin/env python
"""Benchmark of the MCOS Monte-Carlo loop on B200 (BASELINE.json metric: Monte-Carlo simulations per second at
N=500, T=1000, Ledoit-Wolf + DeNoiser, Markowitz + HRP, Sharpe error).

    python bench.py --gpus 1 --steps K --warmup W                 # this engine
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                     # one rank per GPU, simulations sharded by id
    python bench.py --impl reference --gpus 1 --steps K --warmup W # the reference's CPU path (oracle port) on host cores

A "step" is one pass of the whole loop body (sample -> moments -> de-noise -> Markowitz + HRP -> errors) over a batch of
`--sims-per-step` simulations per GPU; the 100 000-simulation job of BASELINE config 3 is ceil(1e5 / batch) such steps
and its throughput in simulations/s is the same number.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ASSETS, N_OBS = 500, 1000
WORKLOAD = "BASELINE config 3: synthetic block-diagonal truth N=500 (10 blocks x 50, rho=0.5), T=1000, " \
           "MuCovLedoitWolf + DeNoiser(0.25), Markowitz + HRP, SharpeRatioErrorEstimator"


def load_peaks():
    """Roofline denominators: HBM from the driver-written MEASURED_PEAKS.json (fallback 6650 GB/s), FP64 from this
    repo's own measurement on the same pool (profiles/fp64_peaks_r01.json; MEASURED_PEAKS.json has no FP64 entry)."""
    hbm, src = 6650.0, "fallback"
    try:
        hbm, src = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        pass
    fp64 = {"dfma_tflops": 36.7, "dmma_tflops": 37.0, "dmul_dadd_tinstr": 18.28}
    try:
        raw = json.load(open(os.path.join(ROOT, "profiles", "fp64_peaks_r01.json")))
        fp64 = {"dfma_tflops": raw["dfma_tflops_b2_t1024"], "dmma_tflops": raw["dmma1688_tflops_b2_t512"],
                "dmul_dadd_tinstr": raw["dmul_dadd_tinstr"]}
    except Exception:
        pass
    return hbm, src, fp64


def trmm_flops(cov, t):
    """Flop per simulation of X = Z L' as the TRMM executes it: per set of 32 columns of X (= 32 rows of L, one warp
    column) it visits the 16-wide k-slabs of L that hold a non-zero there (mcosb_set_truth lists them; a dense factor
    gives the triangle T N^2, a block-diagonal truth much less)."""
    low = np.linalg.cholesky(np.asarray(cov, dtype=np.float64))
    n = low.shape[0]
    slabs = 0
    for r0 in range(0, n, 32):
        rows = low[r0:r0 + 32]
        slabs += sum(bool(np.any(rows[:, k0:k0 + 16] != 0.0)) for k0 in range(0, n, 16))
    return 2.0 * t * 32 * 16 * slabs


def kernel_models(n, t, cov=None):
    """Algorithmic work per simulation of each kernel: (bound, work per simulation, unit)."""
    n3 = float(n) ** 3
    return {
        "philox_normal": ("hbm", 8.0 * t * n, "B"),                       # writes Z once
        # triangular Z L'; fewer when whole slabs of L are zero and skipped (not the case for the permuted block-diagonal truth)
        "trmm": ("tensor", min(trmm_flops(cov, t), 2.0 * t * n * n / 2) if cov is not None else 2.0 * t * n * n / 2, "flop"),
        "colmean": ("hbm", 8.0 * t * n, "B"),
        "syrk": ("tensor", 2.0 * t * n * n / 2, "flop"),                  # lower triangle of Xc'Xc
        "lw_rownorm": ("hbm", 8.0 * t * n, "B"),
        "lw_finish": ("hbm", 16.0 * n * n, "B"),
        "corr": ("hbm", 16.0 * n * n, "B"),
        "tridiag": ("fp64", 4.0 * n3 / 3, "flop"),                        # Householder tridiagonalisation
        "tridiag2": ("fp64", 4.0 * n3 / 3, "flop"),                       # blocked, lower triangle only
        "sy2sb": ("tensor", 4.0 * n3 / 3, "flop"),                        # dense -> band (b = 8): rank-16 update + A V on DMMA
        # batched dense -> band (sy2sb3.cu): the pass reads and writes the trailing triangle once per panel of 8 columns,
        # 2 x 8 B x sum_p m_p^2 / 2 = N^3 / 3 bytes (41.7 MB at N = 500; 6.4 us at the measured HBM peak against 4.5 us for its
        # 4 N^3 / 3 flop on the DMMA pipe, so HBM is the binding roofline); the panel kernel moves ~6 x 64 B per row and panel
        "sy2sb_pass": ("hbm", n3 / 3, "B"),
        "sy2sb_panel": ("hbm", 6 * 64.0 * n * n / 16, "B"),
        "sb2st": ("fp64", 6.0 * 8 * n * n, "flop"),                       # bulge chasing, 6 N^2 b flop
        "q2back": ("hbm", 8.0 * n * n, "B"),                              # stage-2 reflectors read once per simulation
        # one multisection count (4 flop per recurrence step) + ~13 count-and-Newton evaluations (8 flop per step; the
        # slowest lane of a warp, scratch/eig_newton_proto.py) per eigenvalue, N steps each
        "eigvals": ("fp64", (4.0 + 13 * 8.0) * n * n, "flop"),
        # Taylor table of the kernel sum: 48 nodes x N x (exp ~25 + 13 x 4 flop) once, then ~11 objective passes x 1000 points
        # x (2 Horner recurrences of degree 12 + the MP density ~40 flop)
        "mpfit": ("fp64", 48.0 * n * 77 + 11 * 1000.0 * 92, "flop"),
        "eigvec": ("hbm", 3 * 8.0 * n * n, "B"),                          # reflectors read + cov written and rescaled
        "inviter": ("hbm", 8 * 5 * 8.0 * n * 10, "B"),                    # ~8 passes over 5 N-vectors for ~10 eigenvectors
        "backtransform": ("hbm", 8.0 * n * n / 2, "B"),                   # reflectors read once (shared by the vectors via L1/L2)
        "reconstruct": ("hbm", 8.0 * n * n, "B"),                         # de-noised covariance written once
        "markowitz": ("hbm", 8.0 * n * 50 * 60, "B"),                     # ~60 iterations x |F|~50 rows of C from L2
        "hrp_dist": ("hbm", 16.0 * n * n, "B"),
        "hrp_gram": ("tensor", n3, "flop"),                               # lower triangle of D D' on DMMA (fast path)
        "hrp_euclid": ("fp64_nofma", 3.0 * n3 / 2, "instr"),              # sub, mul, add per (i<j, k): exact fallback only
        "hrp_prim": ("hbm", 8.0 * n * n, "B"),                            # Prim reads every row of the squared distances once
        "hrp_tree": ("hbm", 2 * 8.0 * n * n, "B"),                        # merge heights: two rows of D per edge
        "hrp_permute": ("hbm", 16.0 * n * n, "B"),                        # covariance gathered into quasi-diagonal order
        "hrp_bisect": ("hbm", 16.0 * n * n, "B"),                         # every level reads its diagonal blocks: 2 N^2 total
        "errors": ("hbm", 8.0 * n * n / 8, "B"),                          # true cov shared by 8 simulations per CTA
    }


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
        "clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        busy = sorted(x for x in sm if mx is None or x > 0.3 * mx) or sorted(sm)
        med = busy[len(busy) // 2] if busy else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference loop on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def _cpu_worker(seed):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import oracle
    from mcos_b200.synthetic import block_diagonal_truth
    mu, cov = block_diagonal_truth(N_ASSETS)
    np.random.seed(seed)
    t0 = time.time()
    oracle.simulate_optimizations(mu, cov, N_OBS, 1, simulator="mucovledoitwolf", transformers=[("denoise", 0.25)],
                                  optimizers=("markowitz", "HRP"), estimator="sharpe")
    return time.time() - t0


def cpu_reference_pass(cores, seed0):
    """One bounded sample: `cores` simulations of the headline workload, one per worker process with BLAS threads
    pinned to 1 (the reference itself is single-threaded Python). Returns (simulations, wall seconds)."""
    import multiprocessing as mp
    saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
    for k in saved:
        os.environ[k] = "1"
    try:
        ctx = mp.get_context("spawn")
        t0 = time.time()
        with ctx.Pool(cores) as pool:
            pool.map(_cpu_worker, [seed0 + i for i in range(cores)])
        return cores, time.time() - t0
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(os.cpu_count() or 1, args.cpu_cores or (os.cpu_count() or 1))
    budget = float(os.environ.get("MCOSB_REF_BUDGET_S", "420"))
    t_start = time.time()
    # every pass costs about one reference simulation per core (~1 min at N=500: two SLSQP solves of ~25 s each), so
    # the W + K passes are cut short when the next one would overrun the time budget; the LAST min(K, done) passes are
    # the timed ones, whatever ran before them is warm-up
    passes = []
    for i in range(args.warmup + args.steps):
        elapsed = time.time() - t_start
        if passes and elapsed + elapsed / len(passes) > budget:
            break
        passes.append(cpu_reference_pass(cores, 1000 * i))
    timed = passes[-min(args.steps, len(passes)):]
    done_warm = len(passes) - len(timed)
    times = [dt for _, dt in timed]
    sims = sum(n for n, _ in timed)
    total = sum(times)
    value = sims / total
    sample = f"{cores} simulations per step (one per worker process, BLAS threads = 1), {len(times)} timed step(s), " \
             f"{done_warm} warm-up step(s); time budget {budget:.0f} s"
    line = {
        "impl": "reference", "metric": "Monte Carlo sims/sec (N=500,T=1000, DeNoise+Markowitz+HRP)", "value": value,
        "unit": "sims/s", "n_gpus": args.gpus, "steps": len(times), "steps_requested": args.steps,
        "warmup": done_warm, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_assets": N_ASSETS, "n_observations": N_OBS, "sims_per_step": cores,
                   "note": "oracle port of mcos/mcos.py:22-33 (numpy/scipy/sklearn + PyPortfolioOpt-0.5.2-semantics "
                           "SLSQP), ideal allocation recomputed per simulation as the reference does"},
        "cpu_baseline": {"value": value, "unit": "sims/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



# ---------------------------------------------------------------------------------------------------------------------
# BASELINE configs 1, 2, 4, 5 (configs[0], [1], [3], [4]); config 3 (configs[2]) is the headline above
# ---------------------------------------------------------------------------------------------------------------------
def _oracle_check(mb, name, mu, cov, t, sim, opts, ests, trs, x, gpu_errs):
    """cpu_baseline leg of one configuration: the oracle port runs the same loop body (mcos/mcos.py:22-33) on the
    observations the GPU drew for its first simulations -- timed, one process -- and its per-simulation errors are the
    parity check of the GPU's.  Two liberties, both in the CPU's favour and both stated in the record: Markowitz goes
    through the exact QP (what the GPU solves; the reference's SLSQP takes minutes per call at N = 1000) and the ideal
    allocation is computed once instead of once per simulation."""
    import oracle
    kinds = {"ExpectedOutcomeErrorEstimator": "expected_outcome", "VarianceErrorEstimator": "variance",
             "SharpeRatioErrorEstimator": "sharpe"}
    moments = {"MuCovObservationSimulator": oracle.moments_mucov,
               "MuCovLedoitWolfObservationSimulator": oracle.moments_ledoit_wolf,
               "MuCovJackknifeObservationSimulator": oracle.moments_jackknife}[type(sim).__name__]

    def allocate(opt, m, c):
        if opt.name == "markowitz":
            return oracle.markowitz_exact(m, c)
        if opt.name == "HRP":
            return oracle.hrp(c)
        if opt.name == "NCO":
            return oracle.nco(m, c, opt.max_num_clusters, 1)      # every clustering trial repeats the same fits
        return oracle.risk_parity(c, opt.target_risk)

    tol = {"markowitz": 5e-3, "HRP": 1e-7, "NCO": 1e-5, "Risk Parity": 1e-6}   # Markowitz: one 1e-5 rounding flip
    t0 = time.time()
    ideal = [allocate(o, mu, cov) for o in opts]
    t_ideal = time.time() - t0
    worst, ok = {}, True
    t0 = time.time()
    for i in range(len(x)):
        mu_hat, cov_hat = moments(x[i])
        for tr in trs:
            cov_hat = oracle.denoise(cov_hat, t, tr.bandwidth) if hasattr(tr, "bandwidth") else oracle.detone(cov_hat, tr.n_remove)
        for o, opt in enumerate(opts):
            w = allocate(opt, mu_hat, cov_hat)
            for k, est in enumerate(ests):
                kind = kinds[type(est).__name__]
                want = float(np.ravel(oracle.estimate_error(kind, mu, cov, w, ideal[o]))[0])
                got = float(gpu_errs[i, o, k])
                rel = 0.0 if (np.isnan(want) and np.isnan(got)) else abs(got - want) / max(abs(want), 1e-300)
                if rel != rel:
                    rel = float("inf")       # NaN on one side only
                key = f"{opt.name}/{kind}"
                worst[key] = max(worst.get(key, 0.0), rel)
                ok &= rel <= tol[opt.name]
    dt = time.time() - t0
    return {"parity": "pass" if ok else "fail", "max_rel_diff": worst, "checked_sims": len(x),
            "cpu_sims_per_s": len(x) / dt,
            "cpu_note": "oracle port, one process, exact-QP Markowitz, ideal allocations hoisted (%.1f s, not counted)" % t_ideal}


def run_other_configs(args, mb, local, rank, world, barrier, drop_engines):
    import torch
    import torch.distributed as dist
    from mcos_b200.engine import get_engine
    from mcos_b200.synthetic import block_diagonal_truth
    stock = np.load(os.path.join(ROOT, "tests", "golden", "stock.npz"))
    mu100, cov100 = block_diagonal_truth(100)
    mu1k, cov1k = block_diagonal_truth(1000)
    three = [mb.ExpectedOutcomeErrorEstimator(), mb.VarianceErrorEstimator(), mb.SharpeRatioErrorEstimator()]
    cases = {
        "1": dict(desc="README sample: tests/stock_prices.csv (20 stocks), MuCov, T=50, 50 sims, DeNoiser, Markowitz+HRP, Variance",
                  mu=stock["mu"], cov=stock["cov"], t=50, n_sims=50, check=4,
                  sim=mb.MuCovObservationSimulator, opts=[mb.MarkowitzOptimizer(), mb.HRPOptimizer()],
                  ests=mb.VarianceErrorEstimator(), trs=[mb.DeNoiserCovarianceTransformer()]),
        "2": dict(desc="N=100, T=200, 10k sims, MuCov + DeNoiser + Markowitz, Variance", mu=mu100, cov=cov100, t=200,
                  n_sims=10000, check=8, sim=mb.MuCovObservationSimulator, opts=[mb.MarkowitzOptimizer()],
                  ests=mb.VarianceErrorEstimator(), trs=[mb.DeNoiserCovarianceTransformer()]),
        "4": dict(desc="N=100, T=200, 50k sims, Jackknife + NCO(max_num_clusters=10, num_clustering_trials=10), Variance",
                  mu=mu100, cov=cov100, t=200, n_sims=50000, check=4, sim=mb.MuCovJackknifeObservationSimulator,
                  opts=[mb.NCOOptimizer(10, 10)], ests=mb.VarianceErrorEstimator(), trs=[]),
        "5": dict(desc="N=1000, T=2000, Ledoit-Wolf + DeNoiser, Markowitz+HRP+NCO(10)+RiskParity, all three estimators in one "
                       "pass; bounded sample of the 1M-simulation job: %d sims per GPU" % args.config5_sims,
                  mu=mu1k, cov=cov1k, t=2000, n_sims=args.config5_sims * world, check=2,
                  sim=mb.MuCovLedoitWolfObservationSimulator,
                  opts=[mb.MarkowitzOptimizer(), mb.HRPOptimizer(), mb.NCOOptimizer(10, 10), mb.RiskParityOptimizer()],
                  ests=three, trs=[mb.DeNoiserCovarianceTransformer()]),
        "5_without_nco": dict(desc="config 5 without the NCO optimizer (Markowitz+HRP+RiskParity, three estimators in one pass): "
                                   "%d sims per GPU" % args.config5_sims,
                              mu=mu1k, cov=cov1k, t=2000, n_sims=args.config5_sims * world, check=1,
                              sim=mb.MuCovLedoitWolfObservationSimulator,
                              opts=[mb.MarkowitzOptimizer(), mb.HRPOptimizer(), mb.RiskParityOptimizer()],
                              ests=three, trs=[mb.DeNoiserCovarianceTransformer()]),
    }
    out = {}
    for key, c in cases.items():
        import warnings
        sim = c["sim"](c["mu"], c["cov"], c["t"], seed=args.seed, device=local)
        ests = c["ests"]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            mb.simulate_errors(sim, c["n_sims"], c["opts"], ests, c["trs"], device=local)   # warm-up: arena sized, ideal cached
            barrier()
            t0 = time.perf_counter()
            errs, status = mb.simulate_errors(sim, c["n_sims"], c["opts"], ests, c["trs"], device=local, sim_id0=1000)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt.item())
        rec = {"workload": c["desc"], "n_sims": c["n_sims"], "wall_s": dt, "sims_per_s": c["n_sims"] / dt,
               "status_nonzero": int((status != 0).sum()), "nan_errors": int(np.isnan(errs).sum()),
               "api": "mcos_b200.simulate_errors (host numpy in/out)"}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            n = np.asarray(c["cov"]).shape[0]
            eng = get_engine(n, c["t"], local, args.seed)
            x = eng.sample(c["check"], 1000)
            e3 = errs if errs.ndim == 3 else errs[:, :, None]
            try:
                rec.update(_oracle_check(mb, key, np.asarray(c["mu"]), np.asarray(c["cov"]), c["t"], sim, c["opts"],
                                         ests if isinstance(ests, list) else [ests], c["trs"], x, e3))
            except Exception as exc:  # pragma: no cover
                rec.update({"parity": f"check failed to run: {exc}"})
        out[key] = rec
        drop_engines()
    return out


# ---------------------------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist

    from mcos_b200 import _lib
    import mcos_b200 as mb
    from mcos_b200.engine import get_engine
    from mcos_b200.synthetic import block_diagonal_truth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")        # keep stdout to the one JSON line (NCCL prints its version there at these levels)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.sims_per_step <= 0:
        args.sims_per_step = 48 * torch.cuda.get_device_properties(local).multi_processor_count
    spp = args.sims_per_step
    mu, cov = block_diagonal_truth(N_ASSETS)
    sim = mb.MuCovLedoitWolfObservationSimulator(mu, cov, N_OBS, seed=args.seed, device=local)
    optimizers = [mb.MarkowitzOptimizer(), mb.HRPOptimizer()]
    estimator = mb.SharpeRatioErrorEstimator()
    transformers = [mb.DeNoiserCovarianceTransformer()]
    plan = mb.build_plan(sim, optimizers, estimator, transformers)
    eng = get_engine(N_ASSETS, N_OBS, local, args.seed)
    eng.ensure_truth(mu, cov)          # inputs resident in HBM before the timed region
    # one explicit stream for everything: the kernels are launched on it and the CUDA events that time them are
    # recorded on it (torch.cuda.Event sees only torch's current stream)
    bench_stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(bench_stream)
    eng.use_torch_stream()

    def step(i, out_torch=True):
        # rank r, step i works on its own contiguous block of simulation ids (weak scaling, no data-path collective)
        sim0 = (i * world + rank) * spp
        return eng.run(plan, spp, sim0, out_torch=out_torch)

    for i in range(args.warmup):
        step(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count
    eng.hrp_fallback_count()           # reset: count the exact-fallback simulations of the timed steps only
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    status_bad = 0
    for i in range(args.steps):
        err, st = step(args.warmup + i)
    e1.record()
    barrier()
    ms_local = e0.elapsed_time(e1)
    launches = eng.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    hrp_fallback = eng.hrp_fallback_count()
    stage_ms = eng.last_stage_ms()
    status_bad = int((st != 0).sum().item())
    # per-kernel device times (one CUDA event per launch): a separate, untimed pass over one more step -- the product
    # path records no events, so neither does the timed region
    eng.set_profiling(True)
    step(args.warmup + args.steps)
    ktimes = eng.kernel_times()
    eng.set_profiling(False)
    nan_frac = float(torch.isnan(err).float().mean().item())

    t = torch.tensor([ms_local], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * spp * args.steps / (ms_total * 1e-3)

    # ---- end to end through the public API with host buffers (truth re-uploaded from pinned memory every step) --------
    e2e = None
    if not args.no_e2e:
        pin_mu = torch.empty(N_ASSETS, dtype=torch.float64).pin_memory()
        pin_cov = torch.empty((N_ASSETS, N_ASSETS), dtype=torch.float64).pin_memory()
        pin_mu.copy_(torch.from_numpy(mu)); pin_cov.copy_(torch.from_numpy(cov))
        sim_h = mb.MuCovLedoitWolfObservationSimulator(pin_mu.numpy(), pin_cov.numpy(), N_OBS, seed=args.seed, device=local)

        def e2e_step(i):
            eng.mu = None        # forget the resident truth: this step uploads (mu, cov) again and refactors it
            errors, _ = mb.simulate_errors(sim_h, spp * world, optimizers, estimator, transformers, device=local,
                                           sim_id0=i * world * spp)
            return errors        # host numpy [world * spp, 2], gathered over ranks in simulation-id order
        e2e_step(0)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            errors = e2e_step(1 + i)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * spp * args.steps / float(dt.item()), "unit": "sims/s",
               "h2d_bytes_per_step": 8 * (N_ASSETS + N_ASSETS * N_ASSETS),
               "d2h_bytes_per_step": spp * (2 * 8 + 4),
               "api": "mcos_b200.simulate_errors (host numpy in/out; pinned truth uploaded every step)"}

    # ---- the other BASELINE configurations and the fixed 100 000-simulation job, same process, after the timed region ----
    from mcos_b200 import engine as engine_mod

    def drop_engines():
        for e in list(engine_mod._ENGINES.values()):
            e.close()
        engine_mod._ENGINES.clear()

    configs, job = None, None
    if not args.no_configs:
        drop_engines()                 # the headline engine holds a full-wave arena; every config gets HBM to itself
        configs = run_other_configs(args, mb, local, rank, world, barrier, drop_engines)
        drop_engines()
        # BASELINE config 3 as stated: ONE job of 100 000 simulations through the public API from a cold engine (context
        # creation, scratch allocation, truth upload + Cholesky, ideal allocations, ragged last wave and the gather of
        # errors + status included), wall clock, max over ranks: STRONG scaling next to the weak-scaling `value`
        n_job = args.job_sims
        sim_j = mb.MuCovLedoitWolfObservationSimulator(mu, cov, N_OBS, seed=args.seed + 1, device=local)
        barrier()
        t0 = time.perf_counter()
        df = mb.simulate_optimizations(sim_j, n_job, optimizers, estimator, transformers, device=local)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        job = {"n_sims": n_job, "wall_s": float(dt.item()), "sims_per_s": n_job / float(dt.item()), "scaling": "strong",
               "api": "mcos_b200.simulate_optimizations from a cold engine", "n_gpus": world,
               "mean": {k: float(v) for k, v in df["mean"].items()}, "stdev": {k: float(v) for k, v in df["stdev"].items()}}

    # ---- roofline of the dominant kernel -----------------------------------------------------------------------------------
    hbm, hbm_src, fp64 = load_peaks()
    models = kernel_models(N_ASSETS, N_OBS, cov)
    sims_timed = spp          # the profiling pass is one step
    table = {}
    for name, (ms, cnt) in ktimes.items():
        if not cnt or name not in models:
            continue
        if name == "eigvec":   # only simulations with more than 16 signal eigenvalues reach it: no per-simulation model
            continue
        bound, work, unit = models[name]
        rate = work * sims_timed / (ms * 1e-3)
        if bound == "hbm":
            ach, peak, u = rate / 1e9, hbm, "GB/s"
        elif bound == "tensor":
            ach, peak, u = rate / 1e12, fp64["dmma_tflops"], "TFLOP/s"
        elif bound == "fp64":
            ach, peak, u = rate / 1e12, fp64["dfma_tflops"], "TFLOP/s"
        else:
            ach, peak, u = rate / 1e12, fp64["dmul_dadd_tinstr"], "Tinstr/s"
        table[name] = {"ms": round(ms, 3), "launches": cnt, "share": None, "bound": bound, "achieved": ach, "peak": peak,
                       "unit": u, "frac": ach / peak}
        if name == "sy2sb_pass":   # the same launches against the FP64 tensor roofline (4 N^3 / 3 flop)
            table[name]["frac_tensor"] = 4.0 * N_ASSETS ** 3 / 3 * sims_timed / (ms * 1e-3) / 1e12 / fp64["dmma_tflops"]
    tot = sum(v["ms"] for v in table.values()) or 1.0
    # launches that do next to nothing in this plan (the Ledoit-Wolf coefficient kernel: the shrink itself is folded into
    # the de-noiser's correlation pass; HRP's exact-distance kernel, which only flagged simulations enter) have no
    # meaningful roofline entry
    table = {k: v for k, v in table.items() if v["ms"] / tot >= 2e-3}
    for v in table.values():
        v["share"] = round(v["ms"] / tot, 4)
    top = max(table, key=lambda k: table[k]["ms"])
    tv = table[top]
    traffic = None   # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
    traffic_note = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "dominant_traffic.json")))[top]
        if tr["n_assets"] == N_ASSETS:
            traffic = tr["dram_bytes_per_matrix"] * spp
            traffic_note = tr["note"]
    except Exception:
        pass
    roofline = {"kernel": top + "_kernel", "bound": tv["bound"] if tv["bound"] in ("hbm", "tensor") else "fp64",
                "achieved": tv["achieved"], "peak": tv["peak"], "unit": tv["unit"], "frac": tv["frac"], "traffic": traffic,
                "traffic_note": traffic_note,
                "share_of_step": tv["share"], "ms_per_launch": tv["ms"] / tv["launches"],
                "peak_source": ("MEASURED_PEAKS.json (%s)" % hbm_src) if tv["bound"] == "hbm"
                else "profiles/fp64_peaks_r01.json (measured on this pool; MEASURED_PEAKS.json has no FP64 entry)"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = min(os.cpu_count() or 1, args.cpu_cores or (os.cpu_count() or 1))
        try:
            n, dt = cpu_reference_pass(cores, 12345)
            cpu_baseline = {"value": n / dt, "unit": "sims/s", "cores": cores, "kind": "port",
                            "sample": f"{n} simulations of the same workload, one per worker process (BLAS threads = 1), "
                                      f"{dt:.1f} s wall"}
        except Exception as exc:  # pragma: no cover
            cpu_baseline = {"value": None, "unit": "sims/s", "cores": cores, "kind": "port", "sample": f"failed: {exc}"}

    if rank == 0:
        line = {
            "metric": "Monte Carlo sims/sec (N=500,T=1000, DeNoise+Markowitz+HRP)", "value": value, "unit": "sims/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_assets": N_ASSETS, "n_observations": N_OBS, "sims_per_step_per_gpu": spp,
                       "job_100k_sims_steps": -(-100000 // (spp * world)), "parallelism": f"sim-id sharding x{world}",
                       "l2": "working set per step (>= 4 N^2 x 8 B x sims = %.1f GB) is far larger than the 126 MB L2; "
                             "no flush needed" % (4 * N_ASSETS * N_ASSETS * 8 * spp / 1e9),
                       "seed": args.seed},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches * world, "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "stage_ms_last_step": {k: round(v, 3) for k, v in stage_ms.items()},
            "kernels": table, "hrp_exact_fallback_sims": hrp_fallback, "hrp_sims_timed": spp * args.steps,
            "status_nonzero_last_step": status_bad, "nan_fraction_last_step": nan_frac,
            "configs": configs, "job_100k": job,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--sims-per-step", type=int, default=0,
                    help="simulations per GPU and step; default 48 x SM count (7104 on B200): one wave of the fused loop, a "
                         "whole number of CTA rounds for every kernel that runs 1, 2, 3, 4, 6, 8 or 16 simulations per SM")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-cores", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs 1, 2, 4, 5 and the 100k-simulation job")
    ap.add_argument("--job-sims", type=int, default=100000)
    ap.add_argument("--config5-sims", type=int, default=2048, help="config 5: simulations per GPU in the bounded sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
