"""Thin host wrapper around one mcosb context (one GPU, one problem shape).

Arrays may be numpy (host; the library stages them through device scratch) or torch CUDA tensors on the engine's
device (passed by data_ptr(); the call is then asynchronous on torch's current stream).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import MCOSBError, Plan

try:  # torch is plumbing only: device tensors, streams, torch.distributed
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


class Engine:
    """One `mcosb_ctx`. Not thread-safe; one per GPU."""

    def __init__(self, n_assets: int, n_observations: int, device: int = 0, seed: int = 0):
        self.lib = _lib.load()
        self.N, self.T, self.device, self.seed = int(n_assets), int(n_observations), int(device), int(seed)
        handle = ctypes.c_void_p()
        rc = self.lib.mcosb_create(ctypes.byref(handle), self.device, self.N, self.T, ctypes.c_uint64(self.seed))
        if rc != 0:
            raise MCOSBError(f"mcosb_create failed with code {rc}: a CUDA device is required (no CPU fallback)")
        self.h = handle
        self.mu = self.cov = None
        self.truth_pd = False

    def close(self):
        if getattr(self, "h", None):
            self.lib.mcosb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers -------------------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            msg = self.lib.mcosb_last_error(self.h)
            raise MCOSBError(f"mcosb error {rc}: {msg.decode() if msg else ''}")

    def _in(self, x, shape, dtype=np.float64):
        """Returns (keepalive, pointer) for an input of the given shape."""
        if x is None:
            return None, None
        if _is_torch(x):
            want = torch.float64 if dtype == np.float64 else torch.int32
            assert x.is_cuda and x.device.index == self.device and x.dtype == want and x.is_contiguous()
            assert tuple(x.shape) == tuple(shape), (tuple(x.shape), shape)
            return x, ctypes.c_void_p(x.data_ptr())
        a = np.ascontiguousarray(np.asarray(x, dtype=dtype).reshape(shape))
        return a, a.ctypes.data_as(ctypes.c_void_p)

    def _out(self, like_torch, shape, dtype=np.float64):
        if like_torch:
            # device tensors in, device tensors out: nothing synchronises with the host, so the kernels must run on the
            # stream the caller's torch work is on
            self.use_torch_stream()
            t = torch.empty(shape, dtype=torch.float64 if dtype == np.float64 else torch.int32,
                            device=f"cuda:{self.device}")
            return t, ctypes.c_void_p(t.data_ptr())
        a = np.empty(shape, dtype=dtype)
        return a, a.ctypes.data_as(ctypes.c_void_p)

    def use_torch_stream(self):
        """Binds the context to torch's current CUDA stream on this device.  torch's default stream has handle 0, which
        mcosb_set_stream reads as "the context's own stream": it is passed as cudaStreamLegacy (0x1) instead, so that the
        library's kernels are ordered with the torch work around them."""
        handle = torch.cuda.current_stream(self.device).cuda_stream or 0x1
        self._check(self.lib.mcosb_set_stream(self.h, ctypes.c_void_p(handle)))

    def synchronize(self):
        self._check(self.lib.mcosb_synchronize(self.h))

    @property
    def launch_count(self):
        return int(self.lib.mcosb_launch_count(self.h))

    def last_stage_ms(self):
        buf = (ctypes.c_double * _lib.N_STAGES)()
        self._check(self.lib.mcosb_last_stage_ms(self.h, buf))
        return dict(zip(_lib.STAGE_NAMES, list(buf)))

    def set_profiling(self, on=True):
        self._check(self.lib.mcosb_set_profiling(self.h, int(bool(on))))

    def kernel_times(self):
        """{kernel name: (total ms, launches)} since set_profiling(True)."""
        n = self.lib.mcosb_kernel_count()
        ms = (ctypes.c_double * n)()
        cnt = (ctypes.c_longlong * n)()
        self._check(self.lib.mcosb_kernel_times(self.h, ms, cnt))
        return {self.lib.mcosb_kernel_name(i).decode(): (ms[i], int(cnt[i])) for i in range(n)}

    # ---- stages --------------------------------------------------------------------------------------------------
    def set_truth(self, mu, cov, require_pd=True):
        """Uploads the true (mu, cov). With require_pd=False a covariance that has no Cholesky factor is accepted
        (the error estimators only need mu and cov themselves); sampling from it would still fail."""
        mu_a = np.ascontiguousarray(np.asarray(mu, dtype=np.float64).reshape(self.N))
        cov_a = np.ascontiguousarray(np.asarray(cov, dtype=np.float64).reshape(self.N, self.N))
        if self.mu is not None and np.array_equal(mu_a, self.mu) and np.array_equal(cov_a, self.cov) and \
                (self.truth_pd or not require_pd):
            return
        self.mu, self.cov = mu_a, cov_a
        rc = self.lib.mcosb_set_truth(self.h, mu_a.ctypes.data_as(ctypes.c_void_p),
                                      cov_a.ctypes.data_as(ctypes.c_void_p))
        self.truth_pd = rc == 0
        if rc == -4 and not require_pd:
            return
        self._check(rc)

    def ensure_truth(self, mu, cov):
        self.set_truth(mu, cov, require_pd=True)

    def sample(self, n_sims, sim_id0=0, out_torch=False):
        x, px = self._out(out_torch, (n_sims, self.T, self.N))
        self._check(self.lib.mcosb_sample(self.h, ctypes.c_uint64(sim_id0), n_sims, px))
        return x

    def moments(self, x, kind=_lib.MOMENTS_MUCOV):
        s = x.shape[0]
        tt = _is_torch(x)
        kx, px = self._in(x, (s, self.T, self.N))
        mu, pmu = self._out(tt, (s, self.N))
        cov, pcov = self._out(tt, (s, self.N, self.N))
        sh, psh = self._out(tt, (s,))
        self._check(self.lib.mcosb_moments(self.h, s, px, kind, pmu, pcov, psh))
        return mu, cov, sh

    def sample_moments(self, n_sims, sim_id0=0, kind=_lib.MOMENTS_MUCOV, out_torch=False):
        """(mu_hat, cov_hat, shrink) of simulations sim_id0 .. as mcosb_run forms them (draws kept on the device, centred by
        the true mean): the stage call that composes with denoise / markowitz / ... to mcosb_run bit for bit."""
        mu, pmu = self._out(out_torch, (n_sims, self.N))
        cov, pcov = self._out(out_torch, (n_sims, self.N, self.N))
        sh, psh = self._out(out_torch, (n_sims,))
        self._check(self.lib.mcosb_sample_moments(self.h, ctypes.c_uint64(sim_id0), n_sims, kind, pmu, pcov, psh))
        return mu, cov, sh

    def denoise(self, cov, n_observations=None, bandwidth=0.25):
        s = cov.shape[0]
        tt = _is_torch(cov)
        work = cov.clone() if tt else np.array(cov, dtype=np.float64, order="C", copy=True)
        kc, pc = self._in(work, (s, self.N, self.N))
        eig, pe = self._out(tt, (s, self.N))
        var, pv = self._out(tt, (s,))
        nf, pn = self._out(tt, (s,), np.int32)
        st, ps = self._out(tt, (s,), np.int32)
        t = self.T if n_observations is None else int(n_observations)
        self._check(self.lib.mcosb_denoise(self.h, s, pc, t, float(bandwidth), pe, pv, pn, ps))
        return kc, {"eigvals": eig, "var": var, "n_facts": nf, "status": st}

    def set_hrp_mode(self, mode: int):
        """0 = tensor-core distances + guarded Prim + exact fallback (default), 1 = exact distances for every simulation,
        2 = fast pass with every simulation sent to the fallback, 3 = the fast-path variant used above 1024 assets (test aids)."""
        self._check(self.lib.mcosb_set_hrp_mode(self.h, int(mode)))

    def hrp_fallback_count(self) -> int:
        """Simulations that took HRP's exact fallback since the last call."""
        c = ctypes.c_int(0)
        self._check(self.lib.mcosb_hrp_fallback_count(self.h, ctypes.byref(c)))
        return c.value

    def set_eig_method(self, method: int):
        """0 = automatic, 1 = one-stage Householder tridiagonalisation, 2 = two-stage (band reduction + bulge chasing)."""
        self._check(self.lib.mcosb_set_eig_method(self.h, int(method)))

    def tridiagonalize(self, sym, want_work=False):
        """(d, e[, work]) of the symmetric matrices sym[S][N][N]: the first half of np.linalg.eigh on the GPU."""
        s = sym.shape[0]
        tt = _is_torch(sym)
        ks, ps = self._in(sym, (s, self.N, self.N))
        d, pd_ = self._out(tt, (s, self.N))
        e, pe = self._out(tt, (s, self.N))
        wk, pw = self._out(tt, (s, self.N, self.N)) if want_work else (None, None)
        self._check(self.lib.mcosb_tridiagonalize(self.h, s, ps, pd_, pe, pw))
        return (d, e, wk) if want_work else (d, e)

    def detone(self, cov, n_remove):
        s = cov.shape[0]
        tt = _is_torch(cov)
        work = cov.clone() if tt else np.array(cov, dtype=np.float64, order="C", copy=True)
        kc, pc = self._in(work, (s, self.N, self.N))
        st, ps = self._out(tt, (s,), np.int32)
        self._check(self.lib.mcosb_detone(self.h, s, pc, int(n_remove), ps))
        return kc, {"status": st}

    def markowitz(self, mu, cov, rf=0.02, clean=True):
        s = cov.shape[0]
        tt = _is_torch(cov)
        km, pm = self._in(mu, (s, self.N))
        kc, pc = self._in(cov, (s, self.N, self.N))
        w, pw = self._out(tt, (s, self.N))
        it, pi = self._out(tt, (s,), np.int32)
        st, ps = self._out(tt, (s,), np.int32)
        self._check(self.lib.mcosb_markowitz(self.h, s, pm, pc, float(rf), int(bool(clean)), pw, pi, ps))
        return w, {"iters": it, "status": st}

    def hrp(self, cov):
        s = cov.shape[0]
        tt = _is_torch(cov)
        kc, pc = self._in(cov, (s, self.N, self.N))
        w, pw = self._out(tt, (s, self.N))
        order, po = self._out(tt, (s, self.N), np.int32)
        link, pl = self._out(tt, (s, max(self.N - 1, 0), 4))
        st, ps = self._out(tt, (s,), np.int32)
        self._check(self.lib.mcosb_hrp(self.h, s, pc, pw, po, pl, ps))
        return w, {"order": order, "linkage": link, "status": st}

    def nco(self, mu, cov, max_num_clusters=None, num_clustering_trials=10, labels=None):
        s = cov.shape[0]
        tt = _is_torch(cov)
        km, pm = self._in(mu, (s, self.N))
        kc, pc = self._in(cov, (s, self.N, self.N))
        kl, plab = self._in(labels, (s, self.N), np.int32)
        w, pw = self._out(tt, (s, self.N))
        lab, plo = self._out(tt, (s, self.N), np.int32)
        st, ps = self._out(tt, (s,), np.int32)
        max_k = 0 if max_num_clusters is None else int(max_num_clusters)
        self._check(self.lib.mcosb_nco(self.h, s, pm, pc, max_k, int(num_clustering_trials), plab, pw, plo, ps))
        return w, {"labels": lab, "status": st}

    def risk_parity(self, cov, budget=None):
        s = cov.shape[0]
        tt = _is_torch(cov)
        kc, pc = self._in(cov, (s, self.N, self.N))
        kb, pb = self._in(budget, (self.N,))
        w, pw = self._out(tt, (s, self.N))
        it, pi = self._out(tt, (s,), np.int32)
        st, ps = self._out(tt, (s,), np.int32)
        self._check(self.lib.mcosb_risk_parity(self.h, s, pc, pb, pw, pi, ps))
        return w, {"iters": it, "status": st}

    def errors(self, w, w_star, kind=_lib.ERR_VARIANCE):
        s = w.shape[0]
        tt = _is_torch(w)
        kw, pw = self._in(w, (s, self.N))
        batched = int(np.prod(tuple(w_star.shape)) == s * self.N and not (s == 1 and len(tuple(w_star.shape)) == 1))
        kws, pws = self._in(w_star, (s, self.N) if batched else (self.N,))
        err, pe = self._out(tt, (s,))
        self._check(self.lib.mcosb_errors(self.h, s, pw, pws, batched, kind, pe))
        return err

    def errors_multi(self, w, w_star, kinds):
        """err[S, len(kinds)]: several estimators from one pass over the weights (column k bit-equal to errors(kinds[k]))."""
        s = w.shape[0]
        tt = _is_torch(w)
        kw, pw = self._in(w, (s, self.N))
        batched = int(np.prod(tuple(w_star.shape)) == s * self.N and not (s == 1 and len(tuple(w_star.shape)) == 1))
        kws, pws = self._in(w_star, (s, self.N) if batched else (self.N,))
        kk = (ctypes.c_int * len(kinds))(*[int(k) for k in kinds])
        err, pe = self._out(tt, (s, len(kinds)))
        self._check(self.lib.mcosb_errors_multi(self.h, s, pw, pws, batched, len(kinds), kk, pe))
        return err

    def run(self, plan: Plan, n_sims, sim_id0=0, out_torch=False):
        shape = (n_sims, plan.n_optimizers, plan.n_error_kinds) if plan.n_error_kinds else (n_sims, plan.n_optimizers)
        err, pe = self._out(out_torch, shape)
        st, ps = self._out(out_torch, (n_sims,), np.int32)
        rc = self.lib.mcosb_run(self.h, ctypes.byref(plan), ctypes.c_uint64(sim_id0), n_sims, pe, ps)
        if rc == _lib.ERR_IDEAL:  # optimizer.allocate(true mu, true cov) raises in the reference (mcos.py:30)
            bits = int((self.lib.mcosb_last_error(self.h) or b"0").decode().split()[-1])
            if bits & _lib.ST_HRP_SUM:
                raise ValueError("Portfolio allocations don't sum to 1.")
        self._check(rc)
        return err, st


_ENGINES = {}


def get_engine(n_assets, n_observations, device=0, seed=0) -> Engine:
    """Process-wide engine cache keyed by (device, N, T, seed)."""
    key = (int(device), int(n_assets), int(n_observations), int(seed))
    eng = _ENGINES.get(key)
    if eng is None:
        eng = _ENGINES[key] = Engine(n_assets, n_observations, device, seed)
    return eng
