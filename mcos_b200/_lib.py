"""ctypes binding of libmcosb200.so (the C ABI declared in include/mcosb.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is usable, the first call raises.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_longlong, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# MCOSB_LIB_PATH: an alternative build of the same library (timing / ablation variants for A/B runs)
LIB_PATH = os.environ.get("MCOSB_LIB_PATH") or os.path.join(_HERE, "libmcosb200.so")

N_STAGES = 7
STAGE_NAMES = ("sample", "moments", "denoise", "markowitz", "hrp", "nco", "errors")

MOMENTS_MUCOV, MOMENTS_LEDOIT_WOLF, MOMENTS_JACKKNIFE = 0, 1, 2
ERR_EXPECTED_OUTCOME, ERR_VARIANCE, ERR_SHARPE = 0, 1, 2
OPT_MARKOWITZ, OPT_HRP, OPT_NCO, OPT_RISK_PARITY = 0, 1, 2, 3
TR_DENOISE, TR_DETONE = 0, 1
MAX_TRANSFORMERS = 4

ERR_IDEAL = -5
ST_NOT_PD, ST_NO_EXCESS, ST_QP_CAP, ST_MPFIT_FAIL, ST_EIGVEC, ST_HRP_SUM, ST_NONFINITE = 1, 2, 4, 8, 16, 32, 64
ST_NAMES = {ST_NOT_PD: "NOT_PD", ST_NO_EXCESS: "NO_EXCESS", ST_QP_CAP: "QP_CAP", ST_MPFIT_FAIL: "MPFIT_FAIL",
            ST_EIGVEC: "EIGVEC", ST_HRP_SUM: "HRP_SUM", ST_NONFINITE: "NONFINITE"}


class MCOSBError(RuntimeError):
    pass


class Plan(ctypes.Structure):
    _fields_ = [("moments_kind", c_int), ("denoise", c_int), ("bandwidth", c_double), ("n_optimizers", c_int),
                ("optimizers", c_int * 4), ("error_kind", c_int), ("risk_free_rate", c_double),
                ("nco_max_clusters", c_int), ("nco_trials", c_int), ("wave", c_int), ("detone_n_remove", c_int),
                ("n_error_kinds", c_int), ("error_kinds", c_int * 3), ("n_transformers", c_int),
                ("transformer_kind", c_int * MAX_TRANSFORMERS), ("transformer_arg", c_double * MAX_TRANSFORMERS),
                ("risk_budget", c_void_p)]


# name -> (restype, argtypes); every symbol include/mcosb.h declares
SIGNATURES = {
    "mcosb_version": (c_int, []),
    "mcosb_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_uint64]),
    "mcosb_destroy": (c_int, [c_void_p]),
    "mcosb_last_error": (c_char_p, [c_void_p]),
    "mcosb_set_stream": (c_int, [c_void_p, c_void_p]),
    "mcosb_synchronize": (c_int, [c_void_p]),
    "mcosb_set_truth": (c_int, [c_void_p, c_void_p, c_void_p]),
    "mcosb_sample": (c_int, [c_void_p, c_uint64, c_int, c_void_p]),
    "mcosb_moments": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "mcosb_sample_moments": (c_int, [c_void_p, c_uint64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mcosb_denoise": (c_int, [c_void_p, c_int, c_void_p, c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcosb_set_eig_method": (c_int, [c_void_p, c_int]),
    "mcosb_tridiagonalize": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcosb_detone": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "mcosb_markowitz": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_double, c_int, c_void_p, c_void_p, c_void_p]),
    "mcosb_hrp": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcosb_set_hrp_mode": (c_int, [c_void_p, c_int]),
    "mcosb_hrp_fallback_count": (c_int, [c_void_p, POINTER(c_int)]),
    "mcosb_nco": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcosb_risk_parity": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcosb_errors": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "mcosb_errors_multi": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, POINTER(c_int), c_void_p]),
    "mcosb_run": (c_int, [c_void_p, POINTER(Plan), c_uint64, c_int, c_void_p, c_void_p]),
    "mcosb_last_stage_ms": (c_int, [c_void_p, POINTER(c_double)]),
    "mcosb_launch_count": (c_longlong, [c_void_p]),
    "mcosb_set_profiling": (c_int, [c_void_p, c_int]),
    "mcosb_kernel_count": (c_int, []),
    "mcosb_kernel_name": (c_char_p, [c_int]),
    "mcosb_kernel_times": (c_int, [c_void_p, POINTER(c_double), POINTER(c_longlong)]),
}

_lib = None


def load():
    """Loads the library once and types every entry point. Raises MCOSBError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MCOSBError(f"{LIB_PATH} is missing: build it with `make -C mcos_b200/csrc` "
                         "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise MCOSBError(f"{LIB_PATH} does not export {name}") from exc
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib
