"""ctypes binding of libpfb200.so (C ABI in include/pfb200.h).

This is the binding a maintainer of a host-language front end adds (INTEGRATION.md shows the
Julia ``ccall`` twin).  There is NO CPU fallback: if the CUDA library is missing or no GPU is
present, every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PFB200_LIB") or os.path.join(_HERE, "libpfb200.so")  # PFB200_LIB: A/B builds

PFB_MODEL_ISONORMAL = 0
PFB_MODEL_FUNNEL = 1
PFB_MODEL_DIAGNORMAL = 2
PFB_MODEL_DENSENORMAL = 3
PFB_MODEL_HLOGISTIC = 4
PFB_MODEL_HOSTCALLBACK = 5


class PfbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libpfb200 error {code}: {msg}")
        self.code = code


class pfb_config(C.Structure):
    _fields_ = [
        ("device", C.c_int32),
        ("history_length", C.c_int32),
        ("ndraws_elbo", C.c_int32),
        ("materialize_all", C.c_int32),
        ("elbo_mode", C.c_int32),
        ("reserved", C.c_int32),
        ("eps", C.c_double),
    ]


_dp = C.c_void_p
# void cb(void* user, const double* x /* n x m column-major */, int64_t n, int64_t m, double* logp_out)
pfb_logp_callback = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_double), C.c_int64, C.c_int64, C.POINTER(C.c_double))


class pfb_elbo_out(C.Structure):
    _fields_ = [(name, _dp) for name in (
        "elbo", "elbo_se", "logp", "logq", "best_iter", "success", "n_rejected", "draws",
        "draws_logp", "draws_logq", "fit_mu", "fit_alpha", "fit_vh", "fit_T", "fit_Vc",
        "fit_logdet", "fit_jeff", "all_draws")]


class pfb_resample_out(C.Structure):
    _fields_ = [(name, _dp) for name in (
        "log_weights", "weights", "pareto_k", "tail_len", "inds", "ids", "draws")]


class pfb_lbfgs_opts(C.Structure):
    _fields_ = [("maxiters", C.c_int32), ("max_points", C.c_int32), ("gtol", C.c_double), ("ftol", C.c_double)]


class pfb_device_view(C.Structure):
    _fields_ = [
        ("pool_draws", _dp), ("pool_logp", _dp), ("pool_logq", _dp), ("elbo", _dp), ("stream", _dp),
        ("n", C.c_int64), ("K", C.c_int64), ("P", C.c_int64), ("U", C.c_int64),
    ]


# every symbol include/pfb200.h declares: (restype, argtypes)
SYMBOLS = {
    "pfb_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(pfb_config)]),
    "pfb_destroy": (C.c_int, [C.c_void_p]),
    "pfb_last_error": (C.c_char_p, [C.c_void_p]),
    "pfb_kp": (C.c_int, [C.c_void_p]),
    "pfb_register_model": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp, C.c_size_t]),
    "pfb_register_host_model": (C.c_int, [C.c_void_p, C.c_int, pfb_logp_callback, C.c_void_p]),
    "pfb_elbo_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp,
                                 C.POINTER(pfb_elbo_out)]),
    "pfb_batch_upload": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]),
    "pfb_batch_run": (C.c_int, [C.c_void_p]),
    "pfb_batch_sync": (C.c_int, [C.c_void_p]),
    "pfb_batch_download": (C.c_int, [C.c_void_p, C.POINTER(pfb_elbo_out)]),
    "pfb_lbfgs_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp, C.POINTER(pfb_lbfgs_opts), _dp, _dp, _dp]),
    "pfb_batch_from_lbfgs": (C.c_int, [C.c_void_p, _dp]),
    "pfb_lbfgs_download": (C.c_int, [C.c_void_p, _dp, _dp, _dp]),
    "pfb_lbfgs_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "pfb_batch_fit_only": (C.c_int, [C.c_void_p, _dp]),
    "pfb_draw_from_fits": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, C.c_int]),
    "pfb_unit_draws": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp]),
    "pfb_unit_fits": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "pfb_set_fallback_seeds": (C.c_int, [C.c_void_p, C.c_int, _dp]),
    "pfb_comm_unique_id": (C.c_int, [_dp]),
    "pfb_comm_init": (C.c_int, [C.c_void_p, _dp, C.c_int, C.c_int]),
    "pfb_comm_init_all": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "pfb_comm_destroy": (C.c_int, [C.c_void_p]),
    "pfb_pool_exchange_resample": (C.c_int, [C.c_void_p, _dp, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                             C.POINTER(pfb_resample_out)]),
    "pfb_pool_exchange_resample_all": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, _dp, C.c_uint64, C.c_int, C.c_int,
                                                 C.c_int, C.POINTER(pfb_resample_out)]),
    "pfb_pool_set": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp]),
    "pfb_pool_materialize": (C.c_int, [C.c_void_p]),
    "pfb_pool_columns_device": (C.c_int, [C.c_void_p, C.c_int, _dp, C.c_int64, _dp]),
    "pfb_pool_download": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp]),
    "pfb_psis_resample": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                    C.POINTER(pfb_resample_out)]),
    "pfb_psis_resample_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, _dp, _dp,
                                         C.c_uint64, C.c_int, C.c_int, C.c_int, C.POINTER(pfb_resample_out)]),
    "pfb_batch_device_view": (C.c_int, [C.c_void_p, C.POINTER(pfb_device_view)]),
    "pfb_psis_resample_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, _dp, _dp, _dp,
                                           C.c_uint64, C.c_int, C.c_int, C.c_int, C.POINTER(pfb_resample_out)]),
    "pfb_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "pfb_host_unregister": (C.c_int, [C.c_void_p]),
    "pfb_get_timings": (C.c_int, [C.c_void_p, _dp]),
    "pfb_measure_fp64_fma_tflops": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "pfb_measure_fp64_dmma_tflops": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """Load libpfb200.so and bind every declared symbol.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PfbError(-100, f"{LIB_PATH} not found: build it with `python -c 'import "
                             f"__graft_entry__ as g; g.build()'` or `make -C pathfinder_b200/csrc -j`")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(handle, rc):
    if rc != 0:
        msg = load().pfb_last_error(handle)
        raise PfbError(rc, msg.decode() if msg else "")
