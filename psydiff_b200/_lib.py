"""ctypes binding of libpsydiff_b200.so (C ABI in include/psydiff_b200.h).

The library is built in-tree by ``psydiff_b200.build.build_library()`` (called from ``__graft_entry__.build``).
There is no Python or CPU fallback for the compute entry points: if the shared object is missing, or no sm_100
device is present when a compute call is made, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpsydiff_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")
CACHE_DIR = os.environ.get("PSY_KERNEL_CACHE", os.path.join(_HERE, "_kernel_cache"))

_lib = None


class PsydiffError(RuntimeError):
    """Raised where the reference raises an R error (Rcpp::stop / BEGIN_RCPP, src/RcppExports.cpp:18-26)."""


c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_int64_p = C.POINTER(C.c_int64)
c_ubyte_p = C.POINTER(C.c_ubyte)


class psy_container(C.Structure):
    """include/psydiff_b200.h: one entry of model$pars$parameterList as the code generator sees it."""
    _fields_ = [("name", C.c_char_p), ("is_scalar", C.c_int), ("rows", C.c_int), ("cols", C.c_int),
                ("values", c_double_p), ("slots", c_int_p)]


c_container_p = C.POINTER(psy_container)
c_size_p = C.POINTER(C.c_size_t)
c_uint_p = C.POINTER(C.c_uint)

# name -> (restype, argtypes); kept in one table so tests can check every symbol of include/psydiff_b200.h
SIGNATURES = {
    "psy_last_error": (C.c_char_p, []),
    "psy_version": (C.c_int, []),
    "psy_rk4_step_count": (C.c_int, [C.c_double, C.c_double]),
    "psy_emit_model_source": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p, c_container_p, C.c_int, C.c_int, C.c_int,
                                        C.c_char_p, C.c_size_t, c_size_p]),
    "psy_model_template": (C.c_int, [C.c_int, C.c_char_p, C.c_size_t, c_size_p]),
    "psy_drift_dependency": (C.c_int, [C.c_int, C.c_char_p, c_container_p, C.c_int, c_uint_p]),
    "psy_model_compile": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t]),
    "psy_model_create": (C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_char_p]),
    "psy_model_destroy": (None, [C.c_void_p]),
    "psy_model_load_module": (C.c_int, [C.c_void_p, C.c_char_p]),
    "psy_device_count": (C.c_int, []),
    "psy_model_set_devices": (C.c_int, [C.c_void_p, C.c_int, c_int_p]),
    "psy_model_devices": (C.c_int, [C.c_void_p, c_int_p, C.c_int]),
    "psy_shard_info": (C.c_int, [C.c_void_p, C.c_int, c_int_p, c_int_p, c_int_p, c_int64_p, c_double_p]),
    "psy_model_settings": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, c_double_p, C.c_int, C.c_int]),
    "psy_upload": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, c_double_p, c_double_p, c_int64_p, c_int_p]),
    "psy_set_slots": (C.c_int, [C.c_void_p, C.c_int, c_int_p]),
    "psy_fit": (C.c_int, [C.c_void_p, c_double_p, C.c_int, c_double_p, c_double_p, c_double_p]),
    "psy_fit_many": (C.c_int, [C.c_void_p, c_double_p, C.c_int, C.c_int, c_double_p, c_double_p]),
    "psy_gradients": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int, C.c_int, c_double_p, c_double_p]),
    "psy_gradient": (C.c_int, [C.c_void_p, c_double_p, C.c_int, c_double_p, C.c_int, C.c_int, c_double_p]),
    "psy_partial_len": (C.c_int64, [C.c_void_p]),
    "psy_grad_level": (C.c_int, [C.c_void_p, c_double_p, C.c_double, C.c_int, c_ubyte_p, C.POINTER(C.c_void_p)]),
    "psy_grad_resolve": (C.c_int, [C.c_void_p, c_double_p, C.c_double, C.c_int, C.c_int, C.c_int, c_ubyte_p, c_int_p]),
    "psy_grad_finish": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p]),
    "psy_sweep_create": (C.c_void_p, [C.c_int]),
    "psy_sweep_destroy": (None, [C.c_void_p]),
    "psy_sweep_resolve": (C.c_int, [C.c_void_p, c_double_p, C.c_double, C.c_int, C.c_int, C.c_int, c_ubyte_p, c_int_p]),
    "psy_sweep_finish": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p]),
    "psy_last_kernel_ms": (C.c_double, [C.c_void_p]),
    "psy_last_kernel_instances": (C.c_int64, [C.c_void_p]),
    "psy_launch_count": (C.c_int64, []),
    "psy_counters_reset": (None, []),
    "psy_n_instances": (C.c_int64, [C.c_void_p]),
    "psy_measure_fp64_peak": (C.c_int, [c_double_p]),
    "psy_kernel_info": (C.c_int, [C.c_void_p, c_int_p, c_int_p, c_int_p]),
    "psy_kernel_lanes": (C.c_int, [C.c_void_p]),
    # primitives (src/exposeInlineFunctions.cpp)
    "psy_getMeanWeights": (None, [C.c_int, C.c_double, C.c_double, C.c_double, c_double_p]),
    "psy_getCovWeights": (None, [C.c_int, C.c_double, C.c_double, C.c_double, c_double_p]),
    "psy_getWMatrix": (None, [C.c_int, c_double_p, c_double_p, c_double_p]),
    "psy_getSigmaPoints": (C.c_int, [C.c_int, c_double_p, c_double_p, C.c_int, C.c_double, c_double_p]),
    "psy_computeMeanFromSigmaPoints": (None, [C.c_int, C.c_int, c_double_p, c_double_p, c_double_p]),
    "psy_getAFromSigmaPoints": (None, [C.c_int, c_double_p, c_double_p, C.c_double, c_double_p]),
    "psy_computePFromSigmaPoints": (None, [C.c_int, c_double_p, c_double_p, C.c_double, c_double_p]),
    "psy_getPhi": (None, [C.c_int, c_double_p, c_double_p]),
    "psy_predictMu": (None, [C.c_int, C.c_int, c_double_p, c_double_p, c_double_p]),
    "psy_predictS": (None, [C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    "psy_predictC": (None, [C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    "psy_computeK": (C.c_int, [C.c_int, C.c_int, c_double_p, c_double_p, c_double_p]),
    "psy_updateM": (None, [C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    "psy_updateP": (None, [C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    "psy_computeIndividualM2LL": (C.c_int, [C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    "psy_computeIndividualM2LLChol": (C.c_int, [C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    "psy_cholupdate": (C.c_int, [C.c_int, c_double_p, c_double_p, C.c_double, C.c_char_p]),
    "psy_qr": (None, [C.c_int, C.c_int, c_double_p, c_double_p]),
    "psy_predictSquareRootOfS": (None, [C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, C.c_double, C.c_double, c_double_p]),
    "psy_computeK_SR": (C.c_int, [C.c_int, C.c_int, c_double_p, c_double_p, c_double_p]),
    "psy_logChol2Chol": (None, [C.c_int, c_double_p, c_double_p]),
}


def lib():
    """Load the shared library (once).  Raises if it has not been built — there is nothing to fall back to."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PsydiffError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(psydiff_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise PsydiffError(lib().psy_last_error().decode("utf-8", "replace"))


def dptr(a):
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    return a.ctypes.data_as(c_int_p)


def i64ptr(a):
    return a.ctypes.data_as(c_int64_p)


def u8ptr(a):
    return a.ctypes.data_as(c_ubyte_p)
