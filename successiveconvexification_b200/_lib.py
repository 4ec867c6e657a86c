"""ctypes binding of libscvx_b200.so (C ABI: include/scvx_b200.h).

This mirrors, symbol for symbol, the `ccall` shim in julia/SCvxB200.jl, so the tests exercise the
same ABI a Julia host would.  There is NO fallback: if the shared library is missing the import of
the device path fails loudly (build it with `python successiveconvexification_b200/csrc/build.py`).
"""
from __future__ import annotations

import ctypes
import os

from .defns import CProbInfo

_HERE = os.path.dirname(os.path.abspath(__file__))
# SCVX_B200_LIB selects another build of the same library (kernel-variant A/B runs of profiles/build_variants.py)
LIB_PATH = os.environ.get("SCVX_B200_LIB") or os.path.join(_HERE, "libscvx_b200.so")

_dp = ctypes.POINTER(ctypes.c_double)
_ctx_p = ctypes.c_void_p

# name -> (restype, argtypes): every symbol include/scvx_b200.h declares
SIGNATURES = {
    "scvx_device_count": (ctypes.c_int, []),
    "scvx_last_error": (ctypes.c_char_p, []),
    "scvx_version": (ctypes.c_int, []),
    "scvx_sizeof_probinfo": (ctypes.c_int, []),
    "scvx_create": (ctypes.c_int, [ctypes.POINTER(_ctx_p), ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "scvx_destroy": (None, [_ctx_p]),
    "scvx_set_params": (ctypes.c_int, [_ctx_p, ctypes.POINTER(CProbInfo), ctypes.c_int]),
    "scvx_set_aero_table": (ctypes.c_int, [_ctx_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                           ctypes.c_int]),
    "scvx_get_aero_coefficients": (ctypes.c_int, [_ctx_p, ctypes.c_int, ctypes.c_void_p]),
    "scvx_linearize_batch": (ctypes.c_int, [_ctx_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "scvx_linearize_batch_fins": (ctypes.c_int, [_ctx_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_void_p, ctypes.c_void_p]),
    "scvx_set_fin_table": (ctypes.c_int, [_ctx_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int]),
    "scvx_fin_force_batch": (ctypes.c_int, [_ctx_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_void_p]),
    "scvx_predict_batch": (ctypes.c_int, [_ctx_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_void_p]),
    "scvx_compact_record_doubles": (ctypes.c_int, [ctypes.c_int]),
    "scvx_linearize_batch_compact": (ctypes.c_int, [_ctx_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                    ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                    ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "scvx_compact_layout": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int32)]),
    "scvx_expand_compact": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "scvx_host_alloc": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint64]),
    "scvx_host_free": (ctypes.c_int, [ctypes.c_void_p]),
    "scvx_host_register": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64]),
    "scvx_host_unregister": (ctypes.c_int, [ctypes.c_void_p]),
    "scvx_defect_cost_batch": (ctypes.c_int, [_ctx_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]),
    "scvx_linear_points_batch": (ctypes.c_int, [_ctx_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double,
                                                ctypes.c_double, _dp, _dp, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_void_p, ctypes.c_void_p]),
    "scvx_dispersed_setup_batch": (ctypes.c_int, [_ctx_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "scvx_socp_dims": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                      ctypes.POINTER(ctypes.c_int)]),
    "scvx_socp_pattern": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]),
    "scvx_socp_values_batch": (ctypes.c_int, [_ctx_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "scvx_set_stream": (ctypes.c_int, [_ctx_p, ctypes.c_void_p]),
    "scvx_set_kernel": (ctypes.c_int, [_ctx_p, ctypes.c_int]),
    "scvx_synchronize": (ctypes.c_int, [_ctx_p]),
    "scvx_launch_count": (ctypes.c_int64, [_ctx_p]),
    "scvx_last_kernel_ms": (ctypes.c_int, [_ctx_p, _dp]),
}

_LIB = None


class ScvxError(RuntimeError):
    """A non-zero return code of the C ABI (the Julia shim raises `error(...)` the same way,
    matching the reference's exception style, rocketland.jl:275)."""


def load():
    """Load the shared library and declare every prototype.  Raises if the library is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the B200 path has no CPU fallback. Build it with "
            f"`python successiveconvexification_b200/csrc/build.py` (nvcc, sm_100a).")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.scvx_sizeof_probinfo() != ctypes.sizeof(CProbInfo):
        raise ImportError("scvx_probinfo layout mismatch between the binding and the library")
    _LIB = lib
    return lib


TOOLS_PATH = os.path.join(_HERE, "libscvx_benchtools.so")
_TOOLS = None


def load_benchtools():
    """Measurement helpers of bench.py / profiles/ (DFMA peak, DMMA probe): a separate library, not the product ABI."""
    global _TOOLS
    if _TOOLS is None:
        if not os.path.exists(TOOLS_PATH):
            raise ImportError(f"{TOOLS_PATH} is missing: build it with `python successiveconvexification_b200/csrc/build.py`")
        lib = ctypes.CDLL(TOOLS_PATH)
        lib.scvx_bench_fp64_peak.restype = ctypes.c_int
        lib.scvx_bench_fp64_peak.argtypes = [ctypes.c_int, _dp]
        lib.scvx_bench_dmma_probe.restype = ctypes.c_int
        lib.scvx_bench_dmma_probe.argtypes = [ctypes.c_int, _dp]
        _TOOLS = lib
    return _TOOLS


def check(rc: int):
    if rc != 0:
        msg = load().scvx_last_error()
        raise ScvxError(f"scvx error {rc}: {msg.decode() if msg else '?'}")
