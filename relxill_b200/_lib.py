"""ctypes binding of librelxill_b200.so.  No fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librelxill_b200.so")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")

_lib = None

LMOD_SYMBOLS = {
    "relline": "lmodrelline",
    "relconv": "lmodrelconv",
    "relline_lp": "lmodrellinelp",
    "relconv_lp": "lmodrelconvlp",
    "relxill": "lmodrelxill",
    "relxilllp": "lmodrelxilllp",
    "relxillCp": "lmodrelxilldensnthcomp",
    "relxilllpCp": "lmodrelxilllpdensnthcomp",
    "xillver": "lmodxillver",
    "xillverCp": "lmodxillverdensnthcomp",
    "xillverNS": "lmodxillverns",
    "relxillNS": "lmodrelxillns",
    "xillverCO": "lmodxillverco",
    "relxillCO": "lmodrelxillco",
}

ABI_SYMBOLS = [
    "relxill_b200_init", "relxill_b200_shutdown", "relxill_b200_set_num_zones", "relxill_b200_num_params",
    "relxill_b200_default_params", "relxill_b200_last_error", "relxill_batch_eval", "relxill_batch_eval_device",
    "relxill_b200_prepare", "relxill_b200_run", "relxill_b200_batch_status", "relxill_b200_free_batch",
    "relxill_b200_algorithmic_bytes", "relxill_b200_last_launches", "relxill_b200_set_profiling", "relxill_b200_keep_intermediates",
    "relxill_b200_kernel_times", "relxill_b200_probe", "relxill_b200_update_params", "relxill_b200_update_energy",
    "relxill_b200_reuse_counts", "relxill_b200_set_cache", "relxill_b200_set_xill_grid", "relxill_b200_get_xill_grid", "relxill_b200_set_xill_generic",
    "relxill_b200_last_eval_reuse", "relxill_b200_init_devices", "relxill_b200_num_devices", "relxill_b200_set_sharding", "relxill_b200_prepare_on",
    "relxill_b200_measure_fp64_peak",
] + sorted(LMOD_SYMBOLS.values())


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m relxill_b200.build` "
            "(there is no CPU fallback for the relxill_b200 hot path)")
    L = C.CDLL(LIB_PATH)
    L.relxill_b200_init.argtypes = [C.c_char_p, C.c_int]
    L.relxill_b200_init.restype = C.c_int
    L.relxill_b200_set_num_zones.argtypes = [C.c_int]
    L.relxill_b200_num_params.argtypes = [C.c_char_p]
    L.relxill_b200_default_params.argtypes = [C.c_char_p, _dp]
    L.relxill_b200_last_error.restype = C.c_char_p
    L.relxill_batch_eval.argtypes = [C.c_char_p, _dp, C.c_int, _dp, C.c_long, _dp, _ip]
    L.relxill_batch_eval.restype = C.c_int
    L.relxill_batch_eval_device.argtypes = [C.c_char_p, _dp, C.c_int, _dp, C.c_long, C.c_void_p, _ip, C.c_void_p]
    L.relxill_batch_eval_device.restype = C.c_int
    L.relxill_b200_prepare.argtypes = [C.c_char_p, _dp, C.c_int, _dp, C.c_long]
    L.relxill_b200_prepare.restype = C.c_void_p
    L.relxill_b200_prepare_on.argtypes = [C.c_int, C.c_char_p, _dp, C.c_int, _dp, C.c_long]
    L.relxill_b200_prepare_on.restype = C.c_void_p
    L.relxill_b200_init_devices.argtypes = [C.c_char_p, C.c_int]
    L.relxill_b200_init_devices.restype = C.c_int
    L.relxill_b200_num_devices.restype = C.c_int
    L.relxill_b200_set_sharding.argtypes = [C.c_int]
    L.relxill_b200_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.relxill_b200_run.restype = C.c_int
    L.relxill_b200_batch_status.argtypes = [C.c_void_p, _ip]
    L.relxill_b200_free_batch.argtypes = [C.c_void_p]
    L.relxill_b200_algorithmic_bytes.argtypes = [C.c_void_p, _dp]
    L.relxill_b200_last_launches.argtypes = [C.c_void_p]
    L.relxill_b200_last_launches.restype = C.c_long
    L.relxill_b200_set_profiling.argtypes = [C.c_int]
    L.relxill_b200_keep_intermediates.argtypes = [C.c_int]
    L.relxill_b200_kernel_times.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), _dp, np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS"), C.c_int]
    L.relxill_b200_kernel_times.restype = C.c_int
    L.relxill_b200_probe.argtypes = [C.c_void_p, C.c_long, C.c_char_p, _dp, C.c_long]
    L.relxill_b200_probe.restype = C.c_int
    L.relxill_b200_update_params.argtypes = [C.c_void_p, _dp]
    L.relxill_b200_update_params.restype = C.c_int
    L.relxill_b200_update_energy.argtypes = [C.c_void_p, _dp, C.c_int]
    L.relxill_b200_update_energy.restype = C.c_int
    L.relxill_b200_reuse_counts.argtypes = [C.c_void_p, np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")]
    L.relxill_b200_reuse_counts.restype = C.c_int
    L.relxill_b200_last_eval_reuse.argtypes = [np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")]
    L.relxill_b200_last_eval_reuse.restype = C.c_int
    L.relxill_b200_set_cache.argtypes = [C.c_int]
    L.relxill_b200_set_xill_grid.argtypes = [C.c_int]
    L.relxill_b200_get_xill_grid.restype = C.c_int
    L.relxill_b200_set_xill_generic.argtypes = [C.c_int]
    L.relxill_b200_measure_fp64_peak.restype = C.c_double
    for sym in LMOD_SYMBOLS.values():
        f = getattr(L, sym)
        f.argtypes = [_dp, C.c_int, _dp, C.c_int, _dp, C.c_void_p, C.c_char_p]
        f.restype = None
    _lib = L
    return L


def last_error() -> str:
    return lib().relxill_b200_last_error().decode()
