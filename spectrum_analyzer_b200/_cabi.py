"""ctypes view of the C ABI declared in include/spectrum_analyzer_cuda.h.

The shared library is REQUIRED: importing this module raises if it has not been built
(``python -c "import __graft_entry__ as g; g.build()"``).  There is no Python/CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._build import LIB_PATH

# enums (values identical to the header)
WINDOW_NONE, WINDOW_HANN, WINDOW_HAMMING, WINDOW_BLACKMAN_HARRIS_4TERM, WINDOW_BLACKMAN_HARRIS_7TERM = range(5)
LIMIT_ALL, LIMIT_MIN, LIMIT_MAX, LIMIT_RANGE = range(4)
SCALING_NONE, SCALING_DIVIDE_BY_N, SCALING_DIVIDE_BY_N_SQRT, SCALING_ZERO_TO_ONE, SCALING_20_TIMES_LOG10 = range(5)
STATS_MINMAX, STATS_AVERAGE, STATS_MEDIAN, STATS_ALL = 1, 2, 4, 7

OK = 0
ERR_TOO_FEW_SAMPLES = 1
ERR_NAN_VALUES = 2
ERR_INFINITY_VALUES = 3
ERR_NOT_POWER_OF_TWO = 4
ERR_LIMIT_BELOW_MINIMUM = 5
ERR_LIMIT_ABOVE_NYQUIST = 6
ERR_LIMIT_INVALID_RANGE = 7
ERR_LIMIT_TOO_NARROW = 8
ERR_SCALING = 9
SCALING_ERR_INDEX_MASK, SCALING_ERR_NAN, SCALING_ERR_POS_INF, SCALING_ERR_NEG_INF = 0xffff, 0x80000000, 0x40000000, 0x20000000
ERR_UNSUPPORTED_LENGTH = 10
ERR_NON_FINITE_RESULT = 11
ERR_INVALID_ARGUMENT = 100
ERR_CUDA = 101
ERR_OUT_OF_MEMORY = 102


class FrameStats(C.Structure):
    _fields_ = [("min_val", C.c_float), ("min_bin", C.c_uint32), ("max_val", C.c_float),
                ("max_bin", C.c_uint32), ("average", C.c_float), ("median", C.c_float),
                ("status", C.c_uint32), ("reserved", C.c_uint32)]


FRAME_STATS_DTYPE = np.dtype([("min_val", "<f4"), ("min_bin", "<u4"), ("max_val", "<f4"), ("max_bin", "<u4"),
                              ("average", "<f4"), ("median", "<f4"), ("status", "<u4"), ("reserved", "<u4")])
assert FRAME_STATS_DTYPE.itemsize == 32 and C.sizeof(FrameStats) == 32

# every symbol the header declares: name -> (restype, argtypes)
_vp, _u32, _u64, _i, _f = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_float
_u32p = C.POINTER(C.c_uint32)
SYMBOLS = {
    "sa_abi_version": (_u32, []),
    "sa_status_string": (C.c_char_p, [_i]),
    "sa_last_error_string": (C.c_char_p, []),
    "sa_ctx_create": (_i, [_i, C.POINTER(_vp)]),
    "sa_ctx_create_on_stream": (_i, [_i, _vp, C.POINTER(_vp)]),
    "sa_ctx_destroy": (_i, [_vp]),
    "sa_ctx_sync": (_i, [_vp]),
    "sa_ctx_set_fast_log10": (_i, [_vp, _i]),
    "sa_ctx_launch_count": (_u64, [_vp]),
    "sa_ctx_last_kernel": (C.c_char_p, [_vp]),
    "sa_window_coeffs": (_i, [_i, _u32, _vp]),
    "sa_limit_to_bins": (_i, [_u32, _u32, _i, _f, _f, _u32p, _u32p]),
    "sa_bin_frequencies": (_i, [_u32, _u32, _u32, _u32, _vp]),
    "sa_spectrum_batched": (_i, [_vp, _vp, _u64, _u32, _u64, _u32, _i, _f, _f, _i, _i, _u32, _vp, _u64, _vp,
                                 _u32p, _u32p]),
    "sa_spectrum_batched_host": (_i, [_vp, _vp, _u64, _u32, _u64, _u32, _i, _f, _f, _i, _i, _u32, _vp, _u64, _vp,
                                      _u32p, _u32p]),
    "sa_spectrum_batched_i16": (_i, [_vp, _vp, _u64, _u32, _u64, _u32, _i, _f, _f, _i, _i, _u32, _vp, _u64, _vp,
                                     _u32p, _u32p]),
    "sa_spectrum_batched_host_i16": (_i, [_vp, _vp, _u64, _u32, _u64, _u32, _i, _f, _f, _i, _i, _u32, _vp, _u64, _vp,
                                          _u32p, _u32p]),
    "sa_spectrum_single": (_i, [_vp, _vp, _u64, _u32, _i, _f, _f, _i, _i, _vp, _vp, _u32p, _vp]),
    "sa_apply_scaling_batched": (_i, [_vp, _vp, _u64, _u32, _u64, _u32, _u32, _i, _u32, _vp]),
    "sa_apply_scaling_chain_batched": (_i, [_vp, _vp, _u64, _u32, _u64, _u32, _u32, _vp, _u32, _u32, _vp]),
    "sa_apply_scaling_chain_host": (_i, [_vp, _vp, _u64, _u32, _u64, _u32, _u32, _vp, _u32, _u32, _vp]),
    "sa_window_batched": (_i, [_vp, _i, _vp, _u64, _u32, _u64, _vp]),
    "sa_window_host": (_i, [_vp, _i, _vp, _u64, _u32, _u64, _vp]),
    "sa_generate_noise": (_i, [_vp, _vp, _u64, _u64, _u64]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen the CUDA library; loud failure when it is missing (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it first (python -c 'import __graft_entry__ as g; g.build()'). "
            "spectrum_analyzer_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.sa_abi_version() != 1:
        raise ImportError("libspectrum_analyzer_cuda.so ABI version mismatch")
    _lib = lib
    return lib


def status_string(code: int) -> str:
    return load().sa_status_string(code).decode()


def last_error() -> str:
    return load().sa_last_error_string().decode()
