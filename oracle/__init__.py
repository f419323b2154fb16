"""ctypes binding of the CPU oracle (oracle/sa_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs; the product package ``spectrum_analyzer_b200`` never
imports it.  The oracle restates the reference's CPU path (see sa_oracle.h for the
file:line map and the parity status of each part).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

WINDOW_NONE, WINDOW_HANN, WINDOW_HAMMING, WINDOW_BH4, WINDOW_BH7 = range(5)
LIMIT_ALL, LIMIT_MIN, LIMIT_MAX, LIMIT_RANGE = range(4)
SCALING_NONE, SCALING_DIVIDE_BY_N, SCALING_DIVIDE_BY_N_SQRT, SCALING_ZERO_TO_ONE, SCALING_20_TIMES_LOG10 = range(5)

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
ERR_UNSUPPORTED_LENGTH = 10
ERR_NON_FINITE_RESULT = 11


class Stats(C.Structure):
    _fields_ = [("min_val", C.c_float), ("min_bin", C.c_uint32), ("max_val", C.c_float),
                ("max_bin", C.c_uint32), ("average", C.c_float), ("median", C.c_float)]


STATS_DTYPE = np.dtype([("min_val", "<f4"), ("min_bin", "<u4"), ("max_val", "<f4"),
                        ("max_bin", "<u4"), ("average", "<f4"), ("median", "<f4")])


def build(native: bool = False, force: bool = False) -> str:
    """Compile the oracle with gcc (``make`` in oracle/).  ``native`` adds -march=native
    (used only for the timed CPU baseline, on the box it is timed on)."""
    name = "libsa_oracle_native.so" if native else "libsa_oracle.so"
    path = os.path.join(_HERE, name)
    src = [os.path.join(_HERE, f) for f in ("sa_oracle.c", "sa_oracle.h", "Makefile")]
    stale = (not os.path.exists(path)) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in src)
    if force or stale:
        if native:
            cmd = ["gcc", "-O3", "-march=native", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
                   "-std=c11", "-shared", "-o", path, os.path.join(_HERE, "sa_oracle.c"), "-lm"]
        else:
            cmd = ["make", "-C", _HERE, "-B", "libsa_oracle.so"]
        subprocess.run(cmd, check=True, capture_output=True)
    return path


_libs: dict[bool, C.CDLL] = {}


def lib(native: bool = False) -> C.CDLL:
    if native in _libs:
        return _libs[native]
    L = C.CDLL(build(native))
    f32p = C.POINTER(C.c_float)
    L.sao_cosf.restype = C.c_float; L.sao_cosf.argtypes = [C.c_float]
    L.sao_log10f.restype = C.c_float; L.sao_log10f.argtypes = [C.c_float]
    L.sao_window_apply.argtypes = [C.c_int, f32p, C.c_uint32, f32p]
    L.sao_window_coeffs.argtypes = [C.c_int, C.c_uint32, f32p]
    L.sao_fft_calc.argtypes = [f32p, C.c_uint32, f32p]
    L.sao_limit_verify.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float]
    L.sao_frequency_resolution.restype = C.c_float
    L.sao_frequency_resolution.argtypes = [C.c_uint32, C.c_uint32]
    L.sao_limit_to_bins.restype = None
    L.sao_limit_to_bins.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_float, C.c_float,
                                    C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.sao_calc_statistics.restype = None
    L.sao_calc_statistics.argtypes = [f32p, C.c_uint32, C.c_uint32, C.POINTER(Stats)]
    L.sao_scale.restype = C.c_float
    L.sao_scale.argtypes = [C.c_int] + [C.c_float] * 6
    L.sao_samples_fft_to_spectrum.argtypes = [f32p, C.c_uint64, C.c_uint32, C.c_int, C.c_float, C.c_float,
                                              C.c_int, f32p, f32p, C.POINTER(C.c_uint32),
                                              C.POINTER(Stats), f32p]
    L.sao_spectrum_batched.argtypes = [f32p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_int,
                                       C.c_float, C.c_float, C.c_int, C.c_int, f32p, C.c_uint64,
                                       C.c_void_p, C.POINTER(C.c_uint32), C.c_int]
    L.sao_apply_scaling_chain_batched.argtypes = [f32p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32,
                                                  C.POINTER(C.c_int), C.c_uint32, C.c_void_p,
                                                  C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), f32p, C.c_int]
    L.sao_sine_wave_audio_data_multiple.restype = C.c_uint64
    L.sao_sine_wave_audio_data_multiple.argtypes = [f32p, C.c_uint32, C.c_uint32, C.c_uint32,
                                                    C.POINTER(C.c_int16), C.c_uint64]
    L.sao_generate_noise.restype = None
    L.sao_generate_noise.argtypes = [f32p, C.c_uint64, C.c_uint64, C.c_uint64]
    L.sao_max_threads.restype = C.c_int
    _libs[native] = L
    return L


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def cosf(x: float) -> float:
    return float(lib().sao_cosf(np.float32(x)))


def log10f(x: float) -> float:
    return float(lib().sao_log10f(np.float32(x)))


def window_coeffs(kind: int, n: int) -> np.ndarray:
    out = np.empty(n, np.float32)
    assert lib().sao_window_coeffs(kind, n, _p(out)) == 0
    return out


def window_apply(kind: int, samples) -> np.ndarray:
    x = _f32(samples)
    out = np.empty_like(x)
    assert lib().sao_window_apply(kind, _p(x), x.size, _p(out)) == 0
    return out


def fft_calc(samples) -> np.ndarray:
    """N/2+1 complex64 bins, as FftImpl::calc returns them."""
    x = _f32(samples)
    out = np.empty(x.size + 2, np.float32)
    code = lib().sao_fft_calc(_p(x), x.size, _p(out))
    if code:
        raise ValueError(f"oracle fft_calc error {code}")
    return out.view(np.complex64)


def limit_verify(kind: int, lo: float, hi: float, nyquist: float) -> int:
    return lib().sao_limit_verify(kind, lo, hi, nyquist)


def frequency_resolution(sampling_rate: int, n: int) -> float:
    return float(lib().sao_frequency_resolution(sampling_rate, n))


def limit_to_bins(n: int, sampling_rate: int, kind: int, lo: float = 0.0, hi: float = 0.0):
    a, b = C.c_uint32(), C.c_uint32()
    lib().sao_limit_to_bins(n, sampling_rate, kind, lo, hi, C.byref(a), C.byref(b))
    return a.value, b.value


def calc_statistics(vals, bin0: int = 0) -> Stats:
    v = _f32(vals)
    st = Stats()
    lib().sao_calc_statistics(_p(v), v.size, bin0, C.byref(st))
    return st


def scale(kind: int, v: float, smin=0.0, smax=0.0, savg=0.0, smedian=0.0, n=0.0) -> float:
    return float(lib().sao_scale(kind, v, smin, smax, savg, smedian, n))


@dataclass
class Spectrum:
    code: int
    freqs: np.ndarray
    vals: np.ndarray
    stats: Stats | None
    scale_err: tuple[float, float] | None = None


def samples_fft_to_spectrum(samples, sampling_rate: int, limit_kind: int = LIMIT_ALL, lo: float = 0.0,
                            hi: float = 0.0, scaling_kind: int = SCALING_NONE) -> Spectrum:
    """Oracle twin of `samples_fft_to_spectrum` (src/lib.rs:157); `samples` already windowed."""
    x = _f32(samples)
    cap = x.size // 2 + 2
    freqs = np.zeros(cap, np.float32)
    vals = np.zeros(cap, np.float32)
    nb = C.c_uint32(0)
    st = Stats()
    serr = np.zeros(2, np.float32)
    xp = _p(x) if x.size else C.cast(None, C.POINTER(C.c_float))
    code = lib().sao_samples_fft_to_spectrum(xp, x.size, sampling_rate, limit_kind, lo, hi, scaling_kind,
                                             _p(freqs), _p(vals), C.byref(nb), C.byref(st), _p(serr))
    m = nb.value
    return Spectrum(code, freqs[:m].copy(), vals[:m].copy(), st if code == OK else None,
                    (float(serr[0]), float(serr[1])) if code == ERR_SCALING else None)


def spectrum_batched(samples, n_frames: int, n: int, frame_stride: int, sampling_rate: int,
                     limit_kind: int = LIMIT_ALL, lo: float = 0.0, hi: float = 0.0,
                     window_kind: int = WINDOW_NONE, scaling_kind: int = SCALING_NONE,
                     threads: int = 0, native: bool = False, out_stride: int | None = None):
    """Returns (code, vals[n_frames, n_bins], stats recarray, status[n_frames], bin_lo)."""
    x = _f32(samples).reshape(-1)
    assert x.size >= (n_frames - 1) * frame_stride + n if n_frames else True
    bin_lo, n_bins = limit_to_bins(n, sampling_rate, limit_kind, lo, hi) if n >= 2 and not (n & (n - 1)) else (0, 0)
    stride = out_stride or max(n // 2 + 1, 1)
    vals = np.zeros((n_frames, stride), np.float32)
    stats = np.zeros(n_frames, STATS_DTYPE)
    status = np.zeros(n_frames, np.uint32)
    L = lib(native)
    if threads <= 0:
        threads = L.sao_max_threads()
    code = L.sao_spectrum_batched(_p(x), n_frames, n, frame_stride, sampling_rate, limit_kind, lo, hi,
                                  window_kind, scaling_kind, _p(vals), stride,
                                  stats.ctypes.data_as(C.c_void_p),
                                  status.ctypes.data_as(C.POINTER(C.c_uint32)), threads)
    return code, vals[:, :n_bins], stats, status, bin_lo


def apply_scaling_chain_batched(vals, bin0: int, n: int, kinds, stats, status, threads: int = 0):
    """Oracle twin of `FrequencySpectrum::apply_scaling_fn(&combined(kinds))` (src/spectrum.rs:128-168) on a batch.
    vals [frames, m] (copied), stats (STATS_DTYPE, copied), status (copied).
    Returns (vals, stats, status, err_index, err_pair[frames, 2])."""
    v = np.array(vals, dtype=np.float32, order="C", copy=True)
    frames, m = v.shape
    st = np.array(stats, dtype=STATS_DTYPE, copy=True)
    code = np.array(status, dtype=np.uint32, copy=True)
    err_index = np.zeros(frames, np.uint32)
    err_pair = np.zeros((frames, 2), np.float32)
    k = (C.c_int * len(kinds))(*kinds)
    L = lib()
    if threads <= 0:
        threads = L.sao_max_threads()
    L.sao_apply_scaling_chain_batched(_p(v), frames, m, m, bin0, n, k, len(kinds), st.ctypes.data_as(C.c_void_p),
                                      code.ctypes.data_as(C.POINTER(C.c_uint32)),
                                      err_index.ctypes.data_as(C.POINTER(C.c_uint32)), _p(err_pair), threads)
    return v, st, code, err_index, err_pair


def sine_wave_audio_data_multiple(frequencies, sampling_rate: int, duration_ms: int) -> np.ndarray:
    f = _f32(frequencies)
    cap = int(sampling_rate * (duration_ms / 1000.0)) + 8
    out = np.zeros(cap, np.int16)
    n = lib().sao_sine_wave_audio_data_multiple(_p(f), f.size, sampling_rate, duration_ms,
                                                out.ctypes.data_as(C.POINTER(C.c_int16)), cap)
    return out[:n].copy()


def generate_noise(first_index: int, count: int, seed: int) -> np.ndarray:
    out = np.empty(count, np.float32)
    lib().sao_generate_noise(_p(out), first_index, count, seed)
    return out


def max_threads() -> int:
    return lib().sao_max_threads()
