"""Host-side mirror of the reference crate's public API for the hot path, on top of the C ABI.

Names, argument meaning and error behaviour follow phip1611/spectrum-analyzer v1.8.0:

* ``samples_fft_to_spectrum(samples, sampling_rate, frequency_limit, scaling_fn)``
  (reference src/lib.rs:157-162) -> :class:`FrequencySpectrum` (src/spectrum.rs:48-253)
* ``FrequencyLimit.{All, Min, Max, Range}`` (src/limit.rs:38-53)
* ``scaling.{divide_by_N, divide_by_N_sqrt, scale_to_zero_to_one, scale_20_times_log10}``
  (src/scaling.rs:104-162) -- tokens, not Python callables: arbitrary closures cannot run on
  the GPU and are rejected (``TypeError``), as BASELINE.json's north_star requires.
* ``windows.{hann_window, hamming_window, blackman_harris_4term, blackman_harris_7term}``
  (src/windows.rs:40,58,79,97)
* errors: :class:`SpectrumAnalyzerError` subclasses (src/error.rs:37-54)
* new: ``samples_fft_to_spectrum_batched`` -- the ``[frames x N]`` entry point.

Everything numeric happens in libspectrum_analyzer_cuda.so; this module only marshals.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from . import _cabi as A

__all__ = [
    "FrequencyLimit", "FrequencySpectrum", "BatchedSpectrum", "SpectrumAnalyzerError", "Context",
    "samples_fft_to_spectrum", "samples_fft_to_spectrum_batched", "scaling", "windows",
    "TooFewSamples", "NaNValuesNotSupported", "InfinityValuesNotSupported", "InvalidFrequencyLimit",
    "FrequencyLimitTooNarrow", "SamplesLengthNotAPowerOfTwo", "ScalingError", "UnsupportedLength",
    "NonFiniteResult", "CudaError",
]


# ---------------------------------------------------------------------------------------------
# errors (src/error.rs:37-54)
# ---------------------------------------------------------------------------------------------
class SpectrumAnalyzerError(Exception):
    code = -1

    def __init__(self, code: int | None = None, detail: str = ""):
        self.code = self.code if code is None else code
        msg = A.status_string(self.code)
        super().__init__(f"{msg}{': ' + detail if detail else ''}")


class TooFewSamples(SpectrumAnalyzerError): code = A.ERR_TOO_FEW_SAMPLES
class NaNValuesNotSupported(SpectrumAnalyzerError): code = A.ERR_NAN_VALUES
class InfinityValuesNotSupported(SpectrumAnalyzerError): code = A.ERR_INFINITY_VALUES
class SamplesLengthNotAPowerOfTwo(SpectrumAnalyzerError): code = A.ERR_NOT_POWER_OF_TWO
class InvalidFrequencyLimit(SpectrumAnalyzerError): pass      # ValueBelowMinimum / ValueAboveNyquist / InvalidRange
class FrequencyLimitTooNarrow(SpectrumAnalyzerError): code = A.ERR_LIMIT_TOO_NARROW
class ScalingError(SpectrumAnalyzerError): code = A.ERR_SCALING
class UnsupportedLength(SpectrumAnalyzerError): code = A.ERR_UNSUPPORTED_LENGTH       # reference panics (src/fft.rs:52)
class NonFiniteResult(SpectrumAnalyzerError): code = A.ERR_NON_FINITE_RESULT         # reference panics (src/frequency.rs:52)
class CudaError(SpectrumAnalyzerError): code = A.ERR_CUDA

_ERRORS = {
    A.ERR_TOO_FEW_SAMPLES: TooFewSamples, A.ERR_NAN_VALUES: NaNValuesNotSupported,
    A.ERR_INFINITY_VALUES: InfinityValuesNotSupported, A.ERR_NOT_POWER_OF_TWO: SamplesLengthNotAPowerOfTwo,
    A.ERR_LIMIT_BELOW_MINIMUM: InvalidFrequencyLimit, A.ERR_LIMIT_ABOVE_NYQUIST: InvalidFrequencyLimit,
    A.ERR_LIMIT_INVALID_RANGE: InvalidFrequencyLimit, A.ERR_LIMIT_TOO_NARROW: FrequencyLimitTooNarrow,
    A.ERR_SCALING: ScalingError, A.ERR_UNSUPPORTED_LENGTH: UnsupportedLength,
    A.ERR_NON_FINITE_RESULT: NonFiniteResult, A.ERR_CUDA: CudaError,
}


def _check(code: int) -> None:
    if code == A.OK:
        return
    cls = _ERRORS.get(code, SpectrumAnalyzerError)
    raise cls(code, A.last_error() if code == A.ERR_CUDA else "")


# ---------------------------------------------------------------------------------------------
# FrequencyLimit (src/limit.rs:38-118)
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class FrequencyLimit:
    kind: int
    lo: float = 0.0
    hi: float = 0.0

    @staticmethod
    def Min(x: float) -> "FrequencyLimit":
        return FrequencyLimit(A.LIMIT_MIN, float(x), 0.0)

    @staticmethod
    def Max(x: float) -> "FrequencyLimit":
        return FrequencyLimit(A.LIMIT_MAX, 0.0, float(x))

    @staticmethod
    def Range(lo: float, hi: float) -> "FrequencyLimit":
        return FrequencyLimit(A.LIMIT_RANGE, float(lo), float(hi))

    def maybe_min(self):
        return self.lo if self.kind in (A.LIMIT_MIN, A.LIMIT_RANGE) else None

    def maybe_max(self):
        return self.hi if self.kind in (A.LIMIT_MAX, A.LIMIT_RANGE) else None

    def to_bins(self, n: int, sampling_rate: int) -> tuple[int, int]:
        """(bin_lo, n_bins) of the kept range; raises the reference's error for this limit."""
        lo, nb = C.c_uint32(), C.c_uint32()
        _check(A.load().sa_limit_to_bins(n, sampling_rate, self.kind, self.lo, self.hi, C.byref(lo), C.byref(nb)))
        return lo.value, nb.value


FrequencyLimit.All = FrequencyLimit(A.LIMIT_ALL)  # class attribute, like the enum variant


# ---------------------------------------------------------------------------------------------
# scaling (src/scaling.rs:104-182): tokens for the built-ins
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class _ScalingFn:
    kind: int
    name: str

    def __call__(self, *a, **k):
        raise TypeError(f"scaling.{self.name} is a device-side built-in token, not a host callable")


@dataclass(frozen=True)
class _ScalingChain:
    kinds: tuple

    def __call__(self, *a, **k):
        raise TypeError("scaling.combined(...) is a device-side token, not a host callable")


class scaling:  # namespace, like the crate's `scaling` module
    @staticmethod
    def combined(fncs) -> "_ScalingChain":
        """`scaling::combined(&[f1, f2, ...])` (src/scaling.rs:174-182): every function of the chain sees the
        same pre-call statistics.  Usable with FrequencySpectrum.apply_scaling_fn."""
        kinds = tuple(_scaling_kind(f) for f in fncs)
        if len(kinds) > 8:
            raise ValueError("at most 8 functions in a combined chain")
        return _ScalingChain(kinds)

    divide_by_N = _ScalingFn(A.SCALING_DIVIDE_BY_N, "divide_by_N")
    divide_by_N_sqrt = _ScalingFn(A.SCALING_DIVIDE_BY_N_SQRT, "divide_by_N_sqrt")
    scale_to_zero_to_one = _ScalingFn(A.SCALING_ZERO_TO_ONE, "scale_to_zero_to_one")
    scale_20_times_log10 = _ScalingFn(A.SCALING_20_TIMES_LOG10, "scale_20_times_log10")


def _scaling_kind(fn) -> int:
    if fn is None:
        return A.SCALING_NONE
    if isinstance(fn, _ScalingFn):
        return fn.kind
    raise TypeError("only the built-in scaling functions (scaling.divide_by_N, divide_by_N_sqrt, "
                    "scale_to_zero_to_one, scale_20_times_log10) can run on the GPU; arbitrary closures "
                    "are rejected (no CPU fallback)")


# ---------------------------------------------------------------------------------------------
# context
# ---------------------------------------------------------------------------------------------
class Context:
    """One GPU + one stream (``sa_ctx``).  Not thread-safe: one per host thread."""

    def __init__(self, device: int = 0, stream: int | None = None):
        lib = A.load()
        self._h = C.c_void_p()
        if stream is None:
            _check(lib.sa_ctx_create(device, C.byref(self._h)))
        else:
            _check(lib.sa_ctx_create_on_stream(device, C.c_void_p(stream), C.byref(self._h)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            A.load().sa_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _check(A.load().sa_ctx_sync(self._h))

    def set_fast_log10(self, enable: bool = True) -> None:
        """Opt-in: fused ``scale_20_times_log10`` through the hardware log2 approximation (within 4 ulp of the
        libm sequence's f32 dB value, 1.5-1.7x faster in that mode).  Default off."""
        _check(A.load().sa_ctx_set_fast_log10(self._h, 1 if enable else 0))

    @property
    def launch_count(self) -> int:
        return int(A.load().sa_ctx_launch_count(self._h))

    @property
    def last_kernel(self) -> str:
        """Name (with template arguments) of the spectrum kernel the last call on this context launched."""
        return (A.load().sa_ctx_last_kernel(self._h) or b"").decode()


_default_ctx: dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


# ---------------------------------------------------------------------------------------------
# FrequencySpectrum (src/spectrum.rs:48-474)
# ---------------------------------------------------------------------------------------------
_F32_EPS = np.float32(1.1920929e-07)


def _approx_eq_ulps3(a: float, b: float) -> bool:
    """float_cmp::approx_eq!(f32, a, b, ulps = 3) (epsilon keeps its default, f32::EPSILON)."""
    a, b = np.float32(a), np.float32(b)
    if a == b:
        return True
    if np.isnan(a) or np.isnan(b):
        return False
    if abs(a - b) <= _F32_EPS:
        return True
    ia, ib = int(a.view(np.int32)), int(b.view(np.int32))
    if (ia < 0) != (ib < 0):
        return False
    return abs(ia - ib) <= 3


class FrequencySpectrum:
    """Result container with the reference's getters; values are f32 (numpy float32)."""

    def __init__(self, freqs: np.ndarray, vals: np.ndarray, frequency_resolution: float, samples_len: int,
                 stats, bin_lo: int, ctx: Context | None = None):
        self._freqs = np.asarray(freqs, np.float32)
        self._vals = np.asarray(vals, np.float32)
        self._res = np.float32(frequency_resolution)
        self._n = int(samples_len)
        self._bin_lo = int(bin_lo)
        self._stats = stats  # numpy record with FRAME_STATS_DTYPE fields
        self._ctx = ctx

    # -- scalar getters (src/spectrum.rs:173-253) --
    def average(self) -> np.float32: return np.float32(self._stats["average"])
    def median(self) -> np.float32: return np.float32(self._stats["median"])

    def max(self) -> tuple[np.float32, np.float32]:
        b = int(self._stats["max_bin"])
        return self._freqs[b - self._bin_lo], np.float32(self._stats["max_val"])

    def min(self) -> tuple[np.float32, np.float32]:
        b = int(self._stats["min_bin"])
        return self._freqs[b - self._bin_lo], np.float32(self._stats["min_val"])

    def range(self) -> np.float32: return np.float32(self.max()[1] - self.min()[1])
    def data(self) -> list[tuple[np.float32, np.float32]]: return list(zip(self._freqs, self._vals))
    def frequencies(self) -> np.ndarray: return self._freqs
    def values(self) -> np.ndarray: return self._vals
    def frequency_resolution(self) -> np.float32: return self._res
    def samples_len(self) -> int: return self._n
    def max_fr(self) -> np.float32: return self._freqs[-1]
    def min_fr(self) -> np.float32: return self._freqs[0]

    def dc_component(self):
        return self._vals[0] if self._freqs[0] == 0.0 else None

    # -- lookups (src/spectrum.rs:300-474), host-side on the finished spectrum --
    def _bounds_check(self, search_fr):
        mn, mx = self._freqs[0], self._freqs[-1]
        if search_fr < mn or search_fr > mx:
            raise IndexError(f"Frequency {search_fr}Hz is out of bounds [{mn}; {mx}]!")  # reference panics

    def freq_val_exact(self, search_fr: float) -> np.float32:
        search_fr = np.float32(search_fr)
        if _approx_eq_ulps3(self._freqs[0], search_fr): return self._vals[0]
        if _approx_eq_ulps3(self._freqs[-1], search_fr): return self._vals[-1]
        self._bounds_check(search_fr)
        i = int(np.searchsorted(self._freqs, search_fr, side="left"))  # first b with freq[b] >= search
        i = max(i, 1)
        ax, ay, bx, by = self._freqs[i - 1], self._vals[i - 1], self._freqs[i], self._vals[i]
        if _approx_eq_ulps3(ax, search_fr):
            return ay
        slope = np.float32((by - ay) / (bx - ax))        # src/spectrum.rs:584-588
        c = np.float32(ay - np.float32(slope * ax))
        return np.float32(np.float32(slope * search_fr) + c)

    def freq_val_closest(self, search_fr: float) -> tuple[np.float32, np.float32]:
        search_fr = np.float32(search_fr)
        if _approx_eq_ulps3(self._freqs[0], search_fr): return self._freqs[0], self._vals[0]
        if _approx_eq_ulps3(self._freqs[-1], search_fr): return self._freqs[-1], self._vals[-1]
        self._bounds_check(search_fr)
        i = max(int(np.searchsorted(self._freqs, search_fr, side="left")), 1)
        ax, bx = self._freqs[i - 1], self._freqs[i]
        if _approx_eq_ulps3(ax, search_fr):
            return ax, self._vals[i - 1]
        if np.float32(np.float32(search_fr - ax) / self._res) < 0.5:
            return ax, self._vals[i - 1]
        return bx, self._vals[i]

    def mel_val(self, mel: float) -> np.float32:
        assert mel >= 0.0
        hz = np.float32(700.0) * (np.float32(math.pow(10.0, float(np.float32(mel) / np.float32(2595.0)))) - np.float32(1.0))
        return self.freq_val_exact(hz)

    def to_map(self) -> dict[int, float]:
        return {int(f): float(v) for f, v in zip(self._freqs, self._vals)}

    def to_mel_map(self) -> dict[int, float]:
        mel = np.float32(2595.0) * np.log10(np.float32(1.0) + self._freqs / np.float32(700.0)).astype(np.float32)
        return {int(m): float(v) for m, v in zip(mel, self._vals)}

    # -- FrequencySpectrum::apply_scaling_fn (src/spectrum.rs:128-168), on the device --
    def apply_scaling_fn(self, scaling_fn) -> None:
        kinds = scaling_fn.kinds if isinstance(scaling_fn, _ScalingChain) else (_scaling_kind(scaling_fn),)
        ctx = self._ctx or default_context()
        vals = np.ascontiguousarray(self._vals, dtype=np.float32).copy()
        st = np.array([self._stats], dtype=A.FRAME_STATS_DTYPE)
        karr = (C.c_int * max(len(kinds), 1))(*kinds)
        st["status"] = A.OK          # a spectrum object exists only for frames that were OK
        _check(A.load().sa_apply_scaling_chain_host(ctx._h, vals.ctypes.data, 1, vals.size, vals.size, self._bin_lo,
                                                    self._n, karr, len(kinds), A.STATS_ALL, st.ctypes.data))
        # the values in front of an offender are replaced, like the reference's in-place loop (src/spectrum.rs:150-163)
        self._vals = vals
        if int(st["status"][0]) == A.ERR_SCALING:
            r = int(st["reserved"][0])
            i = r & A.SCALING_ERR_INDEX_MASK
            new = float("nan") if r & A.SCALING_ERR_NAN else float("inf") if r & A.SCALING_ERR_POS_INF else float("-inf")
            err = ScalingError(A.ERR_SCALING, f"ScalingError({float(vals[i])!r}, {new!r})")
            err.orig, err.new = float(vals[i]), new          # ScalingError(orig, new), src/error.rs:50-53
            raise err
        self._stats = st[0]


# ---------------------------------------------------------------------------------------------
# entry points
# ---------------------------------------------------------------------------------------------
def _window_kind(window) -> int:
    if window is None:
        return A.WINDOW_NONE
    if isinstance(window, int):
        return window
    return {"none": A.WINDOW_NONE, "hann": A.WINDOW_HANN, "hamming": A.WINDOW_HAMMING,
            "blackman_harris_4term": A.WINDOW_BLACKMAN_HARRIS_4TERM,
            "blackman_harris_7term": A.WINDOW_BLACKMAN_HARRIS_7TERM}[str(window).lower()]


def samples_fft_to_spectrum(samples, sampling_rate: int, frequency_limit: FrequencyLimit = FrequencyLimit.All,
                            scaling_fn=None, *, window=None, ctx: Context | None = None) -> FrequencySpectrum:
    """Drop-in for ``samples_fft_to_spectrum`` (src/lib.rs:157-162).  ``window=`` optionally fuses the
    caller-side ``windows::*`` step (the reference applies it before the call)."""
    x = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1)
    kind = _scaling_kind(scaling_fn)
    ctx = ctx or default_context()
    cap = x.size // 2 + 2
    freqs = np.zeros(cap, np.float32)
    vals = np.zeros(cap, np.float32)
    nb = C.c_uint32(0)
    st = np.zeros(1, A.FRAME_STATS_DTYPE)
    code = A.load().sa_spectrum_single(ctx._h, x.ctypes.data if x.size else None, x.size, sampling_rate,
                                       frequency_limit.kind, frequency_limit.lo, frequency_limit.hi,
                                       _window_kind(window), kind, freqs.ctypes.data, vals.ctypes.data,
                                       C.byref(nb), st.ctypes.data)
    _check(code)
    m = nb.value
    bin_lo = frequency_limit.to_bins(x.size, sampling_rate)[0]
    res = np.float32(sampling_rate) / np.float32(x.size)
    return FrequencySpectrum(freqs[:m].copy(), vals[:m].copy(), res, x.size, st[0], bin_lo, ctx)


@dataclass
class BatchedSpectrum:
    """SoA result of the batched entry point: ``vals[frames, n_bins]`` + one shared frequency vector
    + ``stats[frames]`` (numpy record array, or a torch uint8 tensor [frames, 32] for device calls)."""
    vals: object
    stats: object
    frequencies: np.ndarray
    bin_lo: int
    n_bins: int
    samples_len: int
    sampling_rate: int

    def stats_numpy(self) -> np.ndarray:
        if _is_torch(self.stats):
            return np.frombuffer(self.stats.cpu().numpy().tobytes(), A.FRAME_STATS_DTYPE)
        return self.stats

    def spectrum(self, i: int) -> FrequencySpectrum:
        """Reference-shaped view of frame i; raises the frame's error if its status is not OK."""
        st = self.stats_numpy()[i]
        _check(int(st["status"]))
        v = self.vals[i]
        v = v[: self.n_bins].cpu().numpy() if _is_torch(v) else np.asarray(v[: self.n_bins])
        res = np.float32(self.sampling_rate) / np.float32(self.samples_len)
        return FrequencySpectrum(self.frequencies, v, res, self.samples_len, st, self.bin_lo)


def samples_fft_to_spectrum_batched(frames, sampling_rate: int, frequency_limit: FrequencyLimit = FrequencyLimit.All,
                                    scaling_fn=None, *, window=None, n: int | None = None,
                                    frame_stride: int | None = None, n_frames: int | None = None,
                                    stats: int = A.STATS_ALL, out_stride: int | None = None,
                                    out=None, out_stats=None, ctx: Context | None = None) -> BatchedSpectrum:
    """Batched ``samples_fft_to_spectrum``: same arguments, ``frames`` is ``[n_frames, N]`` (or a flat
    stream with ``n``/``frame_stride``/``n_frames`` for STFT-style overlap).

    * numpy input  -> host pipeline (``sa_spectrum_batched_host``), numpy output
    * torch CUDA input -> device-resident call on torch's current stream (``sa_spectrum_batched``),
      torch output; asynchronous.
    * int16 input (numpy or torch) is PCM: converted `x as f32` inside the kernel's load
      (``sa_spectrum_batched_i16`` / ``sa_spectrum_batched_host_i16``), halving the input traffic.
    """
    kind = _scaling_kind(scaling_fn)
    wk = _window_kind(window)
    lib = A.load()
    if n is None:
        n = int(frames.shape[-1])
    if n_frames is None:
        n_frames = int(frames.shape[0]) if frames.ndim == 2 else (max(0, (frames.numel() if _is_torch(frames) else frames.size) - n) // (frame_stride or n) + 1)
    if frame_stride is None:
        frame_stride = n
    blo, nb = C.c_uint32(), C.c_uint32()
    _check(lib.sa_limit_to_bins(n, sampling_rate, frequency_limit.kind, frequency_limit.lo, frequency_limit.hi,
                                C.byref(blo), C.byref(nb)))
    n_bins, bin_lo = nb.value, blo.value
    stride = out_stride or n_bins
    freqs = np.zeros(n_bins, np.float32)
    _check(lib.sa_bin_frequencies(n, sampling_rate, bin_lo, n_bins, freqs.ctypes.data))

    if _is_torch(frames):
        import torch
        assert frames.is_cuda and frames.dtype in (torch.float32, torch.int16) and frames.is_contiguous()
        entry = lib.sa_spectrum_batched_i16 if frames.dtype == torch.int16 else lib.sa_spectrum_batched
        dev = frames.device
        if ctx is None:
            ctx = _torch_ctx(dev.index or 0, torch.cuda.current_stream(dev).cuda_stream)
        vals = out if out is not None else torch.empty((n_frames, stride), dtype=torch.float32, device=dev)
        st = out_stats if out_stats is not None else torch.empty((n_frames, 32), dtype=torch.uint8, device=dev)
        _check(entry(ctx._h, frames.data_ptr(), n_frames, n, frame_stride, sampling_rate,
                                       frequency_limit.kind, frequency_limit.lo, frequency_limit.hi, wk, kind, stats,
                                       vals.data_ptr(), stride, st.data_ptr() if stats else None, None, None))
        return BatchedSpectrum(vals, st, freqs, bin_lo, n_bins, n, sampling_rate)

    is_i16 = isinstance(frames, np.ndarray) and frames.dtype == np.int16
    x = np.ascontiguousarray(frames) if is_i16 else np.ascontiguousarray(frames, dtype=np.float32)
    host_entry = lib.sa_spectrum_batched_host_i16 if is_i16 else lib.sa_spectrum_batched_host
    ctx = ctx or default_context()
    vals = out if out is not None else np.zeros((n_frames, stride), np.float32)
    st = out_stats if out_stats is not None else np.zeros(n_frames, A.FRAME_STATS_DTYPE)
    _check(host_entry(ctx._h, x.ctypes.data, n_frames, n, frame_stride, sampling_rate,
                                        frequency_limit.kind, frequency_limit.lo, frequency_limit.hi, wk, kind, stats,
                                        vals.ctypes.data, stride, st.ctypes.data if stats else None, None, None))
    return BatchedSpectrum(vals, st, freqs, bin_lo, n_bins, n, sampling_rate)


_torch_ctxs: dict[tuple[int, int], Context] = {}


def _torch_ctx(device: int, stream: int) -> Context:
    key = (device, stream)
    if key not in _torch_ctxs:
        _torch_ctxs[key] = Context(device, stream)
    return _torch_ctxs[key]


# ---------------------------------------------------------------------------------------------
# windows (src/windows.rs:40-150)
# ---------------------------------------------------------------------------------------------
def _window_apply(kind: int, samples, ctx: Context | None):
    lib = A.load()
    if _is_torch(samples):
        import torch
        x = samples.contiguous()
        n = x.shape[-1]
        frames = x.numel() // max(n, 1)
        c = ctx or _torch_ctx(x.device.index or 0, torch.cuda.current_stream(x.device).cuda_stream)
        out = torch.empty_like(x)
        if x.numel():
            _check(lib.sa_window_batched(c._h, kind, x.data_ptr(), frames, n, n, out.data_ptr()))
        return out
    x = np.ascontiguousarray(samples, dtype=np.float32)
    if x.size == 0:
        return x.copy()
    c = ctx or default_context()
    n = x.shape[-1]
    out = np.empty_like(x)
    _check(lib.sa_window_host(c._h, kind, x.ctypes.data, x.size // n, n, n, out.ctypes.data))
    return out


class windows:  # namespace, like the crate's `windows` module
    @staticmethod
    def hann_window(samples, ctx=None): return _window_apply(A.WINDOW_HANN, samples, ctx)

    @staticmethod
    def hamming_window(samples, ctx=None): return _window_apply(A.WINDOW_HAMMING, samples, ctx)

    @staticmethod
    def blackman_harris_4term(samples, ctx=None): return _window_apply(A.WINDOW_BLACKMAN_HARRIS_4TERM, samples, ctx)

    @staticmethod
    def blackman_harris_7term(samples, ctx=None): return _window_apply(A.WINDOW_BLACKMAN_HARRIS_7TERM, samples, ctx)

    @staticmethod
    def coefficients(window, n: int) -> np.ndarray:
        """The per-sample multipliers the reference computes (host, bit-exact f32)."""
        out = np.zeros(n, np.float32)
        _check(A.load().sa_window_coeffs(_window_kind(window), n, out.ctypes.data))
        return out
