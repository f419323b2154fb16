"""CPU-side checks of the product: the C-ABI library loads and exports every symbol include/*.h
declares, its host-only entry points (window multipliers, FrequencyLimit -> bin range, bin
frequencies, error taxonomy) agree bit for bit with the oracle, compute entry points fail LOUDLY
without a GPU (no CPU fallback), and the Python mirror rejects closures."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle as O
import spectrum_analyzer_b200 as sa
from spectrum_analyzer_b200 import cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "spectrum_analyzer_cuda.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sa_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(cuda_lib):
    names = header_functions()
    assert len(names) >= 17
    for n in names:
        assert hasattr(cuda_lib, n), f"{n} declared in the header but not exported"
        assert n in cabi.SYMBOLS, f"{n} has no ctypes prototype"
    assert sorted(cabi.SYMBOLS) == names
    assert cuda_lib.sa_abi_version() == 1


def test_header_enums_match_python_and_oracle():
    src = open(HEADER).read()
    for name, val in re.findall(r"\b(SA_[A-Z0-9_]+)\s*=\s*(\d+)", src):
        py = name[3:]
        py = py.replace("BLACKMAN_HARRIS_4TERM", "BLACKMAN_HARRIS_4TERM")
        if hasattr(cabi, py):
            assert getattr(cabi, py) == int(val), name
    assert (cabi.WINDOW_HANN, cabi.WINDOW_BLACKMAN_HARRIS_7TERM) == (O.WINDOW_HANN, O.WINDOW_BH7)
    assert (cabi.SCALING_20_TIMES_LOG10, cabi.LIMIT_RANGE) == (O.SCALING_20_TIMES_LOG10, O.LIMIT_RANGE)
    for k in ("ERR_TOO_FEW_SAMPLES", "ERR_NAN_VALUES", "ERR_INFINITY_VALUES", "ERR_NOT_POWER_OF_TWO",
              "ERR_LIMIT_BELOW_MINIMUM", "ERR_LIMIT_ABOVE_NYQUIST", "ERR_LIMIT_INVALID_RANGE", "ERR_LIMIT_TOO_NARROW",
              "ERR_SCALING", "ERR_UNSUPPORTED_LENGTH", "ERR_NON_FINITE_RESULT"):
        assert getattr(cabi, k) == getattr(O, k)
    assert C.sizeof(cabi.FrameStats) == 32


@pytest.mark.parametrize("kind", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8, 100, 256, 2048, 4096, 16384, 32768])
def test_window_multipliers_bit_exact(cuda_lib, kind, n):
    got = np.zeros(n, np.float32)
    assert cuda_lib.sa_window_coeffs(kind, n, got.ctypes.data) == 0
    want = O.window_coeffs(kind, n)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def _f32_neighbours(c):
    """c - 1 ulp, c, c + 1 ulp as float32 arrays."""
    c = np.asarray(c, np.float32)
    return np.nextafter(c, np.float32(-np.inf)), c, np.nextafter(c, np.float32(np.inf))


@pytest.mark.parametrize("kind", [1, 2])
@pytest.mark.parametrize("n", [2, 3, 8, 100, 256, 1024, 2048, 4096, 8192, 16384, 32768])
def test_window_multipliers_against_an_independent_cos(cuda_lib, kind, n):
    """The product's `cosf` restatement and the oracle's are two readings of the same musl source by the same author;
    comparing them with each other cannot catch a shared misreading.  Independent witness: libm's cosf is documented to
    be within 1 ulp, so for the exact f32 argument the reference forms (src/windows.rs:44-46, 66) the multiplier must
    be one of the three values obtained from the correctly rounded cosine (numpy float64 cos, rounded once) and its two
    f32 neighbours -- and almost always the middle one."""
    got = np.zeros(n, np.float32)
    assert cuda_lib.sa_window_coeffs(kind, n, got.ctypes.data) == 0
    f = np.float32
    i = np.arange(n, dtype=np.float32)
    two_pi = f(2.0) * f(np.pi)                                   # 2.0 * PI, one f32 product
    if kind == 1:
        arg = (two_pi * i).astype(np.float32) / f(n)             # (2.0 * PI * i as f32) / samples_len_f32
        mult = lambda c: (f(0.5) * (f(1.0) - c)).astype(np.float32)
    else:
        arg = (two_pi * i).astype(np.float32) / (f(n) - f(1.0))
        mult = lambda c: (f(0.54) - (f(0.46) * c).astype(np.float32)).astype(np.float32)
    arg = arg.astype(np.float32)
    lo, mid, hi = _f32_neighbours(np.cos(arg.astype(np.float64)).astype(np.float32))
    cands = [mult(c).view(np.uint32) for c in (lo, mid, hi)]
    g = got.view(np.uint32)
    ok = (g == cands[0]) | (g == cands[1]) | (g == cands[2])
    assert ok.all(), f"multiplier {np.nonzero(~ok)[0][:5]} outside the 1-ulp envelope of a correctly rounded cos"
    assert (g == cands[1]).mean() >= 0.98                        # musl's cosf is correctly rounded almost everywhere


@pytest.mark.parametrize("kind,alphas", [(3, [0.35875, -0.48829, 0.14128, -0.01168]),
                                         (4, [0.2710514, -0.43329793, 0.218123, -0.06592545, 0.010811742,
                                              -0.00077658484, 0.000013887217])])
@pytest.mark.parametrize("n", [4, 100, 4096, 16384])
def test_blackman_harris_multipliers_against_float64(cuda_lib, kind, alphas, n):
    """Same witness for the multi-term windows (src/windows.rs:122-147): with every cosine within 1 ulp and one f32
    rounding per multiply-add, the sum lies within sum(|alpha|) * (terms + 2) ulp(1) of the float64 value formed from
    the same f32 arguments."""
    got = np.zeros(n, np.float32)
    assert cuda_lib.sa_window_coeffs(kind, n, got.ctypes.data) == 0
    f = np.float32
    i = np.arange(n, dtype=np.float32)
    acc = np.zeros(n, np.float64)
    for a, alpha in enumerate(alphas):
        two_pi_iteration = f(2.0) * f(a) * f(np.pi)
        arg = ((two_pi_iteration * i).astype(np.float32) / (f(n) - f(1.0))).astype(np.float32)
        acc += float(f(alpha)) * np.cos(arg.astype(np.float64))
    bound = sum(abs(a) for a in alphas) * (len(alphas) + 2) * 2.0 ** -23
    assert np.abs(got.astype(np.float64) - acc).max() <= bound


def test_window_goldens_through_the_abi(cuda_lib):
    assert np.allclose(sa.windows.coefficients("hamming", 4), [0.08, 0.77, 0.77, 0.08], atol=1e-5)          # windows.rs:156-164
    assert np.allclose(2 * sa.windows.coefficients("blackman_harris_4term", 4), [0.00012, 1.04115, 1.04115, 0.00012],
                       atol=1e-5)                                                                            # windows.rs:166-174


def test_limit_to_bins_bit_exact_against_oracle(cuda_lib):
    rng = np.random.default_rng(0)
    lo, nb = C.c_uint32(), C.c_uint32()
    cases = 0
    for n in (2, 4, 8, 64, 512, 2048, 4096, 16384, 32768):
        for sr in (8, 1000, 1024, 8000, 44100, 48000, 96000, 192000):
            nyq = sr / 2
            for _ in range(12):
                kind = int(rng.integers(0, 4))
                a, b = sorted(rng.uniform(0, nyq, 2).astype(np.float32))
                if rng.random() < 0.3:   # hit exact bin frequencies too
                    res = np.float32(sr) / np.float32(n)
                    a = np.float32(rng.integers(0, n // 2 + 1)) * res
                    b = max(a, np.float32(rng.integers(0, n // 2 + 1)) * res)
                    a, b = np.float32(min(a, nyq)), np.float32(min(b, nyq))
                code = cuda_lib.sa_limit_to_bins(n, sr, kind, float(a), float(b), C.byref(lo), C.byref(nb))
                o_lo, o_nb = O.limit_to_bins(n, sr, kind, float(a), float(b))
                want_code = O.limit_verify(kind, float(a), float(b), float(np.float32(sr) / np.float32(2.0)))
                if want_code == 0 and o_nb < 2:
                    want_code = O.ERR_LIMIT_TOO_NARROW
                assert code == want_code, (n, sr, kind, a, b)
                if want_code in (0, O.ERR_LIMIT_TOO_NARROW):
                    assert (lo.value, nb.value) == (o_lo, o_nb), (n, sr, kind, a, b)
                cases += 1
    assert cases > 800


def test_limit_errors_and_precedence(cuda_lib):
    f = cuda_lib.sa_limit_to_bins
    z = C.c_uint32()
    assert f(1, 44100, cabi.LIMIT_ALL, 0, 0, C.byref(z), C.byref(z)) == cabi.ERR_TOO_FEW_SAMPLES
    assert f(3, 44100, cabi.LIMIT_ALL, 0, 0, C.byref(z), C.byref(z)) == cabi.ERR_NOT_POWER_OF_TWO
    assert f(3, 44100, cabi.LIMIT_MIN, -1.0, 0, C.byref(z), C.byref(z)) == cabi.ERR_NOT_POWER_OF_TWO   # pow2 before limit
    assert f(4, 44100, cabi.LIMIT_MIN, -1.0, 0, C.byref(z), C.byref(z)) == cabi.ERR_LIMIT_BELOW_MINIMUM
    assert f(4, 44100, cabi.LIMIT_MAX, 0, 22051.0, C.byref(z), C.byref(z)) == cabi.ERR_LIMIT_ABOVE_NYQUIST
    assert f(4, 44100, cabi.LIMIT_RANGE, 60.0, 50.0, C.byref(z), C.byref(z)) == cabi.ERR_LIMIT_INVALID_RANGE
    assert f(8, 8, cabi.LIMIT_RANGE, 1.1, 1.9, C.byref(z), C.byref(z)) == cabi.ERR_LIMIT_TOO_NARROW        # mod.rs:410-423
    assert f(8, 8, cabi.LIMIT_RANGE, 1.0, 1.0, C.byref(z), C.byref(z)) == cabi.ERR_LIMIT_TOO_NARROW
    assert f(65536, 44100, cabi.LIMIT_ALL, 0, 0, C.byref(z), C.byref(z)) == cabi.ERR_UNSUPPORTED_LENGTH    # fft.rs:52
    assert f(65536, 44100, cabi.LIMIT_MIN, -1.0, 0, C.byref(z), C.byref(z)) == cabi.ERR_LIMIT_BELOW_MINIMUM  # limit first
    assert f(4096, 44100, 9, 0, 0, C.byref(z), C.byref(z)) == cabi.ERR_INVALID_ARGUMENT


def test_bin_frequencies_bit_exact(cuda_lib):
    for n, sr in [(4096, 44100), (512, 1024), (16384, 44100), (2, 8), (32768, 192000)]:
        got = np.zeros(n // 2 + 1, np.float32)
        assert cuda_lib.sa_bin_frequencies(n, sr, 0, n // 2 + 1, got.ctypes.data) == 0
        want = np.arange(n // 2 + 1, dtype=np.float32) * (np.float32(sr) / np.float32(n))
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        r = O.samples_fft_to_spectrum(np.zeros(n, np.float32), sr)
        assert np.array_equal(got, r.freqs)
    got = np.zeros(5554, np.float32)
    cuda_lib.sa_bin_frequencies(16384, 44100, 19, 5554, got.ctypes.data)
    assert got[0] == np.float32(51.141357) and got[-1] == np.float32(14997.876)


def test_status_strings_follow_the_reference_display(cuda_lib):
    assert sa.cabi.status_string(cabi.ERR_TOO_FEW_SAMPLES) == "Too few samples!"                  # error.rs:59
    assert sa.cabi.status_string(cabi.ERR_NAN_VALUES) == "NaN values are not supported!"
    assert sa.cabi.status_string(cabi.ERR_NOT_POWER_OF_TWO) == "Samples length must be a power of two!"
    assert "too few frequency bins" in sa.cabi.status_string(cabi.ERR_LIMIT_TOO_NARROW)


def test_python_mirror_rejects_closures_and_maps_errors():
    with pytest.raises(TypeError):                                        # arbitrary closures cannot run on the GPU
        sa.samples_fft_to_spectrum(np.zeros(8, np.float32), 44100, sa.FrequencyLimit.All, lambda v, s: v - s.min)
    with pytest.raises(TypeError):
        sa.samples_fft_to_spectrum_batched(np.zeros((2, 8), np.float32), 44100, scaling_fn=np.sqrt)
    with pytest.raises(sa.InvalidFrequencyLimit):
        sa.FrequencyLimit.Min(-1.0).to_bins(4, 44100)
    with pytest.raises(sa.FrequencyLimitTooNarrow):
        sa.FrequencyLimit.Range(1.1, 1.9).to_bins(8, 8)
    with pytest.raises(sa.SamplesLengthNotAPowerOfTwo):
        sa.FrequencyLimit.All.to_bins(3, 44100)
    assert sa.FrequencyLimit.Max(4000.0).to_bins(2048, 44100) == (0, 186)          # cfg1
    assert sa.FrequencyLimit.Range(50.0, 15000.0).maybe_min() == 50.0
    assert sa.FrequencyLimit.All.maybe_max() is None


def test_compute_fails_loudly_without_a_gpu(cuda_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert cuda_lib.sa_ctx_create(0, C.byref(h)) == cabi.ERR_CUDA          # no device -> error code, no fallback
    assert cuda_lib.sa_last_error_string().decode() != ""
    with pytest.raises(sa.CudaError):
        sa.samples_fft_to_spectrum(np.zeros(8, np.float32), 44100)
    x = np.zeros(8, np.float32)
    assert cuda_lib.sa_spectrum_batched(None, x.ctypes.data, 1, 8, 8, 44100, 0, 0.0, 0.0, 0, 0, 7, x.ctypes.data, 5,
                                        None, None, None) == cabi.ERR_INVALID_ARGUMENT


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "spectrum_analyzer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "sa_oracle" not in text and "from oracle" not in text, f
