"""Pins the CPU oracle (oracle/sa_oracle.c) against every golden vector / known answer the
reference's own tests hold for the hot path, plus an independent f64 DFT.

Reference tests restated here (paths relative to /root/reference):
  src/windows.rs:156-174      window goldens
  src/fft.rs:141-147          N/2+1 bins
  src/scaling.rs:189-209      scale_to_zero_to_one values
  src/spectrum.rs:654-792     9-bin statistics + lookups; :869-923 all-zero; :925-941 no DC; :943-981 test_max
  src/limit.rs:150-192        FrequencyLimit::verify
  src/tests/mod.rs:57-490     sine mix, power, inclusive limits, Nyquist, error taxonomy, N semantics
The FrequencySpectrum getters exercised below are the product's host-side view
(spectrum_analyzer_b200.api.FrequencySpectrum) fed with oracle output, so they run without a GPU.
"""
import numpy as np
import pytest

import oracle as O
import spectrum_analyzer_b200 as sa
from spectrum_analyzer_b200 import cabi


def as_spectrum(res: O.Spectrum, sampling_rate: int, n: int, bin_lo: int = 0) -> sa.FrequencySpectrum:
    st = np.zeros(1, cabi.FRAME_STATS_DTYPE)[0]
    for k in ("min_val", "min_bin", "max_val", "max_bin", "average", "median"):
        st[k] = getattr(res.stats, k)
    return sa.FrequencySpectrum(res.freqs, res.vals, np.float32(sampling_rate) / np.float32(n), n, st, bin_lo)


def spectrum_from_pairs(pairs, resolution):
    fr = np.array([p[0] for p in pairs], np.float32)
    va = np.array([p[1] for p in pairs], np.float32)
    s = O.calc_statistics(va)
    st = np.zeros(1, cabi.FRAME_STATS_DTYPE)[0]
    for k in ("min_val", "min_bin", "max_val", "max_bin", "average", "median"):
        st[k] = getattr(s, k)
    return sa.FrequencySpectrum(fr, va, resolution, len(pairs), st, 0)


def sine_f32(freqs, sr, ms):
    return O.sine_wave_audio_data_multiple(freqs, sr, ms).astype(np.float32)


# ---- src/windows.rs:156-174 -------------------------------------------------------------------
def test_hamming_window_coefficients():
    got = O.window_apply(O.WINDOW_HAMMING, [1.0] * 4)
    assert np.allclose(got, [0.08, 0.77, 0.77, 0.08], atol=1e-5, rtol=0)


def test_blackman_harris_4term_window_coefficients():
    got = O.window_apply(O.WINDOW_BH4, [2.0] * 4)
    assert np.allclose(got, [0.00012, 1.04115, 1.04115, 0.00012], atol=1e-5, rtol=0)


def test_window_passthrough_for_len_le_1():
    for kind in (O.WINDOW_HAMMING, O.WINDOW_BH4, O.WINDOW_BH7):
        assert np.array_equal(O.window_apply(kind, [3.5]), np.array([3.5], np.float32))  # windows.rs:60-63,123-126
    assert O.window_apply(O.WINDOW_HANN, [3.5])[0] == 0.0                                # hann has no passthrough


# ---- src/fft.rs:141-147 -----------------------------------------------------------------------
def test_fft_adapter_length():
    assert O.fft_calc([1.0, 2.0, 3.0, 4.0]).size == 2 + 1


# ---- src/scaling.rs:189-209 -------------------------------------------------------------------
def test_scale_to_zero_to_one():
    data = np.array([0.0, 1.1, 2.2, 3.3, 4.4, 5.5], np.float32)
    got = [O.scale(O.SCALING_ZERO_TO_ONE, v, smin=data[0], smax=data[-1], n=6.0) for v in data]
    assert np.allclose(got, [0.0, 0.2, 0.4, 0.6, 0.8, 1.0], atol=3e-7)
    assert O.scale(O.SCALING_ZERO_TO_ONE, 3.0, smax=0.0) == 0.0
    assert O.scale(O.SCALING_20_TIMES_LOG10, 0.0) == 0.0
    assert O.scale(O.SCALING_DIVIDE_BY_N, 8.0, n=4.0) == 2.0
    assert O.scale(O.SCALING_DIVIDE_BY_N_SQRT, 8.0, n=16.0) == 2.0


# ---- src/spectrum.rs:654-792 ------------------------------------------------------------------
BASIC = [(0.0, 5.0), (50.0, 50.0), (100.0, 100.0), (150.0, 150.0), (200.0, 100.0), (250.0, 20.0),
         (300.0, 0.0), (450.0, 200.0), (500.0, 100.0)]


def test_spectrum_basic_statistics():
    sp = spectrum_from_pairs(BASIC, 50.0)
    assert sp.dc_component() == 5.0
    assert sp.min_fr() == 0.0 and sp.max_fr() == 500.0
    assert sp.min() == (300.0, 0.0)
    assert sp.max() == (450.0, 200.0)
    assert sp.range() == 200.0
    assert sp.average() == np.float32(80.55556)
    assert sp.median() == 100.0
    assert sp.frequency_resolution() == 50.0


def test_spectrum_basic_lookups():
    sp = spectrum_from_pairs(BASIC, 50.0)
    for fr, want in [(0.0, 5.0), (50.0, 50.0), (150.0, 150.0), (200.0, 100.0), (250.0, 20.0), (300.0, 0.0),
                     (375.0, 100.0), (450.0, 200.0)]:
        assert sp.freq_val_exact(fr) == np.float32(want)
    for fr, want in [(0.0, (0.0, 5.0)), (50.0, (50.0, 50.0)), (450.0, (450.0, 200.0)), (448.0, (450.0, 200.0)),
                     (400.0, (450.0, 200.0)), (47.3, (50.0, 50.0)), (51.3, (50.0, 50.0))]:
        assert sp.freq_val_closest(fr) == want


@pytest.mark.parametrize("fn", ["freq_val_exact", "freq_val_closest"])
@pytest.mark.parametrize("fr", [-1.0, 451.0])
def test_spectrum_lookup_out_of_range_panics(fn, fr):
    sp = spectrum_from_pairs([(0.0, 5.0), (450.0, 200.0)], 50.0)
    with pytest.raises(IndexError):
        getattr(sp, fn)(fr)


def test_nan_safety_all_zero():
    sp = spectrum_from_pairs([(0.0, 0.0)] * 8, 50.0)
    for v in (sp.min()[1], sp.max()[1], sp.average(), sp.median()):
        assert np.isfinite(v) and v == 0.0
    s = O.calc_statistics(np.zeros(8, np.float32))
    assert (s.min_bin, s.max_bin) == (0, 7)   # stable sort: min -> first, max -> last among equals


def test_no_dc_component():
    assert spectrum_from_pairs([(150.0, 150.0), (200.0, 100.0)], 50.0).dc_component() is None


def test_max_golden():
    pairs = [(2.6916504, 22.81816), (5.383301, 2.1004658), (8.074951, 8.704016), (10.766602, 3.4043686),
             (13.458252, 8.649045), (16.149902, 9.210494), (18.841553, 14.937911), (21.533203, 5.1524887),
             (24.224854, 20.706167), (26.916504, 8.359295), (29.608154, 3.7514696), (32.299805, 15.109907),
             (34.991455, 86.791145), (37.683105, 52.140736), (40.374756, 24.108875), (43.066406, 11.070151),
             (45.758057, 10.569871), (48.449707, 6.1969466), (51.141357, 16.722788), (53.833008, 8.93011)]
    sp = spectrum_from_pairs(pairs, 44100.0)
    assert sp.max() == (np.float32(34.991455), np.float32(86.791145))


def test_mel_getter():
    sp = spectrum_from_pairs([(0.0, 5.0), (450.0, 200.0)], 50.0)
    assert np.isfinite(sp.mel_val(450.0))


def test_median_even_and_odd():
    assert O.calc_statistics([4, 1, 3, 2]).median == 2.5          # (s[1]+s[2])/2
    assert O.calc_statistics([5, 1, 3]).median == 3.0             # odd-length fix (CHANGELOG 1.8.0)
    assert O.calc_statistics([1, 1, 1, 9]).median == 1.0


# ---- src/limit.rs:150-192 ---------------------------------------------------------------------
def test_limit_verify():
    assert O.limit_verify(O.LIMIT_MIN, -1.0, 0, 0.0) == O.ERR_LIMIT_BELOW_MINIMUM
    assert O.limit_verify(O.LIMIT_MIN, 1.0, 0, 0.0) == O.ERR_LIMIT_ABOVE_NYQUIST
    assert O.limit_verify(O.LIMIT_MAX, 0, -1.0, 0.0) == O.ERR_LIMIT_BELOW_MINIMUM
    assert O.limit_verify(O.LIMIT_MAX, 0, 1.0, 0.0) == O.ERR_LIMIT_ABOVE_NYQUIST
    assert O.limit_verify(O.LIMIT_RANGE, -1.0, 0.0, 0.0) == O.ERR_LIMIT_BELOW_MINIMUM
    assert O.limit_verify(O.LIMIT_RANGE, 0.0, 1.0, 0.0) == O.ERR_LIMIT_ABOVE_NYQUIST
    assert O.limit_verify(O.LIMIT_RANGE, 0.0, -1.0, 0.0) != O.OK
    assert O.limit_verify(O.LIMIT_RANGE, 60.0, 50.0, 100.0) == O.ERR_LIMIT_INVALID_RANGE
    for args in [(O.LIMIT_MIN, 50.0, 0), (O.LIMIT_MAX, 0, 50.0), (O.LIMIT_RANGE, 50.0, 50.0), (O.LIMIT_RANGE, 50.0, 70.0)]:
        assert O.limit_verify(*args, 100.0) == O.OK


# ---- src/tests/mod.rs ---------------------------------------------------------------------------
def test_sine_mix_50_1000_3777hz():                                # mod.rs:57-144
    audio = sine_f32([50.0, 1000.0, 3777.0], 44100, 1000)
    assert audio.size == 44100
    window = audio[:4096]
    spectra = {}
    for name, kind in [("no-window", O.WINDOW_NONE), ("hamming", O.WINDOW_HAMMING), ("hann", O.WINDOW_HANN),
                       ("bh4", O.WINDOW_BH4), ("bh7", O.WINDOW_BH7)]:
        r = O.samples_fft_to_spectrum(O.window_apply(kind, window), 44100, O.LIMIT_MAX, 0, 4000.0, O.SCALING_ZERO_TO_ONE)
        assert r.code == O.OK
        spectra[name] = as_spectrum(r, 44100, 4096)
    hann = spectra["hann"]
    for fr in (50.0, 1000.0, 3777.0):
        assert hann.freq_val_exact(fr) > 0.85
        assert hann.freq_val_closest(fr)[1] > 0.85
    assert hann.freq_val_exact(500.0) < 0.00001
    assert hann.freq_val_closest(500.0)[1] < 0.00001
    # SURVEY.md 8c known answers (numpy f64 FFT of the same fixture)
    assert len(hann.values()) == 372 and int(np.argmax(hann.values())) == 93
    assert abs(hann.values()[351] - 0.985) < 2e-3 and abs(hann.values()[5] - 0.930) < 2e-3


def test_spectrum_power():                                         # mod.rs:150-222
    audio = sine_f32([2048.0], 44100, 1000)
    a = as_spectrum(O.samples_fft_to_spectrum(audio[:2048], 44100, O.LIMIT_MAX, 0, 4000.0, O.SCALING_DIVIDE_BY_N), 44100, 2048)
    b = as_spectrum(O.samples_fft_to_spectrum(audio[:4096], 44100, O.LIMIT_MAX, 0, 4000.0, O.SCALING_DIVIDE_BY_N), 44100, 4096)
    va, vb = a.freq_val_exact(2048.0), b.freq_val_exact(2048.0)
    assert abs(va - vb) / max(va, vb) < 0.122


def test_frequency_limit_inclusive():                              # mod.rs:224-290
    sr = 1024
    window = O.window_apply(O.WINDOW_HANN, sine_f32([512.0], sr, 1000)[:512])
    r = O.samples_fft_to_spectrum(window, sr, O.LIMIT_MAX, 0, 400.0)
    assert (r.freqs[0], r.freqs[-1]) == (0.0, 400.0)
    r = O.samples_fft_to_spectrum(window, sr, O.LIMIT_MIN, 100.0, 0)
    assert (r.freqs[0], r.freqs[-1]) == (100.0, sr / 2.0)
    r = O.samples_fft_to_spectrum(window, sr, O.LIMIT_RANGE, 412.0, 510.0)
    assert (r.freqs[0], r.freqs[-1]) == (412.0, 510.0)


def test_nyquist_theorem():                                        # mod.rs:293-321
    r = O.samples_fft_to_spectrum(np.zeros(4096, np.float32), 44100)
    assert r.code == O.OK and r.vals.size == 4096 // 2 + 1 and (r.vals == 0.0).all()
    assert r.freqs[0] == 0.0 and r.freqs[-1] == 44100.0 / 2.0


def test_nyquist_theorem2():                                       # mod.rs:324-379
    audio = sine_f32([22049.9], 44100, 1000)
    sp = as_spectrum(O.samples_fft_to_spectrum(audio[:4096], 44100, O.LIMIT_ALL, 0, 0, O.SCALING_ZERO_TO_ONE), 44100, 4096)
    assert sp.min_fr() == 0.0 and sp.max_fr() == 22050.0
    assert sp.max()[1] > 0.99
    assert sp.freq_val_exact(22049.9) >= 0.94
    assert sp.freq_val_exact(22049.0) >= 0.49
    assert sp.freq_val_exact(22035.0) <= 0.26
    assert sp.freq_val_exact(22000.0) <= 0.07
    assert sp.freq_val_exact(21500.0) <= 0.01


def test_invalid_input():                                          # mod.rs:381-432
    f = O.samples_fft_to_spectrum
    assert f([0, 1, 2, 3, np.nan, 4, 5, 6], 44100).code == O.ERR_NAN_VALUES
    assert f([0, 1, 2, 3, np.inf, 4, 5, 6], 44100).code == O.ERR_INFINITY_VALUES
    assert f([0.0], 44100).code == O.ERR_TOO_FEW_SAMPLES
    assert f([0.0] * 4, 44100, O.LIMIT_MIN, -1.0).code == O.ERR_LIMIT_BELOW_MINIMUM
    assert f([0.0] * 8, 8, O.LIMIT_RANGE, 1.1, 1.9).code == O.ERR_LIMIT_TOO_NARROW
    assert f([0.0] * 8, 8, O.LIMIT_RANGE, 1.0, 1.0).code == O.ERR_LIMIT_TOO_NARROW
    assert f([0.0] * 3, 44100).code == O.ERR_NOT_POWER_OF_TWO
    # precedence (src/lib.rs:164-181): NaN before Inf before power-of-two before limit
    assert f([np.inf, np.nan, 0.0], 44100, O.LIMIT_MIN, -1.0).code == O.ERR_NAN_VALUES
    assert f([np.inf, 0.0, 0.0], 44100, O.LIMIT_MIN, -1.0).code == O.ERR_INFINITY_VALUES
    assert f([0.0, 0.0, 0.0], 44100, O.LIMIT_MIN, -1.0).code == O.ERR_NOT_POWER_OF_TWO
    assert f(np.zeros(65536, np.float32), 44100).code == O.ERR_UNSUPPORTED_LENGTH      # src/fft.rs:52 panic


def test_only_null_samples_valid():                                # mod.rs:434-438
    assert O.samples_fft_to_spectrum([0.0, 0.0], 44100).code == O.OK


def test_divide_by_n_has_effect():                                 # mod.rs:453-490
    audio = O.window_apply(O.WINDOW_HANN, sine_f32([100.0, 200.0, 400.0], 1000, 2000)[:1024])
    normal = O.samples_fft_to_spectrum(audio, 1000)
    scaled = O.samples_fft_to_spectrum(audio, 1000, scaling_kind=O.SCALING_DIVIDE_BY_N)
    assert ((normal.vals[:512] / 1024.0 - scaled.vals[:512]) < 0.1).all()
    limited = O.samples_fft_to_spectrum(audio, 1000, O.LIMIT_MAX, 0, 250.0, O.SCALING_DIVIDE_BY_N)
    assert np.array_equal(scaled.vals[:256], limited.vals[:256])   # stats.n is samples_len, not the bin count


# ---- independent truth: f64 DFT ------------------------------------------------------------------
@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768])
def test_oracle_fft_against_f64_dft(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(np.float32)
    got = O.fft_calc(x)
    ref = np.fft.rfft(x.astype(np.float64))
    assert got.size == n // 2 + 1
    assert got[0].imag == 0.0 and got[-1].imag == 0.0            # Nyquist unpacked (src/fft.rs:126-132)
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()


def test_oracle_libm_restatements_within_one_ulp():
    xs = np.concatenate([np.linspace(0, 40, 20001), np.linspace(-40, 0, 5001)]).astype(np.float32)
    mine = np.array([O.cosf(v) for v in xs], np.float32)
    ref = np.cos(xs.astype(np.float64)).astype(np.float32)
    assert np.abs(mine.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64)).max() <= 1
    xs = np.concatenate([np.exp(np.linspace(-100, 88, 20001)), [1.0, 0.5, 2.0, 1e-45, 1e-40]]).astype(np.float32)
    mine = np.array([O.log10f(v) for v in xs], np.float32)
    ref = np.log10(xs.astype(np.float64)).astype(np.float32)
    assert np.abs(mine.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64)).max() <= 1
    assert O.log10f(1.0) == 0.0 and O.log10f(100.0) == 2.0


def test_config_bin_counts_from_survey():
    # SURVEY.md section 8 config sizes
    assert O.limit_to_bins(2048, 44100, O.LIMIT_MAX, 0, 4000.0) == (0, 186)
    assert O.limit_to_bins(4096, 44100, O.LIMIT_MAX, 0, 4000.0) == (0, 372)
    assert O.limit_to_bins(4096, 44100, O.LIMIT_ALL) == (0, 2049)
    assert O.limit_to_bins(16384, 44100, O.LIMIT_RANGE, 50.0, 15000.0) == (19, 5554)
    assert O.limit_to_bins(256, 44100, O.LIMIT_ALL) == (0, 129)
    assert np.float32(O.frequency_resolution(44100, 4096)) == np.float32(10.766602)


def test_batched_driver_equals_single_calls():
    rng = np.random.default_rng(1)
    n, frames, hop = 256, 7, 100
    x = rng.standard_normal((frames - 1) * hop + n).astype(np.float32)
    code, vals, stats, status, lo = O.spectrum_batched(x, frames, n, hop, 8000, O.LIMIT_RANGE, 500.0, 3000.0,
                                                       O.WINDOW_BH7, O.SCALING_20_TIMES_LOG10, threads=3)
    assert code == 0 and (status == 0).all()
    for f in range(frames):
        w = O.window_apply(O.WINDOW_BH7, x[f * hop: f * hop + n])
        r = O.samples_fft_to_spectrum(w, 8000, O.LIMIT_RANGE, 500.0, 3000.0, O.SCALING_20_TIMES_LOG10)
        assert np.array_equal(r.vals, vals[f]) and r.stats.median == stats["median"][f]
        assert r.stats.max_bin == stats["max_bin"][f]
