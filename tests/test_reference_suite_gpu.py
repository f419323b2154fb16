"""The reference's own integration tests (src/tests/mod.rs) run against the CUDA product through the
drop-in API (`samples_fft_to_spectrum` -> sa_spectrum_single -> fused kernel).  Same fixtures, same
assertions; thresholds are the reference's."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sa(gpu_ctx):
    import spectrum_analyzer_b200 as m
    return m


def sine_f32(freqs, sr, ms):
    return O.sine_wave_audio_data_multiple(freqs, sr, ms).astype(np.float32)   # restated src/tests/sine.rs fixture


def test_spectrum_and_visualize_sine_waves_50_1000_3777hz(sa):              # mod.rs:57-144
    window = sine_f32([50.0, 1000.0, 3777.0], 44100, 1000)[:4096]
    windows = {"no-window": window, "hamming": sa.windows.hamming_window(window), "hann": sa.windows.hann_window(window),
               "bh4": sa.windows.blackman_harris_4term(window), "bh7": sa.windows.blackman_harris_7term(window)}
    spectra = {k: sa.samples_fft_to_spectrum(v, 44100, sa.FrequencyLimit.Max(4000.0), sa.scaling.scale_to_zero_to_one)
               for k, v in windows.items()}
    hann = spectra["hann"]
    for fr in (50.0, 1000.0, 3777.0):
        assert hann.freq_val_exact(fr) > 0.85
        assert hann.freq_val_closest(fr)[1] > 0.85
    assert hann.freq_val_exact(500.0) < 0.00001
    assert hann.freq_val_closest(500.0)[1] < 0.00001
    # fused window vs caller-side window: the tuned kernels multiply by the window with a packed multiply that ptxas
    # contracts into the first butterfly additions (FFMA2), so half of the windowed samples are never rounded on their
    # own; the two spectra agree to a few ulp of the frame maximum (1.0 here), not bit for bit
    fused = sa.samples_fft_to_spectrum(window, 44100, sa.FrequencyLimit.Max(4000.0), sa.scaling.scale_to_zero_to_one,
                                       window="hann")
    assert np.abs(fused.values() - hann.values()).max() <= 4 * np.finfo(np.float32).eps
    # against the oracle, per bin
    for name, kind in [("no-window", O.WINDOW_NONE), ("hamming", O.WINDOW_HAMMING), ("hann", O.WINDOW_HANN),
                       ("bh4", O.WINDOW_BH4), ("bh7", O.WINDOW_BH7)]:
        ref = O.samples_fft_to_spectrum(O.window_apply(kind, window), 44100, O.LIMIT_MAX, 0, 4000.0, O.SCALING_ZERO_TO_ONE)
        got = spectra[name]
        assert np.array_equal(got.frequencies(), ref.freqs)
        assert np.abs(got.values() - ref.vals).max() <= 2.5e-4
        assert got.max()[0] == ref.freqs[ref.stats.max_bin] and got.max()[1] == 1.0


def test_spectrum_power(sa):                                                # mod.rs:150-222
    audio = sine_f32([2048.0], 44100, 1000)
    a = sa.samples_fft_to_spectrum(audio[:2048], 44100, sa.FrequencyLimit.Max(4000.0), sa.scaling.divide_by_N)
    b = sa.samples_fft_to_spectrum(audio[:4096], 44100, sa.FrequencyLimit.Max(4000.0), sa.scaling.divide_by_N)
    va, vb = a.freq_val_exact(2048.0), b.freq_val_exact(2048.0)
    assert abs(va - vb) / max(va, vb) < 0.122


def test_spectrum_frequency_limit_inclusive(sa):                            # mod.rs:224-290
    sr = 1024
    window = sa.windows.hann_window(sine_f32([512.0], sr, 1000)[:512])
    s = sa.samples_fft_to_spectrum(window, sr, sa.FrequencyLimit.Max(400.0), None)
    assert (s.min_fr(), s.max_fr()) == (0.0, 400.0)
    s = sa.samples_fft_to_spectrum(window, sr, sa.FrequencyLimit.Min(100.0), None)
    assert (s.min_fr(), s.max_fr()) == (100.0, sr / 2.0)
    s = sa.samples_fft_to_spectrum(window, sr, sa.FrequencyLimit.Range(412.0, 510.0), None)
    assert (s.min_fr(), s.max_fr()) == (412.0, 510.0)


def test_spectrum_nyquist_theorem(sa):                                      # mod.rs:293-321
    s = sa.samples_fft_to_spectrum(np.zeros(4096, np.float32), 44100, sa.FrequencyLimit.All, None)
    assert sum(1 for _, v in s.data() if v == 0.0) == 4096 // 2 + 1
    assert s.min_fr() == 0.0 and s.max_fr() == 44100.0 / 2.0


def test_spectrum_nyquist_theorem2(sa):                                     # mod.rs:324-379
    audio = sine_f32([22049.9], 44100, 1000)
    s = sa.samples_fft_to_spectrum(audio[:4096], 44100, sa.FrequencyLimit.All, sa.scaling.scale_to_zero_to_one)
    assert s.min_fr() == 0.0 and s.max_fr() == 22050.0
    assert s.max()[1] > 0.99
    assert s.freq_val_exact(22049.9) >= 0.94
    assert s.freq_val_exact(22049.0) >= 0.49
    assert s.freq_val_exact(22035.0) <= 0.26
    assert s.freq_val_exact(22000.0) <= 0.07
    assert s.freq_val_exact(21500.0) <= 0.01


def test_invalid_input(sa):                                                 # mod.rs:381-432
    f = sa.samples_fft_to_spectrum
    with pytest.raises(sa.NaNValuesNotSupported):
        f([0.0, 1.0, 2.0, 3.0, np.nan, 4.0, 5.0, 6.0], 44100, sa.FrequencyLimit.All, None)
    with pytest.raises(sa.InfinityValuesNotSupported):
        f([0.0, 1.0, 2.0, 3.0, np.inf, 4.0, 5.0, 6.0], 44100, sa.FrequencyLimit.All, None)
    with pytest.raises(sa.TooFewSamples):
        f([0.0], 44100, sa.FrequencyLimit.All, None)
    with pytest.raises(sa.InvalidFrequencyLimit):
        f([0.0] * 4, 44100, sa.FrequencyLimit.Min(-1.0), None)
    with pytest.raises(sa.FrequencyLimitTooNarrow):
        f([0.0] * 8, 8, sa.FrequencyLimit.Range(1.1, 1.9), None)
    with pytest.raises(sa.FrequencyLimitTooNarrow):
        f([0.0] * 8, 8, sa.FrequencyLimit.Range(1.0, 1.0), None)
    with pytest.raises(sa.SamplesLengthNotAPowerOfTwo):
        f([0.0] * 3, 44100, sa.FrequencyLimit.All, None)
    with pytest.raises(sa.NaNValuesNotSupported):       # precedence: NaN before power-of-two (src/lib.rs:168-176)
        f([0.0, np.nan, 1.0], 44100, sa.FrequencyLimit.All, None)
    with pytest.raises(sa.UnsupportedLength):           # the reference panics (src/fft.rs:52)
        f(np.zeros(65536, np.float32), 44100, sa.FrequencyLimit.All, None)


def test_only_null_samples_valid(sa):                                       # mod.rs:434-438
    s = sa.samples_fft_to_spectrum([0.0, 0.0], 44100, sa.FrequencyLimit.All, None)
    assert len(s.data()) == 2 and s.max()[1] == 0.0


def test_scaling_closure_is_rejected(sa):                                   # mod.rs:440-450 (closures: rejected, not run)
    with pytest.raises(TypeError):
        sa.samples_fft_to_spectrum([1.1, 2.2, 3.3, 4.4, 5.5, 6.6, 7.7, 8.8], 44100, sa.FrequencyLimit.All,
                                   lambda _v, _i: float("nan"))


def test_divide_by_n_has_effect(sa):                                        # mod.rs:453-490
    audio = sa.windows.hann_window(sine_f32([100.0, 200.0, 400.0], 1000, 2000)[:1024])
    normal = sa.samples_fft_to_spectrum(audio, 1000, sa.FrequencyLimit.All, None)
    scaled = sa.samples_fft_to_spectrum(audio, 1000, sa.FrequencyLimit.All, sa.scaling.divide_by_N)
    assert ((normal.values()[:512] / 1024.0 - scaled.values()[:512]) < 0.1).all()
    limited = sa.samples_fft_to_spectrum(audio, 1000, sa.FrequencyLimit.Max(250.0), sa.scaling.divide_by_N)
    assert np.array_equal(scaled.values()[:256], limited.values()[:256])     # stats.n = samples_len (exact)


def test_minimal_example_and_doc_tests(sa):                                 # examples/minimal.rs:30-52, lib.rs:130-153
    samples = [0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0]
    s = sa.samples_fft_to_spectrum(sa.windows.hann_window(samples), 44100, sa.FrequencyLimit.All, sa.scaling.divide_by_N)
    ref = O.samples_fft_to_spectrum(O.window_apply(O.WINDOW_HANN, samples), 44100, O.LIMIT_ALL, 0, 0, O.SCALING_DIVIDE_BY_N)
    assert len(s.data()) == 5 and np.abs(s.values() - ref.vals).max() <= 1e-4 * ref.vals.max()
    s = sa.samples_fft_to_spectrum([0.0, 1.1, 5.5, -5.5], 44100, sa.FrequencyLimit.All, sa.scaling.scale_to_zero_to_one)
    assert len(s.data()) == 3 and s.max()[1] == 1.0


def test_benchmark_shapes(sa):                                              # benches/fft_spectrum_bench.rs:38-55
    rng = np.random.default_rng(0)
    samples = rng.integers(-32768, 32767, 2048).astype(np.float32)
    hann = sa.windows.hann_window(samples)
    assert np.array_equal(hann, O.window_apply(O.WINDOW_HANN, samples))
    for sc_gpu, sc_cpu in [(None, O.SCALING_NONE), (sa.scaling.divide_by_N_sqrt, O.SCALING_DIVIDE_BY_N_SQRT)]:
        s = sa.samples_fft_to_spectrum(hann, 44100, sa.FrequencyLimit.All, sc_gpu)
        ref = O.samples_fft_to_spectrum(hann, 44100, O.LIMIT_ALL, 0, 0, sc_cpu)
        assert np.abs(s.values() - ref.vals).max() <= 1e-4 * ref.vals.max()
        assert abs(s.median() - ref.stats.median) <= 1e-4 * ref.vals.max()
        assert s.max()[0] == ref.freqs[ref.stats.max_bin]
