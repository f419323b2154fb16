"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Parity bar (BASELINE.json north_star):
  * bit-exact: bin frequencies, FrequencyLimit index range, median rank selection, argmin/argmax
    bins, per-frame status codes, window multipliers, and -- given the GPU's own unscaled
    magnitudes -- every scaling function and every order statistic;
  * per-bin magnitudes within 1e-4 x the frame's maximum magnitude of the f32 oracle
    (MAG_RTOL below); the average within the same bound (it is summed in a different order).
"""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

MAG_RTOL = 1e-4  # x frame max, per north_star

ALL_N = [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768]
WINDOWS = [O.WINDOW_NONE, O.WINDOW_HANN, O.WINDOW_HAMMING, O.WINDOW_BH4, O.WINDOW_BH7]
SCALINGS = [O.SCALING_NONE, O.SCALING_DIVIDE_BY_N, O.SCALING_DIVIDE_BY_N_SQRT, O.SCALING_ZERO_TO_ONE,
            O.SCALING_20_TIMES_LOG10]


def sa():
    import spectrum_analyzer_b200 as m
    return m


def run_gpu(x, n_frames, n, stride, sr, limit=(O.LIMIT_ALL, 0.0, 0.0), window=O.WINDOW_NONE,
            scaling=O.SCALING_NONE, stats=7, device=True):
    """Returns (vals[n_frames, n_bins], stats recarray, bin_lo) from the CUDA library."""
    import torch
    m = sa()
    lim = m.FrequencyLimit(limit[0], float(limit[1]), float(limit[2]))
    sc = {0: None, 1: m.scaling.divide_by_N, 2: m.scaling.divide_by_N_sqrt, 3: m.scaling.scale_to_zero_to_one,
          4: m.scaling.scale_20_times_log10}[scaling]
    if device:
        xd = torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda()
        r = m.samples_fft_to_spectrum_batched(xd, sr, lim, sc, window=window, n=n, frame_stride=stride,
                                              n_frames=n_frames, stats=stats)
        torch.cuda.synchronize()
        return r.vals.cpu().numpy()[:, : r.n_bins], r.stats_numpy(), r.bin_lo
    r = m.samples_fft_to_spectrum_batched(np.ascontiguousarray(x, np.float32), sr, lim, sc, window=window, n=n,
                                          frame_stride=stride, n_frames=n_frames, stats=stats)
    return r.vals[:, : r.n_bins], r.stats, r.bin_lo


def check_against_oracle(x, n_frames, n, stride, sr, limit, window, scaling, device=True):
    gv, gs, gl = run_gpu(x, n_frames, n, stride, sr, limit, window, scaling, device=device)
    code, ov, ostats, ostatus, ol = O.spectrum_batched(x, n_frames, n, stride, sr, limit[0], limit[1], limit[2],
                                                       window, scaling)
    assert code == 0
    assert gl == ol and gv.shape == ov.shape                     # FrequencyLimit indices: exact
    assert np.array_equal(gs["status"], ostatus)                 # status codes: exact
    ok = ostatus == 0
    # unscaled magnitudes (for the tolerance scale) from the oracle
    _, ou, _, _, _ = O.spectrum_batched(x, n_frames, n, stride, sr, limit[0], limit[1], limit[2], window,
                                        O.SCALING_NONE)
    gu, _, _ = run_gpu(x, n_frames, n, stride, sr, limit, window, O.SCALING_NONE, device=device)
    fmax = ou.max(axis=1, keepdims=True)
    assert (np.abs(gu - ou)[ok] <= MAG_RTOL * fmax[ok]).all(), \
        f"unscaled magnitudes off: {float((np.abs(gu - ou) / np.maximum(fmax, 1e-30))[ok].max()):.3e} x frame max"
    nf = np.float32(n)
    for f in np.nonzero(ok)[0]:
        # (1) scaling is an exact function of the GPU's own unscaled magnitudes
        pre = O.calc_statistics(gu[f], gl)
        want = np.array([O.scale(scaling, v, pre.min_val, pre.max_val, pre.average, pre.median, nf) for v in gu[f]],
                        np.float32)
        assert np.array_equal(want, gv[f]), f"scaling {scaling} not bit-exact on frame {f}"
        # (2) order statistics are exact functions of the GPU's own scaled values
        s = O.calc_statistics(gv[f], gl)
        assert (s.min_val, s.min_bin) == (gs["min_val"][f], gs["min_bin"][f]), f"min frame {f}"
        assert (s.max_val, s.max_bin) == (gs["max_val"][f], gs["max_bin"][f]), f"max frame {f}"
        assert s.median == gs["median"][f], f"median frame {f}: {s.median} vs {gs['median'][f]}"
        # the average: same values summed in another order (the reference folds the sorted copy, src/spectrum.rs:505-508):
        # the two rounding-error walks are ~sqrt(m) ulp of the mean magnitude apart at most
        m_ = gv[f].size
        assert abs(s.average - gs["average"][f]) <= (8 + np.sqrt(m_)) * 2.0 ** -23 * max(float(np.abs(gv[f]).mean()), 1e-30), \
            f"average frame {f}"
    # (3) scaled values against the oracle's scaled values
    if scaling in (O.SCALING_NONE, O.SCALING_DIVIDE_BY_N, O.SCALING_DIVIDE_BY_N_SQRT):
        div = {O.SCALING_NONE: 1.0, O.SCALING_DIVIDE_BY_N: float(n), O.SCALING_DIVIDE_BY_N_SQRT: float(np.sqrt(n))}[scaling]
        assert (np.abs(gv - ov)[ok] <= MAG_RTOL * (fmax / div)[ok] * 1.0001).all()
    elif scaling == O.SCALING_ZERO_TO_ONE:
        assert (np.abs(gv - ov)[ok] <= 2.5 * MAG_RTOL).all()
    else:  # dB: compare mapped back to linear (relative error explodes for near-zero bins otherwise)
        lin_g = np.where(gv == 0, gu * 0 + np.where(gu == 0, 0, 1), 10.0 ** (gv.astype(np.float64) / 20.0))
        lin_o = np.where(ov == 0, ou * 0 + np.where(ou == 0, 0, 1), 10.0 ** (ov.astype(np.float64) / 20.0))
        assert (np.abs(lin_g - lin_o)[ok] <= 1.5 * MAG_RTOL * fmax[ok]).all()
    return gv, gs


def noise(n_frames, n, seed, amp=1.0):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n_frames * n) * amp).astype(np.float32)


@pytest.mark.parametrize("n", ALL_N)
def test_every_length_every_window(gpu_ctx, n):
    frames = 37 if n <= 4096 else 5
    x = noise(frames, n, seed=n)
    for w in WINDOWS:
        check_against_oracle(x, frames, n, n, 44100, (O.LIMIT_ALL, 0, 0), w, O.SCALING_NONE)


@pytest.mark.parametrize("n", [2, 8, 64, 256, 1024, 2048, 4096, 16384])
@pytest.mark.parametrize("scaling", SCALINGS[1:])
def test_scalings(gpu_ctx, n, scaling):
    frames = 19
    x = noise(frames, n, seed=100 + n, amp=1000.0)
    check_against_oracle(x, frames, n, n, 44100, (O.LIMIT_ALL, 0, 0), O.WINDOW_HANN, scaling)


@pytest.mark.parametrize("n,sr,limit", [
    (2048, 44100, (O.LIMIT_MAX, 0.0, 4000.0)),          # cfg1 (the repo's own test case)
    (4096, 44100, (O.LIMIT_MAX, 0.0, 4000.0)),
    (16384, 44100, (O.LIMIT_RANGE, 50.0, 15000.0)),     # cfg4
    (512, 1024, (O.LIMIT_MIN, 100.0, 0.0)),             # src/tests/mod.rs:224-290
    (512, 1024, (O.LIMIT_MAX, 0.0, 400.0)),
    (512, 1024, (O.LIMIT_RANGE, 412.0, 510.0)),
    (256, 44100, (O.LIMIT_RANGE, 1000.0, 1400.0)),      # 3 bins (odd count)
    (256, 44100, (O.LIMIT_RANGE, 1000.0, 1250.0)),      # 2 bins (even count, smallest legal)
    (64, 44100, (O.LIMIT_MAX, 0.0, 700.0)),
    (8, 44100, (O.LIMIT_MIN, 11000.0, 0.0)),
    (32768, 96000, (O.LIMIT_RANGE, 20.0, 20000.0)),
])
def test_frequency_limits(gpu_ctx, n, sr, limit):
    frames = 11 if n <= 4096 else 3
    x = noise(frames, n, seed=7 * n)
    assert O.limit_to_bins(n, sr, *limit)[1] >= 2
    for sc in (O.SCALING_NONE, O.SCALING_ZERO_TO_ONE, O.SCALING_20_TIMES_LOG10):
        check_against_oracle(x, frames, n, n, sr, limit, O.WINDOW_BH4, sc)


def test_host_pipeline_matches_device_path(gpu_ctx):
    n, frames = 1024, 3001   # several chunks would need > 64 MiB; this at least exercises the host entry
    x = noise(frames, n, seed=5)
    a, sa_, _ = run_gpu(x, frames, n, n, 48000, window=O.WINDOW_HAMMING, scaling=O.SCALING_DIVIDE_BY_N, device=True)
    b, sb, _ = run_gpu(x, frames, n, n, 48000, window=O.WINDOW_HAMMING, scaling=O.SCALING_DIVIDE_BY_N, device=False)
    assert np.array_equal(a, b) and np.array_equal(sa_, sb)


def test_host_pipeline_multi_chunk(gpu_ctx):
    n, frames = 16384, 2500  # 164 MB of input -> 3 chunks of <= 1024 frames
    x = noise(frames, n, seed=6)
    a, sa_, _ = run_gpu(x, frames, n, n, 44100, window=O.WINDOW_HANN, scaling=O.SCALING_DIVIDE_BY_N_SQRT, device=True)
    b, sb, _ = run_gpu(x, frames, n, n, 44100, window=O.WINDOW_HANN, scaling=O.SCALING_DIVIDE_BY_N_SQRT, device=False)
    assert np.array_equal(a, b) and np.array_equal(sa_, sb)


@pytest.mark.parametrize("n,hop", [(2048, 512), (256, 64), (4096, 1000), (1024, 333), (64, 1)])
def test_stft_hop_and_unaligned_strides(gpu_ctx, n, hop):
    frames = 29
    x = noise(1, (frames - 1) * hop + n, seed=hop)
    check_against_oracle(x, frames, n, hop, 44100, (O.LIMIT_ALL, 0, 0), O.WINDOW_HAMMING, O.SCALING_DIVIDE_BY_N)


def test_zero_frames_and_ties(gpu_ctx):
    # all-zero input: every bin exactly 0; min -> lowest bin, max -> highest bin (stable-sort tie rule)
    for n in (8, 256, 4096):
        x = np.zeros(3 * n, np.float32)
        for sc in SCALINGS:
            gv, gs = check_against_oracle(x, 3, n, n, 44100, (O.LIMIT_ALL, 0, 0), O.WINDOW_HANN, sc)
            assert (gv == 0).all() and not np.signbit(gv).any()
            assert (gs["min_bin"] == 0).all() and (gs["max_bin"] == n // 2).all()
            assert (gs["median"] == 0).all() and (gs["average"] == 0).all()


def test_constant_and_impulse_frames(gpu_ctx):
    for n in (64, 1024, 4096):
        x = np.concatenate([np.ones(n), np.eye(1, n, 0)[0], np.eye(1, n, n // 2)[0] * 5, (-1.0) ** np.arange(n)])
        x = x.astype(np.float32)
        gv, gs = check_against_oracle(x, 4, n, n, 8000, (O.LIMIT_ALL, 0, 0), O.WINDOW_NONE, O.SCALING_NONE)
        assert gv[0, 0] == n and np.abs(gv[0, 1:]).max() <= 1e-4 * n        # DC only
        assert np.abs(gv[1] - 1.0).max() <= 1e-4                              # flat spectrum (many ties)
        assert gv[3, n // 2] == n and np.abs(gv[3, :-1]).max() <= 1e-4 * n    # Nyquist only


def test_nan_inf_status_precedence(gpu_ctx):
    n, frames = 512, 8
    x = noise(frames, n, seed=9)
    x[1 * n + 17] = np.nan
    x[2 * n + 400] = np.inf
    x[3 * n + 5] = -np.inf
    x[3 * n + 300] = np.nan            # NaN wins over Inf (src/lib.rs:168-173)
    x[4 * n + 0] = np.inf              # Hann multiplier 0 at i = 0: 0 * inf = NaN in the windowed samples
    x[5 * n + 10] = 3.0e38             # finite input, FFT overflows -> non-finite result
    x[5 * n + 11] = 3.0e38
    for w in (O.WINDOW_NONE, O.WINDOW_HANN):
        gv, gs, _ = run_gpu(x, frames, n, n, 44100, window=w, scaling=O.SCALING_DIVIDE_BY_N)
        code, ov, ostats, ostatus, _ = O.spectrum_batched(x, frames, n, n, 44100, 0, 0, 0, w, O.SCALING_DIVIDE_BY_N)
        assert np.array_equal(gs["status"], ostatus), (gs["status"], ostatus)
        assert ostatus[1] == O.ERR_NAN_VALUES and ostatus[2] == O.ERR_INFINITY_VALUES and ostatus[3] == O.ERR_NAN_VALUES
        assert ostatus[4] == (O.ERR_NAN_VALUES if w == O.WINDOW_HANN else O.ERR_INFINITY_VALUES)
        assert ostatus[0] == 0 and ostatus[6] == 0 and ostatus[7] == 0
        good = ostatus == 0
        assert (np.abs(gv - ov)[good] <= MAG_RTOL * ov[good].max()).all()
    assert ostatus[5] == O.ERR_NON_FINITE_RESULT or ostatus[5] == 0


def test_stats_mask_and_no_stats(gpu_ctx):
    n, frames = 2048, 9
    x = noise(frames, n, seed=21)
    full_v, full_s, _ = run_gpu(x, frames, n, n, 44100, scaling=O.SCALING_ZERO_TO_ONE, stats=7)
    v0, _, _ = run_gpu(x, frames, n, n, 44100, scaling=O.SCALING_ZERO_TO_ONE, stats=0)
    assert np.array_equal(full_v, v0)
    v3, s3, _ = run_gpu(x, frames, n, n, 44100, scaling=O.SCALING_ZERO_TO_ONE, stats=3)
    assert np.array_equal(full_v, v3)
    assert np.array_equal(s3["max_val"], full_s["max_val"]) and np.array_equal(s3["average"], full_s["average"])
    assert (s3["median"] == 0).all()


def test_noise_generator_matches_oracle(gpu_ctx):
    import torch
    m = sa()
    out = torch.empty(100003, dtype=torch.float32, device="cuda")
    ctx = m.api._torch_ctx(0, torch.cuda.current_stream().cuda_stream)
    m.api._check(m.cabi.load().sa_generate_noise(ctx._h, out.data_ptr(), 12345, out.numel(), 0xABCDEF))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), O.generate_noise(12345, out.numel(), 0xABCDEF))


def test_windows_batched_bit_exact(gpu_ctx):
    m = sa()
    rng = np.random.default_rng(3)
    for n in (1, 2, 4, 5, 100, 4096):
        x = rng.standard_normal((3, n)).astype(np.float32)
        for name, kind in (("hann_window", O.WINDOW_HANN), ("hamming_window", O.WINDOW_HAMMING),
                           ("blackman_harris_4term", O.WINDOW_BH4), ("blackman_harris_7term", O.WINDOW_BH7)):
            got = getattr(m.windows, name)(x)
            want = np.stack([O.window_apply(kind, row) for row in x])
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (name, n)


def test_apply_scaling_batched_repeated(gpu_ctx):
    """benches/fft_spectrum_bench.rs:21-36: divide_by_N_sqrt in the main call + 3 more times."""
    m = sa()
    n = 2048
    rng = np.random.default_rng(11)
    x = rng.integers(-32768, 32767, n).astype(np.float32)
    w = O.window_apply(O.WINDOW_HANN, x)
    spec = m.samples_fft_to_spectrum(w, 44100, m.FrequencyLimit.All, m.scaling.divide_by_N_sqrt)
    ref = O.samples_fft_to_spectrum(w, 44100, O.LIMIT_ALL, 0, 0, O.SCALING_DIVIDE_BY_N_SQRT)
    vals = spec.values().copy()
    for _ in range(3):
        spec.apply_scaling_fn(m.scaling.divide_by_N_sqrt)
        vals = (vals / np.float32(np.sqrt(np.float32(n)))).astype(np.float32)
        assert np.array_equal(spec.values(), vals)
        s = O.calc_statistics(spec.values())
        assert (s.min_val, s.max_val, s.median) == (spec.min()[1], spec.max()[1], spec.median())
    assert np.abs(spec.values() * np.float32(n) ** 1.5 - ref.vals).max() <= MAG_RTOL * ref.vals.max()


def test_combined_scaling_chain(gpu_ctx):
    """scaling::combined (src/scaling.rs:174-182): the whole chain sees the SAME pre-call statistics."""
    m = sa()
    n = 1024
    x = noise(1, n, seed=31, amp=300.0)
    w = O.window_apply(O.WINDOW_HANN, x)
    for chain_gpu, chain_cpu in [
        ([m.scaling.divide_by_N, m.scaling.scale_20_times_log10], [O.SCALING_DIVIDE_BY_N, O.SCALING_20_TIMES_LOG10]),
        ([m.scaling.scale_20_times_log10, m.scaling.divide_by_N, m.scaling.divide_by_N_sqrt],
         [O.SCALING_20_TIMES_LOG10, O.SCALING_DIVIDE_BY_N, O.SCALING_DIVIDE_BY_N_SQRT]),      # scaling.rs:214
        ([m.scaling.scale_to_zero_to_one, m.scaling.scale_to_zero_to_one], [O.SCALING_ZERO_TO_ONE] * 2),
    ]:
        spec = m.samples_fft_to_spectrum(w, 44100, m.FrequencyLimit.All, None)
        before = spec.values().copy()
        pre = O.calc_statistics(before)
        spec.apply_scaling_fn(m.scaling.combined(chain_gpu))
        want = before.copy()
        for k in chain_cpu:   # every step with the statistics of the spectrum BEFORE the call
            want = np.array([O.scale(k, v, pre.min_val, pre.max_val, pre.average, pre.median, np.float32(n)) for v in want],
                            np.float32)
        assert np.array_equal(spec.values(), want)
        s = O.calc_statistics(spec.values())
        assert (s.min_val, s.max_val, s.median) == (spec.min()[1], spec.max()[1], spec.median())
    with pytest.raises(TypeError):
        m.scaling.combined([m.scaling.divide_by_N, lambda v, s: v])


@pytest.mark.parametrize("n,hop,offset", [(8, 8, 0), (32, 16, 0), (64, 64, 0), (256, 128, 0), (512, 512, 2),
                                          (1024, 1024, 0), (2048, 512, 0), (4096, 4096, 0), (4096, 1001, 0),
                                          (4096, 4096, 1), (8192, 8192, 0), (16384, 16384, 0), (32768, 32768, 0)])
def test_int16_ingestion_equals_f32_path(gpu_ctx, n, hop, offset):
    """Fused i16 PCM ingestion (`x as f32` on load, reference callers: src/tests/mod.rs:70-73) must give the
    very same bits as converting on the host and calling the f32 entry point -- in every kernel family, for
    aligned and unaligned (odd hop / odd offset) inputs, on the device and through the host pipeline."""
    import torch
    m = sa()
    rng = np.random.default_rng(n + hop)
    frames = 37
    pcm = rng.integers(-32768, 32768, size=offset + (frames - 1) * hop + n, dtype=np.int16)
    pcm[:4] = [-32768, 32767, 0, -1]
    pcm = pcm[offset:]
    lim = m.FrequencyLimit.All
    kw = dict(window=O.WINDOW_HANN, n=n, frame_stride=hop, n_frames=frames, stats=7)
    full = torch.from_numpy(np.concatenate([np.zeros(offset, np.int16), pcm])).cuda()
    # same misalignment for the f32 arm, so that both arms dispatch to the same kernel family
    ref = m.samples_fft_to_spectrum_batched(full.to(torch.float32)[offset:], 44100, lim,
                                            m.scaling.divide_by_N_sqrt, **kw)
    got = m.samples_fft_to_spectrum_batched(full[offset:], 44100, lim, m.scaling.divide_by_N_sqrt, **kw)
    torch.cuda.synchronize()
    assert np.array_equal(ref.vals.cpu().numpy().view(np.uint32), got.vals.cpu().numpy().view(np.uint32))
    assert ref.stats_numpy().tobytes() == got.stats_numpy().tobytes()
    if offset == 0:  # the host pipeline's slots are always aligned
        host = m.samples_fft_to_spectrum_batched(pcm, 44100, lim, m.scaling.divide_by_N_sqrt, **kw)
        assert np.array_equal(host.vals.view(np.uint32), got.vals.cpu().numpy().view(np.uint32))
        assert host.stats.tobytes() == got.stats_numpy().tobytes()
    # and against the oracle fed the converted samples
    x = pcm.astype(np.float32)
    code, ov, _, ostatus, _ = O.spectrum_batched(x, frames, n, hop, 44100, O.LIMIT_ALL, 0.0, 0.0, O.WINDOW_HANN,
                                                 O.SCALING_DIVIDE_BY_N_SQRT)
    assert code == 0 and np.all(ostatus == 0)
    gv = got.vals.cpu().numpy()[:, : got.n_bins]
    assert np.all(np.abs(gv - ov) <= MAG_RTOL * ov.max(axis=1, keepdims=True))


@pytest.mark.parametrize("n,counts", [(2048, [1, 7, 8, 9, 1183, 1184, 1185]), (4096, [1, 3, 4, 5, 591, 592, 593, 1185]),
                                      (8192, [1, 2, 3, 295, 296, 297]), (16384, [1, 2, 147, 148, 149, 300])])
def test_frame_counts_around_grid_boundaries(gpu_ctx, n, counts):
    """The tuned kernels are persistent (one CTA per SM, G frame groups per CTA, frames dealt round-robin):
    frame counts below / at / just above G and 148*G must give, frame for frame, the very bits of a longer run
    (every frame is a pure function of its own samples, src/lib.rs:157-337; the median is exact whatever window
    the group guessed from its previous frame)."""
    import torch
    m = sa()
    total = max(counts)
    x = torch.from_numpy(noise(total, n, seed=7 * n)).cuda()
    kw = dict(window=O.WINDOW_HANN, n=n, frame_stride=n, stats=7)
    ref = m.samples_fft_to_spectrum_batched(x, 44100, m.FrequencyLimit.All, m.scaling.divide_by_N_sqrt, n_frames=total, **kw)
    torch.cuda.synchronize()
    rv, rs = ref.vals.cpu().numpy().view(np.uint32), ref.stats_numpy()
    assert np.all(rs["status"] == 0)
    for f in counts:
        got = m.samples_fft_to_spectrum_batched(x, 44100, m.FrequencyLimit.All, m.scaling.divide_by_N_sqrt, n_frames=f, **kw)
        torch.cuda.synchronize()
        assert np.array_equal(got.vals.cpu().numpy().view(np.uint32), rv[:f]), f"values differ at n_frames={f}"
        assert got.stats_numpy().tobytes() == rs[:f].tobytes(), f"stats differ at n_frames={f}"
    # and the long run itself against the oracle on its last frames (the ones dealt last)
    code, ov, _, ostatus, _ = O.spectrum_batched(x.cpu().numpy()[(total - 4) * n:], 4, n, n, 44100, O.LIMIT_ALL, 0.0, 0.0,
                                                 O.WINDOW_HANN, O.SCALING_DIVIDE_BY_N_SQRT)
    assert code == 0 and np.all(ostatus == 0)
    gv = ref.vals.cpu().numpy()[total - 4:, : ref.n_bins]
    assert np.all(np.abs(gv - ov) <= MAG_RTOL * ov.max(axis=1, keepdims=True))


@pytest.mark.parametrize("n", [256, 1024, 2048, 4096, 8192, 16384])
def test_repeat_launch_is_bit_identical(gpu_ctx, n):
    """Two launches on the same input give the same bits -- values and the whole 32-byte records.  (The kernels
    use shared-memory atomics, 'first warp to finish' protocols and asynchronous copies; any race in them would
    show up here as a difference between runs.)"""
    import torch
    m = sa()
    frames = (1 << 22) // n + 13
    x = torch.from_numpy(noise(frames, n, seed=n + 1)).cuda()
    kw = dict(window=O.WINDOW_HANN, n=n, frame_stride=n, n_frames=frames, stats=7)
    runs = []
    for _ in range(3):
        r = m.samples_fft_to_spectrum_batched(x, 44100, m.FrequencyLimit.All, m.scaling.scale_20_times_log10, **kw)
        torch.cuda.synchronize()
        runs.append((r.vals.view(torch.int32).clone(), r.stats.clone()))
    for v, s in runs[1:]:
        assert torch.equal(v, runs[0][0]) and torch.equal(s, runs[0][1])


@pytest.mark.parametrize("n", [256, 4096])
def test_fast_log10_option(gpu_ctx, n):
    """sa_ctx_set_fast_log10 (opt-in): dB values through lg2.approx stay within 4 ulp of the exact libm
    sequence (src/scaling.rs:104-114), zeros stay zeros, and the statistics are still exact order statistics
    of the values written.  The default (off) is untouched."""
    import torch
    m = sa()
    frames = 64
    x = torch.from_numpy(noise(frames, n, seed=99, amp=1000.0)).cuda()
    x[: n] = 0.0                                                         # an all-zero frame: every dB value is 0
    kw = dict(window=O.WINDOW_HANN, n=n, frame_stride=n, n_frames=frames, stats=7)
    exact = m.samples_fft_to_spectrum_batched(x, 44100, m.FrequencyLimit.All, m.scaling.scale_20_times_log10, **kw)
    torch.cuda.synchronize()
    ctx = m.Context(0, torch.cuda.current_stream().cuda_stream)
    ctx.set_fast_log10(True)
    fast = m.samples_fft_to_spectrum_batched(x, 44100, m.FrequencyLimit.All, m.scaling.scale_20_times_log10, ctx=ctx, **kw)
    torch.cuda.synchronize()
    ev, fv = exact.vals.cpu().numpy(), fast.vals.cpu().numpy()
    assert np.all(fv[0] == 0.0) and np.all(ev[0] == 0.0)
    # lg2.approx carries a relative error of about 2^-22; both results are then rounded to f32
    # (ulp of an 80 dB value: 7.6e-6 dB)
    err = np.abs(ev - fv) / (np.spacing(np.abs(ev)) + 1e-6)
    assert err.max() <= 4.0, f"fast log10 off by {err.max():.2f} ulp"
    assert not np.array_equal(ev, fv)                                    # it really is another code path
    fs = fast.stats_numpy()
    for f in range(0, frames, 7):
        s = O.calc_statistics(fv[f])
        assert (s.min_val, s.min_bin, s.max_val, s.max_bin, s.median) == \
            (fs["min_val"][f], fs["min_bin"][f], fs["max_val"][f], fs["max_bin"][f], fs["median"][f])
    ctx.set_fast_log10(False)
    again = m.samples_fft_to_spectrum_batched(x, 44100, m.FrequencyLimit.All, m.scaling.scale_20_times_log10, ctx=ctx, **kw)
    torch.cuda.synchronize()
    assert np.array_equal(again.vals.cpu().numpy(), ev)
