"""BASELINE.json's full sizes through size-independent properties (the oracle would need minutes):
  cfg2: 1 000 000 x 4096 Hann / All / divide_by_N_sqrt / all stats
  cfg5: 64M x 256 is 65 GB of input -- run here at 8M frames (same kernel, same packing)
  cfg3: STFT hop 512 over a long stream;  cfg4: 16384-point BH4, Range(50,15000), zero_to_one
Properties: Parseval (energy of the windowed frame == energy of the unscaled spectrum), determinism
(two runs bit-identical), status all OK, min <= median <= max and min <= average <= max per frame,
argmin/argmax really point at the min/max, plus sampled frames checked against the oracle."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


def _run(torch, sa, x, frames, n, hop, limit, scaling, window, stats=7):
    r = sa.samples_fft_to_spectrum_batched(x, 44100, limit, scaling, window=window, n=n, frame_stride=hop,
                                           n_frames=frames, stats=stats)
    torch.cuda.synchronize()
    return r


def _stat_fields(torch, r, frames):
    f = r.stats.view(torch.float32).view(frames, 8)
    i = r.stats.view(torch.int32).view(frames, 8)
    return dict(min_val=f[:, 0], min_bin=i[:, 1], max_val=f[:, 2], max_bin=i[:, 3], average=f[:, 4], median=f[:, 5],
                status=i[:, 6])


def _noise(torch, sa, count, seed=0x5A17):
    x = torch.empty(count, dtype=torch.float32, device="cuda")
    ctx = sa.api._torch_ctx(0, torch.cuda.current_stream().cuda_stream)
    sa.api._check(sa.cabi.load().sa_generate_noise(ctx._h, x.data_ptr(), 0, count, seed))
    return x


def _check_sampled_frames(torch, x, r, frames, n, hop, window, limit, scaling, idx):
    vals = r.vals[idx].cpu().numpy()[:, : r.n_bins]
    st = r.stats_numpy()[idx]
    for row, f in enumerate(idx):
        frame = x[f * hop: f * hop + n].cpu().numpy()
        ref = O.samples_fft_to_spectrum(O.window_apply(window, frame), 44100, limit[0], limit[1], limit[2], scaling)
        assert ref.code == 0
        unscaled = O.samples_fft_to_spectrum(O.window_apply(window, frame), 44100, limit[0], limit[1], limit[2])
        tol = 1e-4 * unscaled.vals.max()
        if scaling == O.SCALING_DIVIDE_BY_N_SQRT:
            tol /= np.sqrt(n)
        elif scaling == O.SCALING_ZERO_TO_ONE:
            tol = 2.5e-4
        assert np.abs(vals[row] - ref.vals).max() <= tol
        s = O.calc_statistics(vals[row], r.bin_lo)                # stats exact w.r.t. the GPU's own values
        assert (s.min_val, s.min_bin, s.max_val, s.max_bin, s.median) == \
            (st["min_val"][row], st["min_bin"][row], st["max_val"][row], st["max_bin"][row], st["median"][row])


def test_cfg2_one_million_frames_4096(gpu_ctx):
    import torch
    import spectrum_analyzer_b200 as sa
    frames, n = 1_000_000, 4096
    x = _noise(torch, sa, frames * n)
    r = _run(torch, sa, x, frames, n, n, sa.FrequencyLimit.All, sa.scaling.divide_by_N_sqrt, "hann")
    s = _stat_fields(torch, r, frames)
    assert int((s["status"] != 0).sum()) == 0
    assert bool(((s["min_val"] <= s["median"]) & (s["median"] <= s["max_val"])).all())
    assert bool(((s["min_val"] <= s["average"]) & (s["average"] <= s["max_val"])).all())
    rows = torch.arange(frames, device="cuda")
    assert torch.equal(r.vals[rows, s["max_bin"].long()], s["max_val"])
    assert torch.equal(r.vals[rows, s["min_bin"].long()], s["min_val"])
    assert torch.equal(r.vals.max(dim=1).values, s["max_val"]) and torch.equal(r.vals.min(dim=1).values, s["min_val"])
    # determinism: second run bit-identical
    r2 = _run(torch, sa, x, frames, n, n, sa.FrequencyLimit.All, sa.scaling.divide_by_N_sqrt, "hann")
    assert torch.equal(r.vals, r2.vals) and torch.equal(r.stats, r2.stats)
    del r2
    # Parseval on a slab of frames: sum(w x)^2 == (|X0|^2 + |XN/2|^2 + 2 sum |Xk|^2) / N, X = sqrt(N) * vals
    w = torch.from_numpy(sa.windows.coefficients("hann", n)).cuda()
    for f0 in (0, 499_000, frames - 4096):
        xs = (x[f0 * n:(f0 + 4096) * n].view(4096, n) * w).double()
        e_time = (xs * xs).sum(dim=1)
        v = r.vals[f0:f0 + 4096].double() ** 2 * n
        e_freq = (v[:, 0] + v[:, -1] + 2 * v[:, 1:-1].sum(dim=1)) / n
        assert float(((e_time - e_freq).abs() / e_time).max()) < 2e-5
    # sampled frames against the oracle (values, stats)
    idx = [0, 1, 777, 123_456, 500_000, 999_998, 999_999]
    _check_sampled_frames(torch, x, r, frames, n, n, O.WINDOW_HANN, (O.LIMIT_ALL, 0, 0), O.SCALING_DIVIDE_BY_N_SQRT, idx)
    # median == torch's exact median (odd count: the middle element)
    med = r.vals[:65536].median(dim=1).values
    assert torch.equal(med, s["median"][:65536])


def test_cfg5_small_n_many_frames_256(gpu_ctx):
    import torch
    import spectrum_analyzer_b200 as sa
    frames, n = 8 * 1024 * 1024, 256
    x = _noise(torch, sa, frames * n, seed=5)
    r = _run(torch, sa, x, frames, n, n, sa.FrequencyLimit.All, sa.scaling.scale_20_times_log10, None)
    s = _stat_fields(torch, r, frames)
    assert int((s["status"] != 0).sum()) == 0
    assert bool(((s["min_val"] <= s["median"]) & (s["median"] <= s["max_val"])).all())
    med = r.vals[:262144].median(dim=1).values                    # 129 bins (odd): exact middle element
    assert torch.equal(med, s["median"][:262144])
    idx = [0, 5, 4_000_000, frames - 1]
    vals = r.vals[idx].cpu().numpy()
    for row, f in enumerate(idx):
        frame = x[f * n:(f + 1) * n].cpu().numpy()
        ref_u = O.samples_fft_to_spectrum(frame, 44100)
        lin = np.where(vals[row] == 0, 0.0, 10.0 ** (vals[row].astype(np.float64) / 20.0))
        assert np.abs(lin - ref_u.vals).max() <= 1.5e-4 * ref_u.vals.max()


def test_cfg3_stft_hop_512(gpu_ctx):
    import torch
    import spectrum_analyzer_b200 as sa
    n, hop = 2048, 512
    n_samples = 400_000_000                                        # 2.5 h of 44.1 kHz audio (10 h = 6.4 GB also fits)
    frames = (n_samples - n) // hop + 1
    x = _noise(torch, sa, n_samples, seed=3)
    r = _run(torch, sa, x, frames, n, hop, sa.FrequencyLimit.All, None, "hamming")
    s = _stat_fields(torch, r, frames)
    assert int((s["status"] != 0).sum()) == 0
    assert bool(((s["min_val"] <= s["median"]) & (s["median"] <= s["max_val"])).all())
    idx = [0, 1, 2, 3, 4, frames // 2, frames - 1]
    _check_sampled_frames(torch, x, r, frames, n, hop, O.WINDOW_HAMMING, (O.LIMIT_ALL, 0, 0), O.SCALING_NONE, idx)
    # overlapping frames share samples: frame f+4 starts where frame f ends -> identical to a dense run
    dense = _run(torch, sa, x[: 4096 * n], 4096, n, n, sa.FrequencyLimit.All, None, "hamming")
    assert torch.equal(dense.vals, r.vals[0:4096 * 4:4])


def test_cfg4_large_n_range_limit(gpu_ctx):
    import torch
    import spectrum_analyzer_b200 as sa
    frames, n = 65536, 16384
    x = _noise(torch, sa, frames * n, seed=4)
    lim = sa.FrequencyLimit.Range(50.0, 15000.0)
    r = _run(torch, sa, x, frames, n, n, lim, sa.scaling.scale_to_zero_to_one, "blackman_harris_4term", stats=3)
    assert (r.bin_lo, r.n_bins) == (19, 5554)
    s = _stat_fields(torch, r, frames)
    assert int((s["status"] != 0).sum()) == 0
    assert bool((s["max_val"] == 1.0).all()) and bool((s["min_val"] >= 0).all())
    rows = torch.arange(frames, device="cuda")
    assert torch.equal(r.vals[rows, (s["max_bin"] - 19).long()], s["max_val"])
    r7 = _run(torch, sa, x[: 1024 * n], 1024, n, n, lim, sa.scaling.scale_to_zero_to_one, "blackman_harris_4term")
    _check_sampled_frames(torch, x, r7, 1024, n, n, O.WINDOW_BH4, (O.LIMIT_RANGE, 50.0, 15000.0), O.SCALING_ZERO_TO_ONE,
                          [0, 511, 1023])
    assert torch.equal(r7.vals, r.vals[:1024])


@pytest.mark.parametrize("dtype,hop", [("f32", 4096), ("i16", 1000)])
def test_pageable_host_buffers_through_the_staging_ring(gpu_ctx, dtype, hop):
    """sa_spectrum_batched_host from plain (pageable) numpy arrays: more chunks than pipeline slots, so every staging
    block is reused; values and records must equal the device-resident call bit for bit."""
    import numpy as np
    import torch
    import spectrum_analyzer_b200 as sa
    n, frames = 4096, 30_011 if hop == 4096 else 90_017
    rng = np.random.default_rng(8)
    if dtype == "f32":
        x = rng.standard_normal((frames - 1) * hop + n).astype(np.float32)
    else:
        x = rng.integers(-32768, 32767, (frames - 1) * hop + n).astype(np.int16)
    lim = sa.FrequencyLimit.Range(100.0, 18000.0)
    host = sa.samples_fft_to_spectrum_batched(x, 44100, lim, sa.scaling.scale_20_times_log10, window="hann", n=n,
                                              frame_stride=hop, n_frames=frames)
    dev = sa.samples_fft_to_spectrum_batched(torch.from_numpy(x).cuda(), 44100, lim, sa.scaling.scale_20_times_log10,
                                             window="hann", n=n, frame_stride=hop, n_frames=frames)
    torch.cuda.synchronize()
    assert np.array_equal(host.vals.view(np.uint32), dev.vals.cpu().numpy().view(np.uint32))
    assert np.array_equal(host.stats.view(np.uint8), dev.stats.cpu().numpy().view(np.uint8).reshape(-1))
