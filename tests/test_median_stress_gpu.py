"""Stress of the exact-median machinery of the tuned / small kernels: the fast path guesses a key
window from the previous frame of the same frame group, so adversarial SEQUENCES of frames (scale
jumps, silence, impulses, heavy ties, quantised input) must hit the fallbacks (window miss, crowded
bucket narrowing, all-equal) and still give statistics that are exact functions of the GPU's own values.

Every frame of every batch is checked against torch's exact order statistics on the device and a sample
against the oracle's stable-sort statistics (src/spectrum.rs:481-539)."""
import os

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

# soak runs: SA_STRESS_SEED=k shifts every generator seed (profiles/README.md: eight seeds after the round-2 median changes)
SEED0 = int(os.environ.get("SA_STRESS_SEED", "0")) * 1000003


def _adversarial_batch(rng, frames, n):
    x = rng.standard_normal((frames, n)).astype(np.float32)
    kind = rng.integers(0, 10, frames)
    scale = np.where(kind < 3, 10.0 ** rng.uniform(-6, 6, frames), 1.0).astype(np.float32)
    x *= scale[:, None]                                       # median jumps by orders of magnitude
    x[kind == 3] = 0.0                                        # silence: every bin exactly 0
    imp = np.nonzero(kind == 4)[0]
    x[imp] = 0.0
    x[imp, rng.integers(0, n, imp.size)] = 3.0                # impulse: flat spectrum, near-equal magnitudes
    q = kind == 5
    x[q] = np.round(x[q] * 2.0) / 2.0                         # coarse quantisation: many repeated values
    c = kind == 6
    x[c] = 1.0                                                # constant: DC only, everything else ~0
    s = np.nonzero(kind == 7)[0]
    t = np.arange(n, dtype=np.float32)
    for f in s:                                               # pure tones: two huge bins, the rest leakage
        x[f] = np.sin(2 * np.pi * t * float(rng.integers(1, n // 2)) / n).astype(np.float32)
    d = kind == 8
    x[d] *= np.float32(1e-30)                                 # tiny values (denormal squares)
    return x


def _check(torch, sa, x, n, limit, scaling, window, sr=44100):
    frames = x.shape[0]
    xd = torch.from_numpy(x.reshape(-1)).cuda()
    r = sa.samples_fft_to_spectrum_batched(xd, sr, limit, scaling, window=window, n=n, frame_stride=n, n_frames=frames)
    torch.cuda.synchronize()
    nb = r.n_bins
    vals = r.vals[:, :nb]
    st_f = r.stats.view(torch.float32).view(frames, 8)
    st_i = r.stats.view(torch.int32).view(frames, 8)
    ok = st_i[:, 6] == 0
    assert bool(ok.all()), "unexpected non-OK status"
    srt = vals.sort(dim=1).values
    mid = nb // 2
    med = srt[:, mid] if nb % 2 else (srt[:, mid - 1] + srt[:, mid]) / 2.0
    bad = torch.nonzero(med != st_f[:, 5]).flatten()
    assert bad.numel() == 0, f"median mismatch on frames {bad[:8].tolist()} (n={n}, bins={nb})"
    assert torch.equal(srt[:, 0], st_f[:, 0]) and torch.equal(srt[:, -1], st_f[:, 2])
    # tie rules: min -> lowest bin, max -> highest bin among equals
    lo_bin = (vals == st_f[:, 0:1]).float().argmax(dim=1) + r.bin_lo
    hi_bin = nb - 1 - (vals.flip(1) == st_f[:, 2:3]).float().argmax(dim=1) + r.bin_lo
    assert torch.equal(lo_bin.int(), st_i[:, 1]) and torch.equal(hi_bin.int(), st_i[:, 3])
    # a sample of frames against the oracle's stable sort (incl. the average within tolerance)
    hv = vals.cpu().numpy()
    hs = r.stats_numpy()
    for f in list(range(0, frames, max(1, frames // 40)))[:48]:
        s = O.calc_statistics(hv[f], r.bin_lo)
        assert (s.min_val, s.min_bin, s.max_val, s.max_bin, s.median) == \
            (hs["min_val"][f], hs["min_bin"][f], hs["max_val"][f], hs["max_bin"][f], hs["median"][f]), f
        # average: the GPU sums pairwise (error ~ log2(n) ulp of mean|v|), the reference folds the sorted copy sequentially
        # in f32 (src/spectrum.rs:505-508: error up to ~ n ulp -- for 8193 near-equal values it leaves [min, max]).  Each
        # is checked against the f64 mean with its own bound; soak seeds found frames where 1e-4 x max between the
        # two was too tight for the REFERENCE's rounding, not for the kernel's.
        exact = float(np.mean(hv[f].astype(np.float64)))
        scale = max(float(np.mean(np.abs(hv[f].astype(np.float64)))), 1e-30)
        eps = float(np.finfo(np.float32).eps)
        assert abs(float(hs["average"][f]) - exact) <= 32 * eps * scale, (f, float(hs["average"][f]), exact)
        assert abs(float(s.average) - exact) <= (hv[f].size + 8) * eps * scale, (f, float(s.average), exact)


@pytest.mark.parametrize("n,frames", [(4096, 6000), (1024, 20000), (2048, 9000), (8192, 2500), (16384, 1200),
                                      (256, 70000), (64, 150000), (512, 40000), (128, 90000)])
def test_adversarial_sequences_all_limit(gpu_ctx, n, frames):
    import torch
    import spectrum_analyzer_b200 as sa
    rng = np.random.default_rng(n + SEED0)
    x = _adversarial_batch(rng, frames, n)
    for scaling, window in [(sa.scaling.divide_by_N_sqrt, "hann"), (None, None), (sa.scaling.scale_to_zero_to_one, "hamming"),
                            (sa.scaling.scale_20_times_log10, "blackman_harris_4term")]:
        _check(torch, sa, x, n, sa.FrequencyLimit.All, scaling, window)


@pytest.mark.parametrize("n,limit", [(4096, (100.0, 9000.0)), (4096, (0.0, 10000.0)), (1024, (500.0, 15000.0)),
                                     (256, (1000.0, 20000.0)), (16384, (50.0, 15000.0)), (512, (0.0, 8000.0))])
def test_adversarial_sequences_even_and_odd_counts(gpu_ctx, n, limit):
    import torch
    import spectrum_analyzer_b200 as sa
    rng = np.random.default_rng(n + 1 + SEED0)
    frames = 3000 if n >= 4096 else 12000
    x = _adversarial_batch(rng, frames, n)
    for hi_adj in (0.0, 11.0 * 44100 / n):     # shift the upper limit by some bins: covers even and odd bin counts
        lim = sa.FrequencyLimit.Range(limit[0], limit[1] + hi_adj)
        for scaling in (sa.scaling.divide_by_N, sa.scaling.scale_20_times_log10):
            _check(torch, sa, x, n, lim, scaling, "hann")


def test_stationary_then_jump_sequence(gpu_ctx):
    """Long stationary run (window shrinks to its floor) followed by abrupt level changes."""
    import torch
    import spectrum_analyzer_b200 as sa
    rng = np.random.default_rng(99 + SEED0)
    n, frames = 4096, 20000
    x = rng.standard_normal((frames, n)).astype(np.float32)
    x[12000:] *= np.float32(1.0001)        # tiny drift: still inside the window
    x[15000:] *= np.float32(7.0)           # jump: window miss on the first frame of every group
    x[17000:17003] = 0.0
    x[18000:] *= np.float32(1e-4)
    _check(torch, sa, x, n, sa.FrequencyLimit.All, sa.scaling.divide_by_N_sqrt, "hann")
