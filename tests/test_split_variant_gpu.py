"""The warp-specialised variant of the 4096-point kernel (csrc/sa_spectrum_v3.cuh, opt-in with SA_SPLIT=1: transform
warps + statistics warps) must produce the very same bytes as the default one-role kernel -- values AND records --
for every mode, limit, mask and input form, including the rare paths of the exact median (first frames, window
misses, crowded buckets, even bin counts) and non-finite frames.  The default kernel is checked against the oracle by
test_parity_gpu.py / test_median_stress_gpu.py, so equality here carries that parity over."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 4096
SR = 44100


def sa():
    import spectrum_analyzer_b200 as m
    return m


@pytest.fixture(scope="module")
def ctx_pair(gpu_ctx):
    """(default context, context created with SA_SPLIT=1), both on torch's current stream."""
    import torch
    m = sa()
    stream = torch.cuda.current_stream().cuda_stream
    a = m.Context(0, stream)
    os.environ["SA_SPLIT"] = "1"
    try:
        b = m.Context(0, stream)
    finally:
        del os.environ["SA_SPLIT"]
    yield a, b
    a.close()
    b.close()


def run_both(ctx_pair, x, limit, scaling, stats=7, window="hann", **kw):
    import torch
    m = sa()
    outs = []
    for ctx in ctx_pair:
        launches = ctx.launch_count
        r = m.samples_fft_to_spectrum_batched(x, SR, limit, scaling, window=window, stats=stats, ctx=ctx, **kw)
        torch.cuda.synchronize()
        assert ctx.launch_count == launches + 1
        outs.append((r.vals.clone(), r.stats.clone() if stats else None, r))
    (va, sa_, ra), (vb, sb, rb) = outs
    assert torch.equal(va.view(torch.int32), vb.view(torch.int32)), "values differ between the two kernels"
    if stats:
        assert torch.equal(sa_, sb), "records differ between the two kernels"
    return ra


def noise(frames, seed, amp=1.0):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(frames, N, generator=g, device="cuda", dtype=torch.float32) * amp


@pytest.mark.parametrize("scaling", ["none", "divide_by_N", "divide_by_N_sqrt", "scale_to_zero_to_one", "scale_20_times_log10"])
@pytest.mark.parametrize("limit", [("all",), ("max", 4000.0), ("range", 50.0, 12000.0), ("range", 50.0, 12005.0), ("min", 21000.0)])
def test_same_bytes_all_modes_and_limits(ctx_pair, scaling, limit):
    m = sa()
    lim = {"all": lambda: m.FrequencyLimit.All, "max": m.FrequencyLimit.Max, "min": m.FrequencyLimit.Min,
           "range": m.FrequencyLimit.Range}[limit[0]](*limit[1:])
    sc = None if scaling == "none" else getattr(m.scaling, scaling)
    x = noise(3001, seed=hash((scaling, limit)) & 0xffff, amp=40.0)
    r = run_both(ctx_pair, x, lim, sc)
    st = r.stats_numpy()
    assert (st["status"] == 0).all()
    assert ((st["min_val"] <= st["median"]) & (st["median"] <= st["max_val"])).all()


@pytest.mark.parametrize("stats", [0, 1, 2, 3, 4, 5, 7])
def test_same_bytes_every_stats_mask(ctx_pair, stats):
    m = sa()
    run_both(ctx_pair, noise(1500, seed=stats), m.FrequencyLimit.All, m.scaling.divide_by_N_sqrt, stats=stats)


@pytest.mark.parametrize("frames", [1, 2, 3, 7, 592, 593, 1777])
def test_same_bytes_few_frames_and_ragged_grids(ctx_pair, frames):
    """1..3 frames: only first-frame (window-less) selections; 592/593: one frame per group, and one group with two."""
    m = sa()
    run_both(ctx_pair, noise(frames, seed=frames), m.FrequencyLimit.All, m.scaling.divide_by_N_sqrt, window="hamming")


def test_same_bytes_hard_medians_and_nonfinite_frames(ctx_pair):
    """Frames built to defeat the guessed window: constant frames (all keys equal), two-valued spectra (crowded bucket),
    amplitude jumps of 1e6 between consecutive frames of a group (window miss), zeros, and NaN / Inf samples."""
    import torch
    m = sa()
    frames = 4 * 148 * 6 + 5
    x = noise(frames, seed=77)
    amp = torch.ones(frames, device="cuda")
    amp[::3] = 1e6
    amp[1::7] = 1e-6
    x = x * amp[:, None]
    x[5] = 0.0                                   # all-zero frame: every value equal
    x[6] = 1.0                                   # constant: one non-zero bin (DC leakage pattern of the window), rest ~0
    x[7, :] = 0.0
    x[7, 100] = 1.0                              # impulse: flat magnitude spectrum -> thousands of near-equal keys
    t = torch.arange(N, device="cuda", dtype=torch.float32)
    x[8] = torch.sin(2 * np.pi * 1000.0 * t / SR) # one tone
    x[700, 17] = float("nan")
    x[701, 4000] = float("inf")
    x[702, 0] = float("-inf")
    x[703] = 3e38                                # finite input, overflowing spectrum
    for lim in (m.FrequencyLimit.All, m.FrequencyLimit.Range(100.0, 20000.0)):
        r = run_both(ctx_pair, x, lim, m.scaling.divide_by_N_sqrt, window=None)
        st = r.stats_numpy()
        assert st["status"][700] == m.cabi.ERR_NAN_VALUES and st["status"][701] == m.cabi.ERR_INFINITY_VALUES
        assert st["status"][702] == m.cabi.ERR_INFINITY_VALUES and st["status"][703] == m.cabi.ERR_NON_FINITE_RESULT
        ok = st["status"] == 0
        assert ok.sum() == frames - 4


def test_same_bytes_int16_and_unaligned_inputs(ctx_pair):
    import torch
    m = sa()
    g = torch.Generator(device="cuda").manual_seed(5)
    pcm = torch.randint(-32768, 32767, (1 + 900 * 441 + N,), generator=g, device="cuda", dtype=torch.int32).to(torch.int16)
    for off, hop in ((0, 4096), (0, 512), (1, 441), (3, 1001)):
        nf = min(900, (pcm.numel() - off - N) // hop + 1)
        for src in (pcm[off:], pcm[off:].to(torch.float32)):
            run_both(ctx_pair, src.contiguous() if off == 0 else src, m.FrequencyLimit.All, m.scaling.divide_by_N,
                     n=N, frame_stride=hop, n_frames=nf)


def test_same_bytes_padded_rows_and_row_alignments(ctx_pair):
    """out_stride 2049 (rows start at every alignment mod 16 bytes), 2052 (always aligned), 2051; limited ranges whose
    first kept bin is not 0: the aligned-chunk stores of the statistics warp against the scalar stores of the default."""
    import torch
    m = sa()
    x = noise(777, seed=3)
    for lim in (m.FrequencyLimit.All, m.FrequencyLimit.Range(1000.0, 9000.0)):
        nb = m.samples_fft_to_spectrum_batched(x[:1], SR, lim, None).n_bins
        for stride in (nb, nb + 1, nb + 2, nb + 3):
            outs = []
            for ctx in ctx_pair:
                vals = torch.full((777, stride), -5.0, device="cuda")
                r = m.samples_fft_to_spectrum_batched(x, SR, lim, m.scaling.divide_by_N_sqrt, window="hann",
                                                      out_stride=stride, out=vals, ctx=ctx)
                torch.cuda.synchronize()
                outs.append((vals, r.stats.clone()))
            assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
            assert (outs[1][0][:, nb:] == -5.0).all()        # padding untouched
