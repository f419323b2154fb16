"""`FrequencySpectrum::apply_scaling_fn` on a RESIDENT BATCH (sa_apply_scaling_chain_batched, called directly through the
C ABI) against the oracle's restatement of src/spectrum.rs:128-168 + scaling::combined (src/scaling.rs:174-185).

Covers what the single-spectrum mirror never reaches: >= 10 000 frames (the grid-stride loop), bin_lo != 0, even and
odd bin counts, padded row strides, stats masks 3 and 7, the three-fold divide_by_N_sqrt of
benches/fft_spectrum_bench.rs:21-36, and ScalingError (src/spectrum.rs:150-160), which IS reachable with the
built-ins: scale_20_times_log10 of an already-dB (negative) value is NaN.

Bar: values, min/max (+ bins), median, status, offender index and class bit-exact; average within 8 + sqrt(m) ulp of
the oracle's (m values summed in a different order)."""
import ctypes as C

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

N_FRAMES = 10_007          # odd, > the kernel's grid (SMs * 8) several times over
SR = 44100


def sa():
    import spectrum_analyzer_b200 as m
    return m


def make_batch(n, limit, amp, seed, stats_mask=7, out_stride=None, scaling=None):
    """A resident batch of GPU spectra: (vals tensor [frames, stride], stats tensor, bin_lo, n_bins)."""
    import torch
    m = sa()
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal(N_FRAMES * n) * amp).astype(np.float32)
    xd = torch.from_numpy(x).cuda().view(N_FRAMES, n)
    r = m.samples_fft_to_spectrum_batched(xd, SR, limit, scaling, window="hann", stats=stats_mask, out_stride=out_stride)
    torch.cuda.synchronize()
    return r


def ulps(a, b):
    a = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7fffffff), a)
    b = np.where(b < 0, -(b & 0x7fffffff), b)
    return np.abs(a - b)


def apply_gpu(r, n, kinds, mask):
    import torch
    m = sa()
    lib = m.cabi.load()
    ctx = m.api._torch_ctx(0, torch.cuda.current_stream().cuda_stream)
    karr = (C.c_int * len(kinds))(*kinds)
    stride = r.vals.shape[1]
    m.api._check(lib.sa_apply_scaling_chain_batched(ctx._h, r.vals.data_ptr(), N_FRAMES, r.n_bins, stride, r.bin_lo, n,
                                                    karr, len(kinds), mask, r.stats.data_ptr()))
    torch.cuda.synchronize()


def oracle_stats_of(vals, bin_lo):
    st = np.zeros(vals.shape[0], O.STATS_DTYPE)
    for f in range(vals.shape[0]):
        s = O.calc_statistics(vals[f], bin_lo)
        st[f] = (s.min_val, s.min_bin, s.max_val, s.max_bin, s.average, s.median)
    return st


def compare(r, want_vals, want_stats, want_status, mask, err_index=None, err_pair=None):
    got = r.vals.cpu().numpy()
    gv = got[:, : r.n_bins]
    gs = r.stats_numpy()
    assert np.array_equal(gs["status"], want_status)
    assert np.array_equal(gv.view(np.uint32), want_vals.view(np.uint32)), "scaled values differ from the oracle's bits"
    ok = want_status == 0
    for k in ("min_val", "min_bin", "max_val", "max_bin"):
        assert np.array_equal(gs[k][ok], want_stats[k][ok]), k
    if mask & 4:
        assert np.array_equal(gs["median"][ok].view(np.uint32), want_stats["median"][ok].view(np.uint32))
    else:
        assert (gs["median"][ok] == 0).all()
    # f32 sums of m values in two different orders: the rounding errors random-walk, ~sqrt(m) ulp apart at most
    assert ulps(gs["average"][ok], want_stats["average"][ok]).max(initial=0) <= 8 + np.sqrt(r.n_bins)
    bad = want_status == O.ERR_SCALING
    if bad.any():
        assert np.array_equal(gs["reserved"][bad] & 0xffff, err_index[bad])
        cls = gs["reserved"][bad] >> 29
        new = err_pair[bad, 1]
        assert np.array_equal(cls == 4, np.isnan(new)) and np.array_equal(cls == 2, new == np.inf) \
            and np.array_equal(cls == 1, new == -np.inf)
        # ScalingError(orig, ..): the offender itself is left as it was
        rows = np.nonzero(bad)[0]
        assert np.array_equal(gv[rows, err_index[bad]].view(np.uint32), err_pair[bad, 0].view(np.uint32))
    if got.shape[1] > r.n_bins:      # padded rows: the padding is never touched
        assert (got[:, r.n_bins:] == PAD).all()


PAD = np.float32(-77.0)


@pytest.mark.parametrize("n,limit,mask,stride_pad", [
    (1024, ("range", 1000.0, 9000.0), 7, 0),      # bin_lo = 24, 186 kept bins (even)
    (1024, ("range", 1000.0, 9040.0), 3, 3),      # bin_lo = 24, 187 kept bins (odd), padded rows, no median
    (2048, ("all",), 7, 0),                       # 1025 bins, the bench's size
    (256, ("min", 5000.0), 7, 1),                 # bin_lo = 30, 99 bins
])
def test_three_fold_divide_by_n_sqrt(gpu_ctx, n, limit, mask, stride_pad):
    """benches/fft_spectrum_bench.rs:21-36 on a batch: divide_by_N_sqrt in the main call, then three more times,
    every call with freshly recomputed statistics; plus the same as ONE combined chain."""
    m = sa()
    lim = {"all": lambda: m.FrequencyLimit.All, "range": lambda a, b: m.FrequencyLimit.Range(a, b),
           "min": lambda a: m.FrequencyLimit.Min(a)}[limit[0]](*limit[1:])
    nb = O.limit_to_bins(n, SR, lim.kind, lim.lo, lim.hi)[1]
    r = make_batch(n, lim, 2000.0, seed=n + mask, out_stride=nb + stride_pad if stride_pad else None,
                   scaling=m.scaling.divide_by_N_sqrt)
    if stride_pad:
        r.vals[:, nb:] = float(PAD)
    vals = r.vals.cpu().numpy()[:, : r.n_bins].copy()
    stats = oracle_stats_of(vals, r.bin_lo)
    status = np.zeros(N_FRAMES, np.uint32)
    for _ in range(3):
        vals, stats, status, ei, ep = O.apply_scaling_chain_batched(vals, r.bin_lo, n, [O.SCALING_DIVIDE_BY_N_SQRT],
                                                                    stats, status)
        apply_gpu(r, n, [O.SCALING_DIVIDE_BY_N_SQRT], mask)
        compare(r, vals, stats, status, mask)
    # the same division three times inside one call (scaling::combined)
    vals, stats, status, ei, ep = O.apply_scaling_chain_batched(vals, r.bin_lo, n, [O.SCALING_DIVIDE_BY_N_SQRT] * 3,
                                                                stats, status)
    apply_gpu(r, n, [O.SCALING_DIVIDE_BY_N_SQRT] * 3, mask)
    compare(r, vals, stats, status, mask)
    assert (status == 0).all()


@pytest.mark.parametrize("mask", [3, 7])
def test_log10_twice_is_a_scaling_error(gpu_ctx, mask):
    """combined([log10, log10]) and log10 applied twice: 20*log10(v) < 0 for v < 1, and log10f of a negative
    number is NaN -> ScalingError at the first such bin; frames whose dB values are all positive pass."""
    m = sa()
    n = 1024
    lim = m.FrequencyLimit.Range(2000.0, 12000.0)
    # amplitude chosen so that most frames contain a magnitude below 1 somewhere and a few hundred do not
    r = make_batch(n, lim, 0.45, seed=5)
    base = r.vals.cpu().numpy()[:, : r.n_bins].copy()
    stats0 = oracle_stats_of(base, r.bin_lo)
    kinds = [O.SCALING_20_TIMES_LOG10, O.SCALING_20_TIMES_LOG10]
    vals, stats, status, ei, ep = O.apply_scaling_chain_batched(base, r.bin_lo, n, kinds, stats0, np.zeros(N_FRAMES, np.uint32))
    n_bad = int((status == O.ERR_SCALING).sum())
    assert 100 < n_bad < N_FRAMES - 100, n_bad       # both outcomes are exercised
    assert np.isnan(ep[status == O.ERR_SCALING, 1]).all()
    apply_gpu(r, n, kinds, mask)
    compare(r, vals, stats, status, mask, ei, ep)

    # the same through two separate calls: the first succeeds everywhere, the second fails where a dB value is negative
    r = make_batch(n, lim, 0.45, seed=5)
    v1, s1, c1, _, _ = O.apply_scaling_chain_batched(base, r.bin_lo, n, kinds[:1], stats0, np.zeros(N_FRAMES, np.uint32))
    apply_gpu(r, n, kinds[:1], mask)
    compare(r, v1, s1, c1, mask)
    assert (c1 == 0).all()
    v2, s2, c2, ei, ep = O.apply_scaling_chain_batched(v1, r.bin_lo, n, kinds[:1], s1, c1)
    apply_gpu(r, n, kinds[:1], mask)
    compare(r, v2, s2, c2, mask, ei, ep)
    # a third call leaves the failed frames alone (no spectrum exists for them) and keeps scaling the others
    v3, s3, c3, ei3, ep3 = O.apply_scaling_chain_batched(v2, r.bin_lo, n, [O.SCALING_DIVIDE_BY_N], s2, c2)
    before = r.stats_numpy()["reserved"].copy()
    apply_gpu(r, n, [O.SCALING_DIVIDE_BY_N], mask)
    got = r.stats_numpy()
    assert np.array_equal(got["status"], c3) and np.array_equal(got["reserved"][c3 != 0], before[c3 != 0])
    assert np.array_equal(r.vals.cpu().numpy()[:, : r.n_bins].view(np.uint32), v3.view(np.uint32))


def test_zero_to_one_does_not_need_minmax_from_the_producer(gpu_ctx):
    """the producing call may have been asked for the median only; scale_to_zero_to_one divides by the
    maximum of the spectrum as it is (src/spectrum.rs:138-145 + :529-530), which the kernel takes from the values."""
    m = sa()
    n = 512
    r = make_batch(n, m.FrequencyLimit.Max(15000.0), 30.0, seed=9, stats_mask=4)
    assert (r.stats_numpy()["max_val"] == 0).all()           # really not computed by the producer
    base = r.vals.cpu().numpy()[:, : r.n_bins].copy()
    stats0 = oracle_stats_of(base, r.bin_lo)
    kinds = [O.SCALING_ZERO_TO_ONE, O.SCALING_20_TIMES_LOG10]
    vals, stats, status, ei, ep = O.apply_scaling_chain_batched(base, r.bin_lo, n, kinds, stats0, np.zeros(N_FRAMES, np.uint32))
    apply_gpu(r, n, kinds, 7)
    compare(r, vals, stats, status, 7, ei, ep)
    assert (status == 0).all() and (vals.max(axis=1) == 0).all()      # 20*log10(1) = 0 at the maximum


def test_python_mirror_raises_scaling_error(gpu_ctx):
    """FrequencySpectrum.apply_scaling_fn -> Err(ScalingError(orig, new)) (src/spectrum.rs:155-160, src/error.rs:50-53)."""
    m = sa()
    x = O.window_apply(O.WINDOW_HANN, (np.random.default_rng(3).standard_normal(2048) * 0.3).astype(np.float32))
    spec = m.samples_fft_to_spectrum(x, SR, m.FrequencyLimit.All, m.scaling.scale_20_times_log10)
    db = spec.values().copy()
    first = int(np.nonzero(db < 0)[0][0])
    with pytest.raises(m.ScalingError) as e:
        spec.apply_scaling_fn(m.scaling.scale_20_times_log10)
    assert e.value.orig == float(db[first]) and np.isnan(e.value.new)
    after = spec.values()
    want_head = np.array([O.scale(O.SCALING_20_TIMES_LOG10, v) for v in db[:first]], np.float32)
    assert np.array_equal(after[:first], want_head) and np.array_equal(after[first:], db[first:])
