"""Timing-experiment build (-DSA_EXP=0x10000): the statistics warp writes cycle counts into the record
(min_bin = wait for full, max_bin = full -> empty, reserved = empty -> end of the previous frame)."""
import os, sys, torch, numpy as np
sys.path.insert(0, os.getcwd())
import spectrum_analyzer_b200 as sa
frames, n = 262144, 4096
x = torch.empty(frames * n, dtype=torch.float32, device="cuda")
ctx = sa.api._torch_ctx(0, torch.cuda.current_stream().cuda_stream)
sa.api._check(sa.cabi.load().sa_generate_noise(ctx._h, x.data_ptr(), 0, x.numel(), 0x5A17))
for _ in range(3):
    r = sa.samples_fft_to_spectrum_batched(x.view(frames, n), 44100, sa.FrequencyLimit.All, sa.scaling.divide_by_N_sqrt, window="hann")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
r = sa.samples_fft_to_spectrum_batched(x.view(frames, n), 44100, sa.FrequencyLimit.All, sa.scaling.divide_by_N_sqrt, window="hann")
e1.record()
torch.cuda.synchronize()
print(f"kernel {e0.elapsed_time(e1):.3f} ms for {frames} frames -> {frames / e0.elapsed_time(e1) / 1e3:.1f} M frames/s")
st = r.stats_numpy()
sel = slice(20000, None)
raw = r.stats.cpu().numpy().view(np.uint32).reshape(-1, 8)
cols = {"min_val": 0, "min_bin": 1, "max_val": 2, "max_bin": 3, "average": 4, "median": 5, "reserved": 7}
for name, key in (("wait full", "min_bin"), ("full->empty", "max_bin"), ("  store+partials+scan", "min_val"), ("  arg bins", "max_val"),
                  ("  pass step 1", "average"), ("  pass step 2", "median"), ("empty->end (prev)", "reserved")):
    v = raw[sel, cols[key]].astype(np.float64)
    print(f"{name:20s} mean {v.mean():8.0f}  p10 {np.percentile(v,10):8.0f}  p50 {np.percentile(v,50):8.0f}  p90 {np.percentile(v,90):8.0f}  p99 {np.percentile(v,99):8.0f}")
