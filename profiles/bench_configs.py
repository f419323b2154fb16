#!/usr/bin/env python
"""Secondary measurements: device-resident throughput of every BASELINE.json config (the headline
line stays bench.py / configs[1]).  One JSON line per config: frames/s, Msamples/s and fraction of the
measured HBM roofline with SURVEY.md 8(d)'s algorithmic bytes per frame (4*hop + 4*n_bins + 32).

usage (on a B200): python profiles/bench_configs.py [--scale 1.0] > gpurun_out/configs.jsonl
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="scale the frame counts")
    ap.add_argument("--steps", type=int, default=0, help="timed launches per config (0: as many as fill ~0.3 s)")
    ap.add_argument("--only", default="", help="substring filter on the config name")
    args = ap.parse_args()
    import time
    from bench import ClockSampler
    import torch
    import spectrum_analyzer_b200 as sa
    from spectrum_analyzer_b200 import cabi
    sa.build()
    lib = cabi.load()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream(dev)
    ctx = sa.Context(0, stream.cuda_stream)
    chk = sa.api._check
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    L = sa.FrequencyLimit
    configs = [
        # name, n, hop, frames, window, limit, scaling, stats
        ("cfg2 4096 hann all div_sqrt", 4096, 4096, 1_000_000, cabi.WINDOW_HANN, L.All, cabi.SCALING_DIVIDE_BY_N_SQRT, 7),
        ("cfg3 2048 hamming hop512 (2.5h audio)", 2048, 512, 775_000, cabi.WINDOW_HAMMING, L.All, cabi.SCALING_NONE, 7),
        ("cfg4 16384 bh4 range(50,15000) zero_to_one minmax", 16384, 16384, 262_144, cabi.WINDOW_BLACKMAN_HARRIS_4TERM,
         L.Range(50.0, 15000.0), cabi.SCALING_ZERO_TO_ONE, 3),
        ("cfg5 256 log10 median (16M frames)", 256, 256, 16 * 1024 * 1024, cabi.WINDOW_NONE, L.All,
         cabi.SCALING_20_TIMES_LOG10, 7),
        ("1024 hann all div_n", 1024, 1024, 4_000_000, cabi.WINDOW_HANN, L.All, cabi.SCALING_DIVIDE_BY_N, 7),
        ("2048 hann all div_sqrt", 2048, 2048, 2_000_000, cabi.WINDOW_HANN, L.All, cabi.SCALING_DIVIDE_BY_N_SQRT, 7),
        ("8192 hann all div_n", 8192, 8192, 500_000, cabi.WINDOW_HANN, L.All, cabi.SCALING_DIVIDE_BY_N, 7),
        ("64 hann all none", 64, 64, 32 * 1024 * 1024, cabi.WINDOW_HANN, L.All, cabi.SCALING_NONE, 7),
        ("512 hann all none", 512, 512, 8 * 1024 * 1024, cabi.WINDOW_HANN, L.All, cabi.SCALING_NONE, 7),
        ("32768 hann all none", 32768, 32768, 65_536, cabi.WINDOW_HANN, L.All, cabi.SCALING_NONE, 7),
        ("2048 hamming hop441 (odd hop: frames on 4-byte boundaries)", 2048, 441, 775_000, cabi.WINDOW_HAMMING, L.All, cabi.SCALING_NONE, 7),
        ("fastlog cfg5 256 log10 median (sa_ctx_set_fast_log10)", 256, 256, 16 * 1024 * 1024, cabi.WINDOW_NONE, L.All,
         cabi.SCALING_20_TIMES_LOG10, 7),
        ("4096 hann all log10", 4096, 4096, 1_000_000, cabi.WINDOW_HANN, L.All, cabi.SCALING_20_TIMES_LOG10, 7),
        ("fastlog 4096 hann all log10 (sa_ctx_set_fast_log10)", 4096, 4096, 1_000_000, cabi.WINDOW_HANN, L.All,
         cabi.SCALING_20_TIMES_LOG10, 7),
        ("4096 hann max(4000) div_sqrt (372 of 2049 bins kept)", 4096, 4096, 1_000_000, cabi.WINDOW_HANN, L.Max(4000.0),
         cabi.SCALING_DIVIDE_BY_N_SQRT, 7),
        ("4096 hann range(50,12000) div_sqrt (1111 bins kept)", 4096, 4096, 1_000_000, cabi.WINDOW_HANN, L.Range(50.0, 12000.0),
         cabi.SCALING_DIVIDE_BY_N_SQRT, 7),
        ("i16 cfg2 4096 hann all div_sqrt", 4096, 4096, 1_000_000, cabi.WINDOW_HANN, L.All, cabi.SCALING_DIVIDE_BY_N_SQRT, 7),
        ("i16 cfg3 2048 hamming hop512", 2048, 512, 775_000, cabi.WINDOW_HAMMING, L.All, cabi.SCALING_NONE, 7),
    ]
    for name, n, hop, frames, win, lim, sc, st in configs:
        if args.only and args.only not in name:
            continue
        i16 = name.startswith("i16")
        chk(lib.sa_ctx_set_fast_log10(ctx._h, 1 if name.startswith("fastlog") else 0))
        frames = max(1024, int(frames * args.scale))
        nsamp = (frames - 1) * hop + n
        x = torch.empty(nsamp, dtype=torch.float32, device=dev)
        chk(lib.sa_generate_noise(ctx._h, x.data_ptr(), 0, nsamp, 0x5A17))
        if i16:
            x = (x * 8192.0).clamp_(-32768, 32767).to(torch.int16)
        entry = lib.sa_spectrum_batched_i16 if i16 else lib.sa_spectrum_batched
        esz = 2 if i16 else 4
        bin_lo, n_bins = lim.to_bins(n, 44100)
        vals = torch.empty(frames * n_bins, dtype=torch.float32, device=dev)
        stats = torch.empty(frames * 32, dtype=torch.uint8, device=dev)

        def step():
            chk(entry(ctx._h, x.data_ptr(), frames, n, hop, 44100, lim.kind, lim.lo, lim.hi, win, sc, st,
                                        vals.data_ptr(), n_bins, stats.data_ptr(), None, None))
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        # size the timed region (>= 0.3 s) so that the 50 ms clock sampler sees it
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        torch.cuda.synchronize()
        steps = args.steps or max(5, int(300.0 / max(e0.elapsed_time(e1), 1e-3)) + 1)
        sampler = ClockSampler(0)
        sampler.start()
        time.sleep(0.12)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        clocks = sampler.stop(t0, t1)
        ms = e0.elapsed_time(e1) / steps
        bpf = esz * hop + 4 * n_bins + 32
        fps = frames / (ms * 1e-3)
        ok = int((stats.view(torch.int32).view(frames, 8)[:, 6] != 0).sum().item()) == 0
        print(json.dumps({"config": name, "n": n, "hop": hop, "frames": frames, "n_bins": n_bins, "ms": round(ms, 4),
                          "frames_per_s": fps, "msamples_per_s": fps * hop / 1e6, "algorithmic_bytes_per_frame": bpf,
                          "achieved_gbs": bpf * fps / 1e9, "hbm_frac_of_measured": bpf * fps / 1e9 / peak,
                          # roofline ridge of the path at this N (north_star): ~5 N log2 N flop per frame over the
                          # algorithmic bytes; B200 fp32 ridge = 75 TFLOP/s (non-tensor) / measured HBM = ~11.5 flop/B
                          "flop_per_byte": 2.5 * n * (n.bit_length() - 1) / bpf,
                          "fp32_tflops_algorithmic": 2.5 * n * (n.bit_length() - 1) * fps / 1e12,
                          "kernel": ctx.last_kernel, "steps": steps, "clocks": clocks,
                          "all_status_ok": ok}), flush=True)
        del x, vals, stats
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
