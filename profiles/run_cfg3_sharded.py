#!/usr/bin/env python
"""cfg3 of BASELINE.json as specified: STFT-style 2048-point Hamming spectra, hop 512, over 10 h of synthetic
44.1 kHz audio (1 587 600 000 samples, 3 100 778 frames), frames sharded over the ranks by contiguous frame ranges
(spectrum_analyzer_b200.sharding.frame_range) with the N - hop = 1536-sample halo.  STRONG scaling: the stream is
fixed, each rank holds only its own samples.

Per run (one JSON line on rank 0):
  * frames/s of the whole job = all frames / max over ranks of the device time (CUDA events, barrier on both sides);
  * seam check: every rank recomputes, in ONE launch that spans the seam (what a single GPU computes there), the last
    K frames of the previous shard and its own first K; they must equal, bit for bit, what the two shards produced
    (values and records; the previous rank's tail arrives through an all_gather of K rows);
  * oracle check on rank 0: a sample of frames (first, last, around the seams) against the CPU oracle, and the
    statistics against the oracle's stable-sort statistics of the GPU's own values;
  * clocks sampled during the timed region.

usage:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
            profiles/run_cfg3_sharded.py [--hours 10] [--steps 3]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N, HOP, SR, SEED, K = 2048, 512, 44100, 0xC0FFEE, 8


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hours", type=float, default=10.0)
    ap.add_argument("--steps", type=int, default=0, help="0: as many as fill ~0.3 s")
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist

    import spectrum_analyzer_b200 as sa
    from bench import ClockSampler
    from spectrum_analyzer_b200 import cabi, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    n_samples = int(args.hours * 3600 * SR)
    n_frames = sharding.stft_frames(n_samples, N, HOP)
    f0, f1 = sharding.frame_range(n_frames, rank, world)
    s0, s1 = sharding.sample_range(f0, f1, N, HOP)
    nf = f1 - f0
    lib = cabi.load()
    ctx = sa.Context(local, torch.cuda.current_stream(dev).cuda_stream)
    chk = sa.api._check
    x = torch.empty(s1 - s0, dtype=torch.float32, device=dev)
    chk(lib.sa_generate_noise(ctx._h, x.data_ptr(), s0, x.numel(), SEED))     # the GLOBAL stream, this rank's slice
    n_bins = N // 2 + 1
    vals = torch.empty((nf, n_bins), dtype=torch.float32, device=dev)
    stats = torch.empty((nf, 32), dtype=torch.uint8, device=dev)

    def step():
        chk(lib.sa_spectrum_batched(ctx._h, x.data_ptr(), nf, N, HOP, SR, cabi.LIMIT_ALL, 0.0, 0.0, cabi.WINDOW_HAMMING,
                                    cabi.SCALING_NONE, 7, vals.data_ptr(), n_bins, stats.data_ptr(), None, None))

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    if args.steps <= 0:     # size the timed region so that the 50 ms clock sampler sees it (same count on every rank)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        args.steps = max(5, int(300.0 / max(float(t.item()), 1e-3)) + 1)
        barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    t_end = time.perf_counter()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_start, t_end)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = n_frames * args.steps / (ms * 1e-3)

    # ---- seam check -------------------------------------------------------------------------------------------
    tail_v = vals[-K:].contiguous()
    tail_s = stats[-K:].contiguous()
    if world > 1:
        tv = [torch.empty_like(tail_v) for _ in range(world)]
        ts = [torch.empty_like(tail_s) for _ in range(world)]
        dist.all_gather(tv, tail_v)
        dist.all_gather(ts, tail_s)
    seams_ok, seam_frames = True, 0
    if rank > 0:
        a, b = f0 - K, f0 + K                                         # frames [a, b) span the seam
        sa0, sa1 = sharding.sample_range(a, b, N, HOP)
        xs = torch.empty(sa1 - sa0, dtype=torch.float32, device=dev)
        chk(lib.sa_generate_noise(ctx._h, xs.data_ptr(), sa0, xs.numel(), SEED))
        sv = torch.empty((2 * K, n_bins), dtype=torch.float32, device=dev)
        ss = torch.empty((2 * K, 32), dtype=torch.uint8, device=dev)
        chk(lib.sa_spectrum_batched(ctx._h, xs.data_ptr(), 2 * K, N, HOP, SR, cabi.LIMIT_ALL, 0.0, 0.0, cabi.WINDOW_HAMMING,
                                    cabi.SCALING_NONE, 7, sv.data_ptr(), n_bins, ss.data_ptr(), None, None))
        torch.cuda.synchronize(dev)
        seams_ok = (torch.equal(sv[:K].view(torch.int32), tv[rank - 1].view(torch.int32)) and torch.equal(ss[:K], ts[rank - 1])
                    and torch.equal(sv[K:].view(torch.int32), vals[:K].view(torch.int32)) and torch.equal(ss[K:], stats[:K]))
        seam_frames = 2 * K
    if world > 1:
        flag = torch.tensor([1 if seams_ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        seams_ok = bool(flag.item())
        cnt = torch.tensor([seam_frames], device=dev)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        seam_frames = int(cnt.item())

    # ---- oracle check (rank 0's shard: first frames, last frames) ------------------------------------------------
    oracle = None
    if rank == 0:
        import oracle as O
        idx = sorted(set(list(range(4)) + [nf // 2] + list(range(nf - 4, nf))))
        st = stats.cpu().numpy().view(cabi.FRAME_STATS_DTYPE).reshape(-1)
        worst = 0.0
        for f in idx:
            frame = x[f * HOP: f * HOP + N].cpu().numpy()
            ref = O.samples_fft_to_spectrum(O.window_apply(O.WINDOW_HAMMING, frame), SR)
            got = vals[f].cpu().numpy()
            worst = max(worst, float(np.abs(got - ref.vals).max() / ref.vals.max()))
            s = O.calc_statistics(got, 0)
            assert (s.min_val, s.min_bin, s.max_val, s.max_bin, s.median) == \
                (st["min_val"][f], st["min_bin"][f], st["max_val"][f], st["max_bin"][f], st["median"][f]), f
        assert worst <= 1e-4, worst
        assert int((st["status"] != 0).sum()) == 0
        oracle = {"frames_checked": len(idx), "max_abs_err_over_frame_max": worst, "tolerance": 1e-4}

    if rank == 0:
        unique_bytes = 4 * HOP + 4 * n_bins + 32
        try:
            peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            peak = 6650.0
        print(json.dumps({
            "config": "cfg3: 2048-point Hamming, hop 512, %.2f h of 44.1 kHz audio, FrequencyLimit::All, no scaling, all stats" % args.hours,
            "n_gpus": world, "scaling": "strong", "n_samples": n_samples, "n_frames": n_frames, "frames_per_rank": nf,
            "halo_samples": N - HOP, "steps": args.steps, "ms_per_step": ms / args.steps, "frames_per_s": value,
            "msamples_per_s": value * HOP / 1e6, "kernel": ctx.last_kernel,
            "hbm_frac_unique_bytes": value * unique_bytes / 1e9 / (peak * world), "algorithmic_bytes_per_frame": unique_bytes,
            "seams_bit_identical": seams_ok, "seam_frames_checked_all_ranks": seam_frames, "oracle": oracle, "clocks": clocks}))
        assert seams_ok, "frames at a shard seam differ from the single-launch result"
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
