#!/usr/bin/env python
"""The deployment shape BASELINE.json's north_star describes: ONE process, one host thread + one sa_ctx + one CUDA
stream per GPU, frames sharded by contiguous ranges, no collective (bench.py measures the same work with one PROCESS per
GPU under torchrun, as the bench contract requires).  The C-ABI calls are made through ctypes, which releases the GIL,
so the threads run concurrently; a context is used by exactly one thread.

Prints one JSON line: device-resident spectra/s (all GPUs, max over the threads' CUDA-event times) and the end-to-end
rate through sa_spectrum_batched_host from pinned buffers, for cfg2 (4096-point Hann, divide_by_N_sqrt, all stats).

usage: python profiles/threads_one_process.py --gpus 8 [--frames 1000000] [--steps 10] [--e2e-frames 250000]
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N, SR, SEED = 4096, 44100, 0x5A17


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--frames", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--e2e-frames", type=int, default=250_000)
    ap.add_argument("--e2e-steps", type=int, default=2)
    args = ap.parse_args()
    import torch

    import spectrum_analyzer_b200 as sa
    from spectrum_analyzer_b200 import cabi
    g = args.gpus or torch.cuda.device_count()
    lib = cabi.load()
    chk = sa.api._check
    n_bins = N // 2 + 1
    bar = threading.Barrier(g)
    res = [None] * g
    err = []

    def worker(i):
        try:
            torch.cuda.set_device(i)
            dev = torch.device("cuda", i)
            stream = torch.cuda.Stream(dev)
            with torch.cuda.stream(stream):
                ctx = sa.Context(i, stream.cuda_stream)
                x = torch.empty(args.frames * N, dtype=torch.float32, device=dev)
                chk(lib.sa_generate_noise(ctx._h, x.data_ptr(), i * args.frames * N, x.numel(), SEED))
                vals = torch.empty(args.frames * n_bins, dtype=torch.float32, device=dev)
                stats = torch.empty(args.frames * 32, dtype=torch.uint8, device=dev)

                def step():
                    chk(lib.sa_spectrum_batched(ctx._h, x.data_ptr(), args.frames, N, N, SR, cabi.LIMIT_ALL, 0.0, 0.0,
                                                cabi.WINDOW_HANN, cabi.SCALING_DIVIDE_BY_N_SQRT, 7, vals.data_ptr(), n_bins,
                                                stats.data_ptr(), None, None))
                for _ in range(3):
                    step()
                stream.synchronize()
                bar.wait()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter()
                e0.record(stream)
                for _ in range(args.steps):
                    step()
                e1.record(stream)
                stream.synchronize()
                t1 = time.perf_counter()
                bar.wait()
                ms = e0.elapsed_time(e1)
                ok = int((stats.view(torch.int32).view(-1, 8)[:, 6] != 0).sum().item()) == 0
                # end to end from pinned host buffers, own context (own three streams)
                ef = args.e2e_frames
                hx = torch.empty(ef * N, dtype=torch.float32, pin_memory=True)
                hv = torch.empty(ef * n_bins, dtype=torch.float32, pin_memory=True)
                hs = torch.empty(ef * 32, dtype=torch.uint8, pin_memory=True)
                hx.copy_(x[: ef * N])
                stream.synchronize()
                ectx = sa.Context(i)

                def estep():
                    chk(lib.sa_spectrum_batched_host(ectx._h, hx.data_ptr(), ef, N, N, SR, cabi.LIMIT_ALL, 0.0, 0.0,
                                                     cabi.WINDOW_HANN, cabi.SCALING_DIVIDE_BY_N_SQRT, 7, hv.data_ptr(), n_bins,
                                                     hs.data_ptr(), None, None))
                estep()
                bar.wait()
                t2 = time.perf_counter()
                for _ in range(args.e2e_steps):
                    estep()
                t3 = time.perf_counter()
                bar.wait()
                same = torch.equal(hv[: 64 * n_bins], vals[: 64 * n_bins].cpu())
                res[i] = {"ms": ms, "wall": t1 - t0, "e2e_s": t3 - t2, "ok": ok and same, "kernel": ctx.last_kernel}
        except Exception as ex:      # a failed thread must not leave the others at the barrier
            err.append(repr(ex))
            bar.abort()

    th = [threading.Thread(target=worker, args=(i,)) for i in range(g)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if err:
        print(json.dumps({"error": err}))
        return 1
    ms = max(r["ms"] for r in res)
    e2e_s = max(r["e2e_s"] for r in res)
    print(json.dumps({
        "shape": "one process, %d host threads, one sa_ctx + stream per GPU" % g, "n_gpus": g, "config": "cfg2",
        "frames_per_gpu": args.frames, "steps": args.steps, "value": g * args.frames * args.steps / (ms * 1e-3),
        "unit": "spectra/s", "ms_per_step_max_over_threads": ms / args.steps,
        "wall_s_max_over_threads": max(r["wall"] for r in res),
        "e2e": {"value": g * args.e2e_frames * args.e2e_steps / e2e_s, "unit": "spectra/s", "frames_per_gpu": args.e2e_frames,
                "api": "sa_spectrum_batched_host, pinned buffers, one call per thread"},
        "results_ok": all(r["ok"] for r in res), "kernel": res[0]["kernel"]}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
