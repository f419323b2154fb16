#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json configs[1]).

Workload (config.workload): batched 4096-point Hann-windowed spectra, 1 000 000 frames of synthetic
f32 noise per GPU, FrequencyLimit::All, scaling::divide_by_N_sqrt, all statistics
(min/max/average/median), one fused kernel launch per step.

  python bench.py --gpus N --steps K --warmup W            # our CUDA path
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

A "step" = one pass of the hot path over the whole batch.  `value` = whole-job spectra/s with the
inputs resident in HBM; `e2e` = the same metric through the host-buffer C-ABI entry point
(sa_spectrum_batched_host) with H2D/D2H inside the timed region.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N = 4096
SAMPLING_RATE = 44100
FRAMES_DEFAULT = 1_000_000
N_BINS = N // 2 + 1
ALGO_BYTES_PER_FRAME = 4 * N + 4 * N_BINS + 32          # unique input + kept bins + 32 B record
ALGO_FLOPS_PER_FRAME = 2.5 * N * 12                      # 2.5 * N * log2 N (reporting only)
SEED = 0x5A17
METRIC = "spectra_per_sec_4096pt_hann_batch"
UNIT = "spectra/s"


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            inside = t0 <= ts <= t1 + 0.06
            try:
                if inside:
                    sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            if inside:
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     parts[3:7]):
                    if val == "Active":
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def host_threads() -> int:
    """Host cores this process may use (torchrun exports OMP_NUM_THREADS=1, so do not ask OpenMP)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline_run(frames: int, threads: int, native: bool = True):
    """Times the oracle port (restatement of the reference's CPU path: per-frame hann_window +
    samples_fft_to_spectrum + divide_by_N_sqrt, frames split statically over `threads`)."""
    import numpy as np
    import oracle as O
    x = O.generate_noise(0, frames * N, SEED)
    t0 = time.perf_counter()
    code, vals, stats, status, _ = O.spectrum_batched(x, frames, N, N, SAMPLING_RATE, O.LIMIT_ALL, 0, 0, O.WINDOW_HANN,
                                                      O.SCALING_DIVIDE_BY_N_SQRT, threads=threads, native=native)
    dt = time.perf_counter() - t0
    assert code == 0 and (status == 0).all()
    return dt, float(np.float64(vals[:, 1:8].sum()))


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The Rust crate cannot be
    built here (no cargo/rustc, microfft/libm crates absent), so this is the oracle PORT (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle as O
    threads = host_threads()
    O.lib(native=True)
    dt, _ = cpu_baseline_run(512 * max(1, threads // 8), threads)               # calibration (untimed)
    rate = 512 * max(1, threads // 8) / dt
    frames = int(min(max(rate * 2.0, 1024), 262144))                             # ~2 s of CPU work per step
    for _ in range(args.warmup):
        cpu_baseline_run(frames, threads)
    t = 0.0
    for _ in range(args.steps):
        d, _ = cpu_baseline_run(frames, threads)
        t += d
    value = frames * args.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "msamples_per_s": value * N / 1e6,
        "config": {"workload": "batched 4096-point Hann spectra, synthetic f32 noise, divide_by_N_sqrt, "
                               "FrequencyLimit::All (BASELINE.json configs[1])",
                   "frames_per_step": frames, "n": N, "note": "bounded sample of the 1M-frame workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{frames} frames x {N} samples per step, {args.steps} steps, gcc -O3 -march=native, "
                                   f"OpenMP static over frames ({threads} threads)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist

    import spectrum_analyzer_b200 as sa
    from spectrum_analyzer_b200 import cabi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback exists)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sa.build()
    lib = cabi.load()
    frames = args.frames
    stream = torch.cuda.current_stream(dev)
    ctx = sa.Context(local_rank, stream.cuda_stream)
    chk = sa.api._check

    # ---- synthetic input, generated on the device (each rank its own frame range of the stream) ----
    x = torch.empty(frames * N, dtype=torch.float32, device=dev)
    chk(lib.sa_generate_noise(ctx._h, x.data_ptr(), rank * frames * N, x.numel(), SEED))
    out_stride = args.out_stride or N_BINS
    vals = torch.empty(frames * out_stride, dtype=torch.float32, device=dev)
    stats = torch.empty(frames * 32, dtype=torch.uint8, device=dev)
    stats_mask = args.stats

    def step():
        chk(lib.sa_spectrum_batched(ctx._h, x.data_ptr(), frames, N, N, SAMPLING_RATE, cabi.LIMIT_ALL, 0.0, 0.0,
                                    cabi.WINDOW_HANN, cabi.SCALING_DIVIDE_BY_N_SQRT, stats_mask, vals.data_ptr(),
                                    out_stride, stats.data_ptr(), None, None))

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    t_end = time.perf_counter()
    ms_local = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop(t_start, t_end) if rank == 0 else None
    ms_total = max_over_ranks(ms_local)
    ms_per_step = ms_total / args.steps
    value = world * frames * args.steps / (ms_total * 1e-3)

    # sanity on the results of the timed region (all frames OK, stats ordered)
    st = stats.view(torch.int32).view(frames, 8)
    if os.environ.get("SA_LIB_PATH"):
        stats_mask_checked = 0   # timing-experiment build of the library (profiles/microbench/): results not meaningful
    else:
        stats_mask_checked = stats_mask
    if stats_mask_checked:
        assert int((st[:, 6] != 0).sum().item()) == 0, "non-OK frame status"
    stf = stats.view(torch.float32).view(frames, 8)
    if stats_mask_checked == 7:
        assert bool(((stf[:, 0] <= stf[:, 5]) & (stf[:, 5] <= stf[:, 2])).all().item()), "min <= median <= max violated"

    # ---- roofline of the dominant (only) kernel: one launch per step -------------------------------
    peak, peak_src = measured_hbm_peak()
    kernel_ms = ms_local / max(launches, 1)
    achieved = ALGO_BYTES_PER_FRAME * frames / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel": ctx.last_kernel + " (one launch per step)",
                "algorithmic_bytes_per_frame": ALGO_BYTES_PER_FRAME, "kernel_ms": kernel_ms,
                "fp32_tflops_algorithmic": ALGO_FLOPS_PER_FRAME * frames / (kernel_ms * 1e-3) / 1e12}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                t = json.load(f)
            roofline["traffic"] = t.get("dram_bytes_per_frame") and t["dram_bytes_per_frame"] * frames
            roofline["traffic_source"] = t.get("source")
        except Exception:
            pass

    # ---- e2e: host buffers through sa_spectrum_batched_host (H2D + kernel + D2H per step) -----------
    e2e = None
    if not args.no_e2e:
        ef = args.e2e_frames or frames
        hx = hv = hs = None
        while ef >= 1024:
            try:
                hx = torch.empty(ef * N, dtype=torch.float32, pin_memory=True)
                hv = torch.empty(ef * out_stride, dtype=torch.float32, pin_memory=True)
                hs = torch.empty(ef * 32, dtype=torch.uint8, pin_memory=True)
                break
            except RuntimeError:
                hx = hv = hs = None
                ef //= 2
        if hx is not None:
            hx.copy_(x[: ef * N])
            torch.cuda.synchronize(dev)
            ectx = sa.Context(local_rank)   # own streams; the call is synchronous

            def e2e_step():
                chk(lib.sa_spectrum_batched_host(ectx._h, hx.data_ptr(), ef, N, N, SAMPLING_RATE, cabi.LIMIT_ALL, 0.0,
                                                 0.0, cabi.WINDOW_HANN, cabi.SCALING_DIVIDE_BY_N_SQRT, stats_mask,
                                                 hv.data_ptr(), out_stride, hs.data_ptr(), None, None))

            e2e_step()  # warm-up (allocates the staging slots)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                e2e_step()
            torch.cuda.synchronize(dev)
            dt = max_over_ranks(time.perf_counter() - t0)
            barrier()
            assert torch.equal(hv[: 64 * out_stride], vals[: 64 * out_stride].cpu()), "e2e output != device output"
            e2e = {"value": world * ef * args.e2e_steps / dt, "unit": UNIT,
                   "h2d_bytes_per_step": ef * N * 4, "d2h_bytes_per_step": ef * (out_stride * 4 + 32),
                   "frames_per_step": ef, "steps": args.e2e_steps, "ms_per_step": 1e3 * dt / args.e2e_steps,
                   "api": "sa_spectrum_batched_host (pinned host buffers, 3-slot H2D/compute/D2H pipeline)"}
            # copy-only ceiling of this box at this N: the same bytes over the same pinned buffers, H2D and D2H
            # concurrently on two streams, no kernel -- what a perfect pipeline could reach (all ranks at once)
            try:
                s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
                dvals = vals[: ef * out_stride]
                dst = torch.empty(ef * N, dtype=torch.float32, device=dev)

                def copy_step():
                    with torch.cuda.stream(s_in):
                        dst.copy_(hx, non_blocking=True)
                    with torch.cuda.stream(s_out):
                        hv.copy_(dvals, non_blocking=True)
                        hs.copy_(stats[: ef * 32], non_blocking=True)
                copy_step()
                barrier()
                t0 = time.perf_counter()
                for _ in range(args.e2e_steps):
                    copy_step()
                torch.cuda.synchronize(dev)
                dtc = max_over_ranks(time.perf_counter() - t0)
                barrier()
                ceil = world * ef * args.e2e_steps / dtc
                e2e["copy_ceiling"] = {"value": ceil, "unit": UNIT,
                                       "h2d_gbs_per_gpu": ef * N * 4 * args.e2e_steps / dtc / 1e9,
                                       "d2h_gbs_per_gpu": ef * (out_stride * 4 + 32) * args.e2e_steps / dtc / 1e9,
                                       "how": "cudaMemcpyAsync of the step's input (H2D) and output (D2H) on two streams, "
                                              "concurrently on all ranks, no kernel"}
                e2e["frac_of_copy_ceiling"] = e2e["value"] / ceil
                del dst
            except RuntimeError as ex:
                e2e["copy_ceiling"] = {"error": str(ex)[:200]}
            # beside it (not the headline): the same frames as int16 PCM through sa_spectrum_batched_host_i16 --
            # the host->device copy, the limiter of the f32 call, is halved
            try:
                hx16 = torch.empty(ef * N, dtype=torch.int16, pin_memory=True)
                hx16.copy_((hx * 8192.0).clamp_(-32768, 32767).to(torch.int16))

                def e2e_step16():
                    chk(lib.sa_spectrum_batched_host_i16(ectx._h, hx16.data_ptr(), ef, N, N, SAMPLING_RATE, cabi.LIMIT_ALL,
                                                         0.0, 0.0, cabi.WINDOW_HANN, cabi.SCALING_DIVIDE_BY_N_SQRT,
                                                         stats_mask, hv.data_ptr(), out_stride, hs.data_ptr(), None, None))
                e2e_step16()
                barrier()
                t0 = time.perf_counter()
                for _ in range(args.e2e_steps):
                    e2e_step16()
                torch.cuda.synchronize(dev)
                dt16 = max_over_ranks(time.perf_counter() - t0)
                barrier()
                e2e["int16_pcm_input"] = {"value": world * ef * args.e2e_steps / dt16, "unit": UNIT,
                                          "h2d_bytes_per_step": ef * N * 2, "d2h_bytes_per_step": ef * (out_stride * 4 + 32),
                                          "api": "sa_spectrum_batched_host_i16"}
                del hx16
            except RuntimeError:
                pass
            # the same call from PAGEABLE host memory (what a caller holding a plain Vec<f32> / numpy array pays):
            # a quarter of the frames, one step after one warm-up
            try:
                import numpy as np
                pf = max(1024, ef // 4)
                px = np.empty(pf * N, dtype=np.float32)
                px[:] = hx[: pf * N].numpy()
                pv = np.empty(pf * out_stride, dtype=np.float32)
                ps = np.empty(pf * 32, dtype=np.uint8)

                def e2e_step_pageable():
                    chk(lib.sa_spectrum_batched_host(ectx._h, px.ctypes.data, pf, N, N, SAMPLING_RATE, cabi.LIMIT_ALL, 0.0,
                                                     0.0, cabi.WINDOW_HANN, cabi.SCALING_DIVIDE_BY_N_SQRT, stats_mask,
                                                     pv.ctypes.data, out_stride, ps.ctypes.data, None, None))
                e2e_step_pageable()
                barrier()
                t0 = time.perf_counter()
                e2e_step_pageable()
                dtp = max_over_ranks(time.perf_counter() - t0)
                barrier()
                assert np.array_equal(pv[: 64 * out_stride], vals[: 64 * out_stride].cpu().numpy()), "pageable e2e output != device output"
                e2e["pageable_input"] = {"value": world * pf / dtp, "unit": UNIT, "frames_per_step": pf,
                                         "h2d_bytes_per_step": pf * N * 4, "d2h_bytes_per_step": pf * (out_stride * 4 + 32),
                                         "api": "sa_spectrum_batched_host from pageable numpy arrays"}
                del px, pv, ps
            except (RuntimeError, MemoryError) as ex:
                e2e["pageable_input"] = {"error": str(ex)[:200]}
            del hx, hv, hs

    # ---- CPU baseline beside it (rank 0, N=1 only) ---------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle as O
        threads = host_threads()
        O.lib(native=True)
        d0, _ = cpu_baseline_run(256 * max(1, threads // 8), threads)
        rate = 256 * max(1, threads // 8) / d0
        cf = int(min(max(rate * 10.0, 2048), 1 << 20))                    # ~10 s of CPU work
        d, _ = cpu_baseline_run(cf, threads)
        cpu = {"value": cf / d, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{cf} frames x {N} samples of the same noise stream, oracle port (gcc -O3 -march=native), "
                         f"OpenMP static over frames, {d:.1f} s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "msamples_per_s": value * N / 1e6,
            "config": {"workload": "batched 4096-point Hann spectra, synthetic f32 noise, divide_by_N_sqrt, "
                                   "FrequencyLimit::All (BASELINE.json configs[1])",
                       "frames_per_gpu": frames, "n": N, "n_bins": N_BINS, "out_stride": out_stride,
                       "stats_mask": stats_mask, "sharding": f"contiguous frame ranges, {world} rank(s), no collective",
                       "host_numa_node_rank0": numa,
                       "l2": "inputs (16.4 GB/GPU) and outputs (8.2 GB/GPU) are far larger than the 126 MB L2"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_JSON_OUT = None


def bind_to_gpu_numa_node(torch, local_rank: int):
    """Best effort, multi-rank runs only: run this rank on the CPUs next to its GPU so that the pinned host
    buffers of the end-to-end measurement are allocated on that NUMA node (first touch) and the PCIe traffic
    does not cross sockets.  Returns the node or None."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(base + "/numa_node").read())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if node < 0 or not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def emit(line: dict) -> None:
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES_DEFAULT, help="frames per GPU (default: the 1M-frame workload)")
    ap.add_argument("--stats", type=int, default=7, help="stats mask (7 = min/max + average + median)")
    ap.add_argument("--out-stride", type=int, default=0)
    ap.add_argument("--e2e-frames", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    # stdout carries the one JSON line and nothing else: libraries that printf to fd 1 (NCCL's version banner,
    # OpenMP runtimes) are sent to stderr, the line itself goes to a private duplicate of the original stdout
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
