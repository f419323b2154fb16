"""Host-side logic of the N>1 path on CPU: frame-range sharding and the max-over-ranks reduction,
with a real world_size-2 gloo process group (the data path itself has no collective)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
from spectrum_analyzer_b200 import sharding


def test_frame_ranges_partition():
    for n_frames in (0, 1, 7, 8, 1_000_000, 3_100_778):
        for world in (1, 2, 3, 4, 8):
            ranges = [sharding.frame_range(n_frames, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n_frames
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_stft_frame_count_and_halo():
    assert sharding.stft_frames(1_587_600_000, 2048, 512) == 3_100_778      # cfg3
    assert sharding.stft_frames(2047, 2048, 512) == 0
    a, b = sharding.frame_range(3_100_778, 3, 8)
    s0, s1 = sharding.sample_range(a, b, 2048, 512)
    assert s0 == a * 512 and s1 == (b - 1) * 512 + 2048                     # halo = N - hop past the last start
    assert s1 <= 1_587_600_000


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, n, hop, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # each rank regenerates only its own shard of the synthetic stream (counter-based noise)
        f0, f1 = sharding.frame_range(n_frames, rank, world)
        s0, s1 = sharding.sample_range(f0, f1, n, hop)
        x = O.generate_noise(s0, s1 - s0, 0x5A17)
        code, vals, stats, status, _ = O.spectrum_batched(x, f1 - f0, n, hop, 44100, O.LIMIT_ALL, 0, 0,
                                                          O.WINDOW_HAMMING, O.SCALING_NONE, threads=1)
        assert code == 0 and (status == 0).all()
        checksum = torch.tensor([float(np.float64(vals).sum())], dtype=torch.float64)
        dist.all_reduce(checksum)                               # test-only: the product path has no collective
        value, ms = sharding.aggregate_throughput(f1 - f0, 10.0 * (rank + 1))
        dist.barrier()
        out_q.put((rank, f0, f1, float(checksum.item()), value, ms))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shards_equal_single_process():
    n_frames, n, hop, world = 41, 256, 64, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, n, hop, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # shards tile the frame range
    assert results[0][1] == 0 and results[0][2] == results[1][1] and results[1][2] == n_frames
    # the sum of the shard checksums equals the single-process result on the whole stream
    x = O.generate_noise(0, (n_frames - 1) * hop + n, 0x5A17)
    _, vals, _, _, _ = O.spectrum_batched(x, n_frames, n, hop, 44100, O.LIMIT_ALL, 0, 0, O.WINDOW_HAMMING,
                                          O.SCALING_NONE, threads=1)
    want = float(np.float64(vals).sum())
    assert results[0][3] == pytest.approx(want, rel=1e-12) and results[1][3] == pytest.approx(want, rel=1e-12)
    # whole-job throughput = total units / max time over ranks (rank 1 reported 20 ms)
    for r in results:
        assert r[5] == 20.0 and r[4] == pytest.approx(n_frames / 20e-3)
