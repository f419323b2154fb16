"""Multi-GPU sharding of the batched path: contiguous frame ranges, no data-path collective.

`samples_fft_to_spectrum` is a pure function of one frame (reference src/lib.rs:157-337), so frames
shard trivially: rank r of W owns frames [r*F/W, (r+1)*F/W); for STFT-style
input (hop < N) a shard additionally reads N - hop samples of halo past its last frame start.
torch.distributed is used only for the barrier and the max-over-ranks timing reduction.
"""
from __future__ import annotations


def frame_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """[first, last) frames of `rank`; ranges are contiguous, disjoint, cover [0, n_frames), and differ
    in size by at most one frame."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_frames, world)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def sample_range(first_frame: int, last_frame: int, n: int, hop: int) -> tuple[int, int]:
    """[first, last) samples a shard must hold: its frames plus the read-only halo of N - hop samples."""
    if last_frame <= first_frame:
        return first_frame * hop, first_frame * hop
    return first_frame * hop, (last_frame - 1) * hop + n


def stft_frames(n_samples: int, n: int, hop: int) -> int:
    """Number of full frames in a stream (cfg3: 1 587 600 000 samples, N=2048, hop 512 -> 3 100 778)."""
    return 0 if n_samples < n else (n_samples - n) // hop + 1


def aggregate_throughput(units_local: int, ms_local: float, group=None) -> tuple[float, float]:
    """(whole-job units/s, max ms) = sum of units over ranks / max time over ranks (device-timed)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return units_local / (ms_local * 1e-3), ms_local
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor([ms_local], dtype=torch.float64, device=dev)
    u = torch.tensor([float(units_local)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(u, op=dist.ReduceOp.SUM, group=group)
    return float(u.item()) / (float(t.item()) * 1e-3), float(t.item())
