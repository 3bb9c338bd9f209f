"""Sharding of independent camera streams over the GPUs of one box.

A stream never exchanges data with another stream and frame t of a stream needs frame t-1's pyramid, so a stream is
pinned to one GPU for its lifetime: stream s -> rank s mod world.  There is no collective on the data path; the only
cross-rank operations are the barrier around the timed region and the max-over-ranks of the elapsed time, which is
what ``aggregate`` restates for the CPU (gloo) test of the multi-rank path."""
from __future__ import annotations

from typing import List, Sequence


def streams_of_rank(n_streams: int, rank: int, world: int) -> List[int]:
    """Streams owned by ``rank`` (round-robin: stream s lives on rank s mod world)."""
    if world < 1 or not 0 <= rank < world or n_streams < 0:
        raise ValueError("bad rank/world/n_streams")
    return list(range(rank, n_streams, world))


def rank_of_stream(stream: int, world: int) -> int:
    return stream % world


def seed_of_stream(stream: int) -> int:
    """Synthetic sequence of stream s uses seed 1000 + s."""
    return 1000 + stream


def aggregate(frames_per_rank: Sequence[int], ms_per_rank: Sequence[float]) -> float:
    """Whole-job frames/s: all frames of all ranks divided by the slowest rank's elapsed time."""
    worst = max(ms_per_rank)
    return sum(frames_per_rank) / (worst * 1e-3) if worst > 0 else 0.0


def distributed_throughput(dist, torch, frames: int, ms: float, device=None) -> float:
    """The reduction bench.py performs under torchrun: SUM of frames, MAX of elapsed ms over the ranks."""
    t = torch.tensor([float(frames), float(ms)], dtype=torch.float64, device=device)
    f, m = t[:1].clone(), t[1:].clone()
    dist.all_reduce(f, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return float(f[0]) / (float(m[0]) * 1e-3)
