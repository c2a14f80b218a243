"""Stream sharding across ranks: independent streams, block partition, no collective
on the data path.  Pure host logic (no CUDA), shared by bench.py and the gloo tests."""
from __future__ import annotations

from typing import Tuple


def stream_range(rank: int, world: int, n_streams_total: int) -> Tuple[int, int]:
    """Rank `rank` of `world` owns streams [lo, hi): contiguous blocks, sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    lo = n_streams_total * rank // world
    hi = n_streams_total * (rank + 1) // world
    return lo, hi


def weak_scaling_streams(rank: int, world: int, streams_per_rank: int) -> Tuple[int, int]:
    """Weak scaling: every rank gets `streams_per_rank` streams; global ids are contiguous per rank."""
    return rank * streams_per_rank, (rank + 1) * streams_per_rank


def aggregate_throughput(samples_per_rank: int, world: int, max_elapsed_s: float) -> float:
    """Whole-job samples/s: all ranks' units over the slowest rank's time."""
    return samples_per_rank * world / max_elapsed_s
