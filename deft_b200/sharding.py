"""Tree-level data parallelism: which trees a rank owns and how per-rank measurements are combined.

Queries of different decoding trees share no KV, so the path shards by tree with no collective on the
data path: rank r owns a contiguous block of the trees, runs its own launches, and only
timings / checksums cross ranks (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [begin, end) of rank ``rank``; sizes differ by at most one, earlier ranks get the extra."""
    assert 0 <= rank < world and n_items >= 0
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def _active() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def max_over_ranks(value: float, device: torch.device) -> float:
    """Device time of a multi-rank step = the slowest rank's."""
    if not _active():
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: torch.device) -> float:
    """Whole-job totals (units processed, checksums)."""
    if not _active():
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_ranges(n_items: int, world: int) -> List[Tuple[int, int]]:
    return [shard_range(n_items, r, world) for r in range(world)]
