"""Host-side helpers for the multi-GPU paths (one process per GPU, `torch.distributed`).

Rollout inference shards *trajectories*: contiguous blocks per rank, exactly the
`DistributedSampler(shuffle=False)` pattern the reference's datamodule sets up
(data/datamodule.py:96-104,145-166).  Trajectories never interact, so the data
path has NO collective; the only communication is the timing reduction (max over ranks) and an
optional host-side gather of per-rank statistics, both outside any timed region.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def trajectory_shard(n_total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[start, stop) of the contiguous block of trajectories owned by `rank` (sizes differ by <= 1)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(n_total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def micro_batches(start: int, stop: int, batch: int) -> List[Tuple[int, int]]:
    """Split a shard into rollout micro-batches of at most `batch` trajectories."""
    return [(s, min(s + batch, stop)) for s in range(start, stop, batch)]


def max_over_ranks(value: float, device=None) -> float:
    """Device-timed durations are reported as the max over ranks (never wall clock)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(local: torch.Tensor) -> List[torch.Tensor]:
    """Gather small per-rank statistics (e.g. model calls per trajectory) on every rank (host-side plumbing)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local]
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(dist.get_world_size())]
    dist.all_gather(sizes, torch.tensor([local.numel()], dtype=torch.int64, device=local.device))
    mx = int(max(int(s.item()) for s in sizes))
    pad = torch.zeros(mx, dtype=local.dtype, device=local.device)
    pad[: local.numel()] = local.flatten()
    out = [torch.zeros_like(pad) for _ in sizes]
    dist.all_gather(out, pad)
    return [o[: int(s.item())] for o, s in zip(out, sizes)]
