"""Multi-GPU sharding of the hot path: one process per GPU, no collective on the
data path.  Independent units (load combinations of one model, or frames of a parametric batch)
are split across ranks; only result summaries are gathered with ``torch.distributed``
(NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_units: int, world: int, rank: int) -> tuple:
    """Contiguous, balanced [lo, hi) range of units owned by ``rank`` (first ranks get the extras)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_factors(factors: np.ndarray, world: int, rank: int) -> np.ndarray:
    """Rows (combinations) of the factor matrix owned by ``rank`` (strong-scaling split of one model)."""
    lo, hi = shard_range(factors.shape[0], world, rank)
    return factors[lo:hi]


def gather_summary(local: torch.Tensor, counts: list | None = None) -> torch.Tensor:
    """All-gather a per-unit summary tensor (first axis = units of this rank) from every rank and
    concatenate in rank order.  ``counts`` = units per rank when they differ."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if counts is None:
        counts = [local.shape[0]] * world
    width = max(counts)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


def max_over_ranks(value: float, device) -> float:
    """Device-timed durations are reported as the maximum over ranks."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
