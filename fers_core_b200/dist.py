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


def gather_columns(local: torch.Tensor, counts: list) -> torch.Tensor:
    """Result tensors are laid out [..., columns] (right-hand side fastest).  Every rank holds the columns of its own
    combinations (``local[..., :counts[rank]]``; trailing padding columns are ignored); returns the tensor with the columns of
    ALL ranks in rank order, on every rank: ONE all_gather over NCCL / NVLink (gloo in the CPU tests)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local[..., : counts[0]].contiguous()
    world, rank = dist.get_world_size(), dist.get_rank()
    width = max(counts)
    lead = tuple(local.shape[:-1])
    # column-major staging: [width, ...] so that a rank's columns are one contiguous block of the gathered buffer
    mine = torch.zeros((width,) + lead, dtype=local.dtype, device=local.device)
    if counts[rank]:
        mine[: counts[rank]] = local[..., : counts[rank]].movedim(-1, 0)
    flat = torch.empty((world * width,) + lead, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(flat, mine)
    allbuf = flat.view((world, width) + lead)
    keep = torch.cat([allbuf[r, : counts[r]] for r in range(world)], dim=0)           # [total columns, ...]
    return keep.movedim(0, -1).contiguous()


def run_linear_sharded(engine, factors: np.ndarray, gather=("U", "R", "f_local"), detail: bool = False) -> dict:
    """The linear hot path for ONE model on all ranks of the default process group: the model and its
    factorisation are replicated, the combinations (rows of ``factors``) are split into contiguous, balanced shards, every
    rank solves and recovers its own columns, and the result tensors named in ``gather`` are collected with one all_gather
    each -- no collective inside the solve.  ``engine`` is this rank's fers_core_b200.engine.Engine of the same model.
    Returns {name: tensor [..., n_combinations]} on every rank, plus "columns" = (lo, hi) of this rank."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    n = factors.shape[0]
    counts = [shard_range(n, world, r)[1] - shard_range(n, world, r)[0] for r in range(world)]
    lo, hi = shard_range(n, world, rank)
    if hi > lo:
        rec = engine.run_linear(factors[lo:hi], detail=detail)
    else:                                   # more ranks than combinations: this rank only takes part in the gathers
        rec = engine.run_linear(factors[:1], detail=detail)
    out = {"columns": (lo, hi)}
    for name in gather:
        if name in rec:
            out[name] = gather_columns(rec[name], counts)
    return out


def max_over_ranks(value: float, device) -> float:
    """Device-timed durations are reported as the maximum over ranks."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
