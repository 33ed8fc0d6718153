"""Trajectory sharding of the reward path across the GPUs of one node.

The path shards naturally: trajectories are independent
(preconditioner.py:37 loops over them with no cross-iteration state), so rank g
scores a contiguous slice of the batch against its own replica of the context
and the only exchange is an all-gather of the per-trajectory rewards
(8 bytes each). One process per GPU, `torch.distributed` (NCCL on GPUs, gloo in
the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard_rows", "gather_rewards"]


def shard_bounds(batch: int, world: int, rank: int) -> tuple[int, int]:
    """[lo, hi) of the trajectories owned by `rank`: sizes differ by at most one,
    earlier ranks take the extra ones; empty shards are legal (batch < world)."""
    if world <= 0 or not 0 <= rank < world or batch < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rows(actions: torch.Tensor, world: int | None = None, rank: int | None = None) -> torch.Tensor:
    """This rank's rows of the global `actions[B, T]`."""
    world = dist.get_world_size() if world is None else world
    rank = dist.get_rank() if rank is None else rank
    lo, hi = shard_bounds(actions.shape[0], world, rank)
    return actions[lo:hi]


def gather_rewards(local: torch.Tensor, batch: int, group=None) -> torch.Tensor:
    """All-gather the per-shard rewards into the global `[batch]` vector, in
    trajectory order, on every rank. Shards may be ragged (padded to the largest
    shard for the collective, then trimmed)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(batch, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros(width, dtype=local.dtype, device=local.device)
    buf[: local.numel()] = local
    out = torch.empty(world * width, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return torch.cat([out[r * width: r * width + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])


def reward_row_sharded(ctx, actions: torch.Tensor, alpha: float, mode: str = "copy",
                       dtype: torch.dtype = torch.float32, group=None):
    """Secondary partitioning (few trajectories, huge n): every rank evaluates the
    rows [lo, hi) = shard_bounds(n, world, rank) of EVERY trajectory; the partial
    sums of squared row residuals are all-reduced (one f64[B] sum) before the sqrt
    and the mix formula. `actions` is the full CUDA int64 [B, T] batch on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(ctx.n, world, rank)
    res2, nnz = ctx.reward_rows(actions, lo, hi, mode, dtype)
    if world > 1:
        dist.all_reduce(res2, op=dist.ReduceOp.SUM, group=group)
    return ctx.finalize_rewards(res2, nnz, alpha, dtype)
