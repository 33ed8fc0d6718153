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

__all__ = ["shard_bounds", "shard_rows", "gather_rewards", "reward_row_sharded", "dp_trajectory_balance_loss",
           "allreduce_gradients", "data_parallel_step"]


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


# --------------------------------------------------------------------------
# Data-parallel training step (the north star's last collective): trajectories
# are sharded, the trajectory-balance loss is formed shard-locally and only
# scalars + the policy gradients cross NVLink.
# --------------------------------------------------------------------------
def _all_reduce(t: torch.Tensor, op, group=None, async_op: bool = False):
    """all_reduce that also serves host tensors under NCCL (staged through the current CUDA device).
    Returns (work or None, tensor holding the result once the work is done, original)."""
    if dist.get_backend(group) == "nccl" and not t.is_cuda:
        staged = t.to(torch.device("cuda", torch.cuda.current_device()))
        work = dist.all_reduce(staged, op=op, group=group, async_op=async_op)
        if not async_op:
            t.copy_(staged)
        return work, staged, t
    work = dist.all_reduce(t, op=op, group=group, async_op=async_op)
    return work, t, t


def dp_trajectory_balance_loss(total_flow, rewards, fwd_probs, back_probs, global_batch: int, group=None):
    """Shard-local surrogate of the reference's mean trajectory-balance loss
    (gflownet/utils.py:228-278) over the GLOBAL batch.

    Returns (surrogate, loss_value): summing the gradients of `surrogate` over the ranks
    (all-reduce, `allreduce_gradients`) gives exactly the gradient of the single-process loss on
    the concatenated batch, and `loss_value` (a detached scalar, identical on every rank) is that
    loss. The reference subtracts the BATCH maximum of the summed log-probabilities on each side
    (utils.py:262-267) and that maximum carries gradient; here the two maxima and the sum of the
    residuals are exchanged as scalars (two tiny all-reduces) and the maximum's gradient is
    attached on the rank that owns the arg-max trajectory.
    Shapes as in the reference: rewards [b], fwd_probs / back_probs [b, T] for this rank's b
    trajectories (b may be 0); global_batch = sum of b over the ranks."""
    eps = 1e-9
    dt, dev = fwd_probs.dtype, fwd_probs.device
    total_flow = total_flow.to(dev).to(dt)
    rewards = rewards.to(dev).to(dt)
    back_probs = back_probs.to(dev).to(dt)
    lf = torch.log(fwd_probs + eps).sum(dim=-1)
    lb = torch.log(back_probs + eps).sum(dim=-1)
    on = dist.is_initialized() and dist.get_world_size(group) > 1
    neg = torch.finfo(dt).min
    mx = torch.stack([lf.detach().max() if lf.numel() else torch.tensor(neg, dtype=dt, device=dev),
                      lb.detach().max() if lb.numel() else torch.tensor(neg, dtype=dt, device=dev)])
    if on:
        _all_reduce(mx, dist.ReduceOp.MAX, group)
    resid = (torch.log(total_flow + eps) + (lf - mx[0])) - (torch.log(rewards + eps) + (lb - mx[1]))
    stats = torch.stack([resid.detach().sum(), (resid.detach() ** 2).sum()])
    if on:
        _all_reduce(stats, dist.ReduceOp.SUM, group)
    r_sum, loss_value = stats[0], stats[1] / global_batch
    surrogate = (resid ** 2).sum() / global_batch
    # d loss / d max_f = -(2/B) * sum_b resid_b ; d loss / d max_b = +(2/B) * sum_b resid_b.
    # torch.max routes its gradient to the first arg-max element; ties across ranks go to the lowest rank.
    me = dist.get_rank(group) if on else 0
    far = 1 << 30
    owner = torch.tensor([me if (v.numel() and bool(v.detach().max() == mx[i])) else far
                          for i, v in enumerate((lf, lb))], device=dev, dtype=torch.int64)
    if on:                                   # every rank takes part, whatever its shard holds
        _all_reduce(owner, dist.ReduceOp.MIN, group)
    for side, vec, sign in ((0, lf, -1.0), (1, lb, 1.0)):
        if int(owner[side]) == me and vec.requires_grad:
            surrogate = surrogate + sign * (2.0 / global_batch) * r_sum * vec[torch.argmax(vec.detach())]
    return surrogate, loss_value.detach()


def allreduce_gradients(params, group=None, bucket_bytes: int = 64 << 20, average: bool = False):
    """Sum (or average) the gradients of `params` over the ranks in flat buckets of at most
    `bucket_bytes` (the dominant tensor is the forward policy's fc weight, hid x (E+1),
    policy.py:30): one all-reduce per bucket, issued back to back (async) and awaited together,
    so a bucket's reduction overlaps the packing of the next. Parameters without a gradient
    on some rank (an empty shard) contribute zeros. Returns the number of bytes reduced."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    if world == 1:
        return 0
    buckets, cur, cur_bytes = [], [], 0
    for p in params:
        nb = p.grad.numel() * p.grad.element_size()
        if cur and (cur_bytes + nb > bucket_bytes or p.grad.dtype != cur[0].grad.dtype or p.grad.device != cur[0].grad.device):
            buckets.append(cur)
            cur, cur_bytes = [], 0
        cur.append(p)
        cur_bytes += nb
    if cur:
        buckets.append(cur)
    pending, total = [], 0
    for bk in buckets:
        flat = torch.cat([p.grad.reshape(-1) for p in bk])
        total += flat.numel() * flat.element_size()
        work, staged, _ = _all_reduce(flat, dist.ReduceOp.SUM, group, async_op=True)
        pending.append((work, staged, bk))
    for work, flat, bk in pending:
        work.wait()
        if average:
            flat.div_(world)
        off = 0
        for p in bk:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))        # (device -> host when the gradients live on the host)
            off += n
    return total


def data_parallel_step(model, optimizer, s0_shard, global_batch: int, group=None, generator=None,
                       method: str = "step", bucket_bytes: int = 64 << 20):
    """One epoch of the reference's training loop (GFlowNet100.py:291-315) with the batch sharded
    over the ranks: this rank samples and scores `len(s0_shard)` trajectories on its GPU
    (`model.sample_states`, rewards by the SPAI kernels), forms the shard-local surrogate of the
    global trajectory-balance loss, all-reduces the policy gradients and steps the optimiser
    (identical parameters on every rank before and after). Returns (loss, log); the per-trajectory
    rewards stay on the rank (gather them with `gather_rewards` only for logging)."""
    from .sampler import trajectory_balance_loss
    log = model.sample_states(s0_shard, return_log=True, generator=generator, method=method)
    params = [p for p in model.parameters() if p.requires_grad]
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        loss = trajectory_balance_loss(log.total_flow, log.rewards, log.fwd_probs, log.back_probs)
        value = loss.detach()
    else:
        loss, value = dp_trajectory_balance_loss(log.total_flow, log.rewards, log.fwd_probs, log.back_probs,
                                                 global_batch, group)
    if bool(torch.isnan(value) | torch.isinf(value)):          # GFlowNet100.py:306-308: skip the step
        optimizer.zero_grad()
        return value, log
    loss.backward()
    allreduce_gradients(params, group, bucket_bytes)
    optimizer.step()
    optimizer.zero_grad()
    return value, log
