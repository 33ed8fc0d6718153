"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun), NCCL.
Asserts that the trajectory-sharded and the row-sharded evaluations reproduce the
single-rank rewards; prints MULTI_OK on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gflownet_spai_b200 import dist as sdist  # noqa: E402
from gflownet_spai_b200 import synth  # noqa: E402
from gflownet_spai_b200.env import SpaiContext  # noqa: E402


class _Fwd(torch.nn.Module):
    """Stand-in for policy.ForwardPolicy (GATv2Conv is not installed): same call convention."""

    def __init__(self, a):
        super().__init__()
        self.logit = torch.nn.Parameter(torch.linspace(-1.0, 1.0, a, dtype=torch.float64))
        self.alpha = torch.nn.Parameter(torch.tensor(0.0, dtype=torch.float64))

    def forward(self, data, actions):
        return torch.softmax(self.logit[None, :], dim=1), torch.sigmoid(self.alpha)


class _Bwd(torch.nn.Module):
    def __init__(self, t):
        super().__init__()
        self.w = torch.nn.Parameter(torch.linspace(-0.5, 0.5, t, dtype=torch.float64))

    def forward(self, trajectories):
        p = torch.sigmoid(self.w)[None, : trajectories.shape[1]].expand(trajectories.shape[0], -1)
        return torch.where(trajectories >= 0, p, torch.ones_like(p))


def data_parallel_check(dev, rank, world):
    """2 ranks x B/2 reproduce the gradient and the loss of 1 rank x B (GFlowNet100.py:291-315), on
    trajectories sampled by K4 and rewards from the SPAI kernels; then one optimiser step keeps the
    replicas identical."""
    from gflownet_spai_b200.env import PreconditionerEnv
    from gflownet_spai_b200.sampler import GFlowNet, trajectory_balance_loss
    a = synth.poisson2d(6)
    coo = a.tocoo()
    idx = torch.tensor(np.stack([coo.row, coo.col]))
    init = torch.sparse_coo_tensor(idx, torch.tensor(coo.data, dtype=torch.float32), a.shape)
    env = PreconditionerEnv(a.shape[0], init, init.clone(), device=dev.index)
    batch = 8
    model = GFlowNet(_Fwd(env.num_actions), _Bwd(env.num_actions + 1), env)
    gen = torch.Generator(device=dev).manual_seed(123)
    log = model.sample_states([init] * batch, return_log=True, generator=gen)        # same draws on every rank
    params = [q for q in model.parameters() if q.requires_grad]
    full = trajectory_balance_loss(log.total_flow, log.rewards, log.fwd_probs, log.back_probs)
    want = torch.autograd.grad(full, params, retain_graph=True, allow_unused=True)
    lo, hi = sdist.shard_bounds(batch, world, rank)
    loss, value = sdist.dp_trajectory_balance_loss(log.total_flow, log.rewards[lo:hi], log.fwd_probs[lo:hi],
                                                   log.back_probs[lo:hi], batch)
    loss.backward()
    nbytes = sdist.allreduce_gradients(params)
    assert nbytes == sum(q.numel() * q.element_size() for q in params)
    assert abs(float(value) - float(full)) <= 1e-10 * max(1.0, abs(float(full)))
    for q, w in zip(params, want):
        w = torch.zeros_like(q) if w is None else w
        np.testing.assert_allclose(q.grad.cpu().numpy(), w.cpu().numpy(), rtol=1e-9, atol=1e-12)
    # a full step through the public entry point keeps the replicas in lock step
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    opt.zero_grad()
    gen2 = torch.Generator(device=dev).manual_seed(1000 + rank)                       # each rank its own shard
    value, _ = sdist.data_parallel_step(model, opt, [init] * (batch // world), batch // world * world, generator=gen2)
    flat = torch.cat([q.detach().reshape(-1) for q in params]).to(dev)
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(ref, flat), "replicas diverged after the data-parallel step"
    env.ctx.close()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    p = synth.make_problem("cfg2", 0.25)
    coo = p.a.tocoo()
    ctx = SpaiContext(p.n, p.edge_row, p.edge_col, p.edge_val, coo.row, coo.col, coo.data, device=local)
    batch = 256 + 3                                           # ragged shards
    acts = torch.from_numpy(synth.make_trajectories(p.num_edges, batch, seed0=31)).to(dev)
    for mode, dt, tol in (("copy", torch.float32, 1e-6), ("copy", torch.float64, 1e-12), ("ls_gram", torch.float64, 1e-11)):
        full = ctx.reward_batch(acts, 0.5, mode, dt)
        mine = sdist.shard_rows(acts, world, rank)
        local_out = ctx.reward_batch(mine.contiguous(), 0.5, mode, dt)
        got = sdist.gather_rewards(local_out["reward"], batch)
        assert got.shape == full["reward"].shape
        np.testing.assert_allclose(got.cpu().numpy(), full["reward"].cpu().numpy(), rtol=tol, atol=tol * 1e3,
                                   err_msg=f"trajectory sharding {mode} {dt}")
        rows = sdist.reward_row_sharded(ctx, acts, 0.5, mode, dt)
        np.testing.assert_allclose(rows["reward"].cpu().numpy(), full["reward"].cpu().numpy(), rtol=tol, atol=tol * 1e3,
                                   err_msg=f"row sharding {mode} {dt}")
        assert torch.equal(rows["nnz_m"], full["nnz_m"])
    # every rank holds the same gathered vector
    chk = got.clone()
    dist.all_reduce(chk, op=dist.ReduceOp.MAX)
    assert torch.equal(chk, got)
    ctx.close()
    data_parallel_check(dev, rank, world)
    dist.barrier()
    if rank == 0:
        print("MULTI_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
