"""World-size-2 gloo tests of the trajectory sharding (host logic only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gflownet_spai_b200.dist import gather_rewards, shard_bounds, shard_rows


def test_shard_bounds_cover_the_batch_exactly():
    for batch in (0, 1, 2, 7, 4096, 16385):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    acts = torch.arange(batch * 3, dtype=torch.int64).reshape(batch, 3)
    mine = shard_rows(acts)
    local = mine[:, 0].to(torch.float64) * 0.5          # stand-in "reward" = f(trajectory)
    full = gather_rewards(local, batch)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), full.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [5, 8, 1])
def test_two_rank_gather_restores_trajectory_order(tmp_path, batch):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, batch, str(tmp_path)), nprocs=2, join=True)
    want = np.arange(batch) * 3 * 0.5
    for r in range(2):
        got = np.load(tmp_path / f"r{r}.npy")
        assert np.array_equal(got, want)


# ---------------------------------------------------------------------------
# data-parallel training step: 2 ranks x B/2 give the gradient of 1 rank x B
# ---------------------------------------------------------------------------
def _toy_batch(batch, tlen, n_act, seed=0):
    g = torch.Generator().manual_seed(seed)
    actions = torch.randint(0, n_act, (batch, tlen), generator=g)
    lens = torch.randint(1, tlen + 1, (batch,), generator=g)
    valid = torch.arange(tlen)[None, :] < lens[:, None]
    rewards = torch.rand(batch, generator=g, dtype=torch.float64) * 900 + 1
    return actions, valid, rewards


def _toy_probs(theta_f, theta_b, actions, valid):
    """Differentiable stand-ins for Log.fwd_probs / Log.back_probs (1.0 on padding, as the reference logs)."""
    pf = torch.softmax(theta_f, dim=0)[actions]
    pb = torch.sigmoid(theta_b)[None, : actions.shape[1]].expand_as(pf)
    one = torch.ones_like(pf)
    return torch.where(valid, pf, one), torch.where(valid, pb, one)


def _dp_worker(rank, world, port, batch, out_dir):
    from gflownet_spai_b200.dist import allreduce_gradients, dp_trajectory_balance_loss, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_act, tlen = 11, 6
    actions, valid, rewards = _toy_batch(batch, tlen, n_act)
    g = torch.Generator().manual_seed(5)
    theta_f = torch.randn(n_act, generator=g, dtype=torch.float64, requires_grad=True)
    theta_b = torch.randn(tlen, generator=g, dtype=torch.float64, requires_grad=True)
    unused = torch.zeros(3, dtype=torch.float64, requires_grad=True)          # a parameter no rank touches
    lo, hi = shard_bounds(batch, world, rank)
    pf, pb = _toy_probs(theta_f, theta_b, actions[lo:hi], valid[lo:hi])
    loss, value = dp_trajectory_balance_loss(torch.ones(1), rewards[lo:hi], pf, pb, batch)
    loss.backward()
    nbytes = allreduce_gradients([theta_f, theta_b, unused], bucket_bytes=64)   # tiny buckets: several all-reduces
    np.savez(os.path.join(out_dir, f"dp{rank}.npz"), gf=theta_f.grad.numpy(), gb=theta_b.grad.numpy(),
             gu=unused.grad.numpy(), value=float(value), nbytes=nbytes)
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [7, 2, 1])
def test_two_rank_gradient_equals_single_process_gradient(tmp_path, batch):
    from gflownet_spai_b200.sampler import trajectory_balance_loss
    n_act, tlen = 11, 6
    actions, valid, rewards = _toy_batch(batch, tlen, n_act)
    g = torch.Generator().manual_seed(5)
    theta_f = torch.randn(n_act, generator=g, dtype=torch.float64, requires_grad=True)
    theta_b = torch.randn(tlen, generator=g, dtype=torch.float64, requires_grad=True)
    pf, pb = _toy_probs(theta_f, theta_b, actions, valid)
    want = trajectory_balance_loss(torch.ones(1), rewards, pf, pb)             # the reference's loss on the full batch
    want.backward()
    port = _free_port()
    mp.spawn(_dp_worker, args=(2, port, batch, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = np.load(tmp_path / f"dp{r}.npz")
        np.testing.assert_allclose(got["gf"], theta_f.grad.numpy(), rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(got["gb"], theta_b.grad.numpy(), rtol=1e-10, atol=1e-12)
        assert np.all(got["gu"] == 0)
        assert float(got["value"]) == pytest.approx(float(want), rel=1e-12)
        assert int(got["nbytes"]) == 8 * (n_act + tlen + 3)
