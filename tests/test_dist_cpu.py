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
