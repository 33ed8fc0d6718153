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
    # data-parallel training step: 2 ranks x B/2 give the gradient of 1 rank x B
    if hasattr(sdist, "_selftest_data_parallel_step"):
        sdist._selftest_data_parallel_step(ctx, p, dev, rank, world)
    ctx.close()
    dist.barrier()
    if rank == 0:
        print("MULTI_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
