"""Real multi-GPU test (NCCL, one process per GPU): 2 ranks reproduce the 1-rank rewards with
trajectory sharding (all-gather of rewards) and with row sharding (all-reduce of partial sums).
Skipped on a single-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import os
import subprocess
import sys

import pytest
import torch

import conftest

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_ranks_reproduce_single_rank_rewards():
    port = 29500 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(conftest.ROOT, "tests", "_multi_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTI_OK" in out.stdout
