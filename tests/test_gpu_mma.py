"""K3m, the tensor-core copy-mode kernel (tcgen05.mma on the per-row Cholesky factor of the Gram matrix;
rows with <= 32 candidates, fp32, batches of >= 64 trajectories): against the oracle and against the
CUDA-core row sweep K3 (SPAI_K3_MMA=0) on the same inputs."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from gflownet_spai_b200 import synth
from oracle import spai_oracle as orc

pytestmark = pytest.mark.gpu


def _ctx(p):
    from gflownet_spai_b200.env import SpaiContext
    coo = p.a.tocoo()
    return SpaiContext(p.n, p.edge_row, p.edge_col, p.edge_val, coo.row, coo.col, coo.data, device=0)


class _env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _trajectories(num_edges, bsz, seed):
    """dense, short and empty trajectories in one batch"""
    acts = synth.make_trajectories(num_edges, bsz, seed0=seed)
    acts[1, :] = -1
    acts[1, 0] = num_edges                       # nothing removed
    acts[2, 5:] = -1                             # five deletions
    return acts


@pytest.mark.parametrize("cfg,scale,bsz,split", [("cfg3", 0.25, 64, "3"), ("cfg3", 0.25, 200, "2"), ("cfg4", 0.0625, 130, "3"),
                                                 ("cfg4", 0.0625, 600, "2"), ("cfg5", 0.004, 96, "3"), ("cfg5", 0.004, 1100, "2")])
def test_tensor_core_rows_match_row_sweep_and_oracle(cfg, scale, bsz, split):
    p = synth.make_problem(cfg, scale)
    with _env(SPAI_K3M_SPLIT=split):
        ctx = _ctx(p)
        k = ctx.info().max_row_slots
        assert 8 < k <= 32
        acts = _trajectories(p.num_edges, bsz, 17)
        t = torch.from_numpy(acts).cuda()
        got = ctx.reward_batch(t, 0.5, "copy", torch.float32)
        r16, r32 = ctx.k3m_rows()
        assert r16 + r32 == p.n and (r32 > 0) == (k > 16)         # the tensor-core records exist: the call above ran K3m
        with _env(SPAI_K3_MMA="0"):
            ref = ctx.reward_batch(t, 0.5, "copy", torch.float32)
    assert torch.equal(got["nnz_m"], ref["nnz_m"])
    tol = 2e-6 if split == "3" else 2e-5
    assert torch.allclose(got["residual"], ref["residual"], rtol=tol), float(((got["residual"] - ref["residual"]).abs() / ref["residual"]).max())
    sel = [0, 1, 2, bsz - 1]
    want = orc.reward_batch_copy(p.n, p.edge_row, p.edge_col, p.edge_val.astype(np.float32), p.a.astype(np.float32),
                                 acts[sel], 0.5, dtype=np.float32)
    np.testing.assert_allclose(got["reward"].cpu().numpy()[sel], want["reward"], rtol=1e-4, atol=2e-2)
    np.testing.assert_allclose(got["residual"].cpu().numpy()[sel], want["residual"], rtol=1e-4)
    ctx.close()


def test_tensor_core_rows_row_ranges_and_timing_label():
    """Row-sharded evaluation (spai_reward_rows_dev) adds up to the full reward through K3m too."""
    from gflownet_spai_b200.dist import shard_bounds
    p = synth.make_problem("cfg3", 0.25)
    ctx = _ctx(p)
    acts = torch.from_numpy(synth.make_trajectories(p.num_edges, 70, seed0=4)).cuda()
    full = ctx.reward_batch(acts, 0.3, "copy", torch.float32)
    assert sum(ctx.k3m_rows()) == p.n
    tot = torch.zeros(70, dtype=torch.float64, device="cuda")
    for r in range(3):
        lo, hi = shard_bounds(p.n, 3, r)
        part, nnz = ctx.reward_rows(acts, lo, hi, "copy", torch.float32)
        tot += part
    fin = ctx.finalize_rewards(tot, nnz, 0.3, torch.float32)
    assert torch.allclose(fin["reward"], full["reward"], rtol=1e-6, atol=1e-6)
    ctx.close()


def test_tensor_core_rows_semidefinite_and_zero_slots():
    """Explicit zeros in the pattern (a zero column of W), a row of A repeated with a different scale
    (dependent contributions: singular Gram matrix), rows without a diagonal entry in their union."""
    rng = np.random.default_rng(8)
    n = 96
    a = sp.random(n, n, density=0.08, random_state=3, format="lil")
    a.setdiag(rng.uniform(1.0, 2.0, n))
    a[5, :] = 2.5 * a[4, :]                       # rows 4 and 5 of A are parallel
    a = sp.csr_matrix(a)
    rows, cols = [], []
    for i in range(n):
        k = int(rng.integers(9, 15))
        cand = set(rng.choice(n, size=k, replace=False).tolist())
        if i % 7 == 0:
            cand |= {4, 5}
        if i % 11 == 3:
            cand.discard(i)
        cand = sorted(cand)
        rows += [i] * len(cand)
        cols += cand
    rows, cols = np.array(rows, dtype=np.int64), np.array(cols, dtype=np.int64)
    vals = rng.normal(size=rows.size)
    vals[rng.random(rows.size) < 0.1] = 0.0       # explicit zeros stay in the pattern (matrix_flops counts them)
    from gflownet_spai_b200.env import SpaiContext
    coo = a.tocoo()
    ctx = SpaiContext(n, rows, cols, vals, coo.row, coo.col, coo.data, device=0)
    acts = synth.make_trajectories(rows.size, 128, seed0=2)
    t = torch.from_numpy(acts).cuda()
    got = ctx.reward_batch(t, 0.5, "copy", torch.float32)
    assert sum(ctx.k3m_rows()) == n
    with _env(SPAI_K3_MMA="0"):
        ref = ctx.reward_batch(t, 0.5, "copy", torch.float32)
    assert torch.allclose(got["residual"], ref["residual"], rtol=5e-6)
    want = orc.reward_batch_copy(n, rows, cols, vals.astype(np.float32), a.astype(np.float32), acts[:6], 0.5, dtype=np.float32)
    np.testing.assert_allclose(got["residual"].cpu().numpy()[:6], want["residual"], rtol=1e-4)
    np.testing.assert_allclose(got["reward"].cpu().numpy()[:6], want["reward"], rtol=1e-4, atol=2e-2)
    ctx.close()
