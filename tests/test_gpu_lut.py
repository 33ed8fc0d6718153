"""K3t, the table-lookup copy kernel (patterns whose rows have <= 8 candidates, batches
of >= 64 trajectories): against the oracle and against the row sweep K3 on the same inputs."""
import os

import numpy as np
import pytest
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


@pytest.mark.parametrize("scale,bsz", [(0.125, 64), (0.125, 300), (1.0, 1100)])
def test_table_lookup_matches_row_sweep_and_oracle(scale, bsz):
    p = synth.make_problem("cfg2", scale)
    ctx = _ctx(p)
    assert ctx.info().max_row_slots <= 8
    acts = synth.make_trajectories(p.num_edges, bsz, seed0=31)
    t = torch.from_numpy(acts).cuda()
    got32 = ctx.reward_batch(t, 0.5, "copy", torch.float32)
    got64 = ctx.reward_batch(t, 0.5, "copy", torch.float64)
    with _env(SPAI_K3_LUT="0"):
        ref32 = ctx.reward_batch(t, 0.5, "copy", torch.float32)
        ref64 = ctx.reward_batch(t, 0.5, "copy", torch.float64)
    assert torch.equal(got32["nnz_m"], ref32["nnz_m"])
    assert torch.allclose(got32["residual"], ref32["residual"], rtol=2e-6)
    assert torch.allclose(got64["residual"], ref64["residual"], rtol=1e-13)
    sel = [0, bsz // 2, bsz - 1]
    want32 = orc.reward_batch_copy(p.n, p.edge_row, p.edge_col, p.edge_val.astype(np.float32),
                                   p.a.astype(np.float32), acts[sel], 0.5, dtype=np.float32)
    np.testing.assert_allclose(got32["reward"].cpu().numpy()[sel], want32["reward"], rtol=1e-4, atol=2e-2)
    want64 = orc.reward_batch_copy(p.n, p.edge_row, p.edge_col, p.edge_val, p.a, acts[sel[:1]], 0.5, dtype=np.float64)
    np.testing.assert_allclose(got64["reward"].cpu().numpy()[:1], want64["reward"], rtol=1e-10, atol=1e-8)
    # host entry point: same kernel, same numbers
    host = ctx.reward_batch(torch.from_numpy(acts[:80]), 0.5, "copy", torch.float32)
    dev = ctx.reward_batch(t[:80], 0.5, "copy", torch.float32)
    assert torch.equal(host["reward"], dev["reward"].cpu())
    ctx.close()


def test_table_lookup_row_ranges_and_golden():
    import conftest
    from gflownet_spai_b200.dist import shard_bounds
    from gflownet_spai_b200.env import SpaiContext
    p = synth.make_problem("cfg2", 0.125)
    ctx = _ctx(p)
    acts = torch.from_numpy(synth.make_trajectories(p.num_edges, 70, seed0=4)).cuda()
    for dtype, tol in ((torch.float32, 1e-6), (torch.float64, 1e-12)):
        full = ctx.reward_batch(acts, 0.3, "copy", dtype)
        tot = torch.zeros(70, dtype=torch.float64, device="cuda")
        for r in range(3):
            lo, hi = shard_bounds(p.n, 3, r)
            part, nnz = ctx.reward_rows(acts, lo, hi, "copy", dtype)
            tot += part
        fin = ctx.finalize_rewards(tot, nnz, 0.3, dtype)
        assert torch.allclose(fin["reward"], full["reward"], rtol=tol, atol=1e-6)
    ctx.close()
    # golden produced by the reference itself (k <= 8 pattern), batch replicated to reach the table path
    g = conftest.load_golden("poisson32_k8")
    n = int(g["n"])
    ctx = SpaiContext(n, g["edge_row"], g["edge_col"], g["edge_val"].astype(np.float64),
                      g["a_row"], g["a_col"], g["a_val"].astype(np.float64), device=0)
    acts = g["actions"]
    reps = (64 + acts.shape[0] - 1) // acts.shape[0]
    big = np.concatenate([acts] * reps, axis=0)
    out = ctx.reward_batch(torch.from_numpy(big).cuda(), float(g["alpha"]), "copy", torch.float32)
    want = np.concatenate([g["reward"]] * reps)
    np.testing.assert_allclose(out["reward"].cpu().numpy(), want, rtol=1e-4, atol=2e-2)
    ctx.close()


def test_ls_gram_table_matches_gram_kernel_and_oracle():
    p = synth.make_problem("cfg2", 0.125)
    ctx = _ctx(p)
    acts = synth.make_trajectories(p.num_edges, 130, seed0=77)
    t = torch.from_numpy(acts).cuda()
    got64 = ctx.reward_batch(t, 0.5, "ls_gram", torch.float64)
    got32 = ctx.reward_batch(t, 0.5, "ls_gram", torch.float32)
    with _env(SPAI_K3_LUT="0"):
        ref64 = ctx.reward_batch(t, 0.5, "ls_gram", torch.float64)
        ref32 = ctx.reward_batch(t, 0.5, "ls_gram", torch.float32)
    assert torch.allclose(got64["residual"], ref64["residual"], rtol=1e-12, atol=1e-12)
    assert torch.allclose(got32["residual"], ref32["residual"], rtol=1e-5)
    assert torch.equal(got64["nnz_m"], ref64["nnz_m"])
    wls = orc.reward_batch_ls(p.n, p.edge_row, p.edge_col, p.a, acts[:2], 0.5, dtype=np.float64, baseline_dtype=np.float64)
    np.testing.assert_allclose(got64["reward"].cpu().numpy()[:2], wls["reward"], rtol=1e-10, atol=1e-7)
    wls32 = orc.reward_batch_ls(p.n, p.edge_row, p.edge_col, p.a, acts[:2], 0.5, dtype=np.float32, baseline_dtype=np.float32)
    np.testing.assert_allclose(got32["reward"].cpu().numpy()[:2], wls32["reward"], rtol=1e-4, atol=2e-2)
    # row ranges through the table
    from gflownet_spai_b200.dist import shard_bounds
    tot = torch.zeros(130, dtype=torch.float64, device="cuda")
    for r in range(3):
        lo, hi = shard_bounds(p.n, 3, r)
        part, nnz = ctx.reward_rows(t, lo, hi, "ls_gram", torch.float64)
        tot += part
    fin = ctx.finalize_rewards(tot, nnz, 0.5, torch.float64)
    assert torch.allclose(fin["reward"], got64["reward"], rtol=1e-12, atol=1e-6)
    ctx.close()


def test_ls_gram_table_hands_dependent_columns_to_householder():
    """Masks whose Gram elimination meets a small pivot are NaN in the table; the lookup kernel
    must route those (row, trajectory) pairs to the Householder kernel."""
    import scipy.sparse as sp
    rng = np.random.default_rng(11)
    n = 64
    a = sp.random(n, n, density=0.08, random_state=4, format="lil") + sp.identity(n, format="lil") * 2.0
    a = sp.lil_matrix(a)
    a[5, :] = a[3, :]
    a = sp.csr_matrix(a)
    a.sort_indices()
    rows, cols = [], []
    for i in range(n):
        base = [3, 5] if i % 3 == 0 else []
        extra = rng.choice([c for c in range(n) if c not in (3, 5)], size=5, replace=False)
        cc = np.array(base + list(extra))
        rows.append(np.full(cc.size, i))
        cols.append(cc)
    r = np.concatenate(rows).astype(np.int64)
    c = np.concatenate(cols).astype(np.int64)
    v = rng.uniform(-1, 1, r.size)
    coo = a.tocoo()
    from gflownet_spai_b200.env import SpaiContext
    ctx = SpaiContext(n, r, c, v, coo.row, coo.col, coo.data)
    acts = synth.make_trajectories(r.size, 70, seed0=2, max_frac=0.3)
    acts[0, :] = -1
    want = orc.reward_batch_ls(n, r, c, a, acts[:8], 0.5, dtype=np.float64, baseline_dtype=np.float64)
    got = ctx.reward_batch(torch.from_numpy(acts).cuda(), 0.5, "ls_gram", torch.float64)
    np.testing.assert_allclose(got["residual"].cpu().numpy()[:8], want["residual"], rtol=1e-9, atol=1e-9)
    with _env(SPAI_K3_LUT="0"):
        ref = ctx.reward_batch(torch.from_numpy(acts).cuda(), 0.5, "ls_gram", torch.float64)
    assert torch.allclose(got["residual"], ref["residual"], rtol=1e-9, atol=1e-9)
    ctx.close()


def test_householder_table_matches_householder_kernels_and_oracle():
    """ls mode, >= 64 trajectories, rows with <= 8 candidates: the table is filled by the generic
    Householder kernel on the 256 uniform masks; same numbers as the register QR kernels."""
    p = synth.make_problem("cfg2", 0.125)
    ctx = _ctx(p)
    acts = synth.make_trajectories(p.num_edges, 70, seed0=15)
    t = torch.from_numpy(acts).cuda()
    for dtype, tol in ((torch.float64, 1e-11), (torch.float32, 2e-5)):
        got = ctx.reward_batch(t, 0.5, "ls", dtype)
        with _env(SPAI_K3_LUT="0"):
            ref = ctx.reward_batch(t, 0.5, "ls", dtype)
        assert torch.allclose(got["residual"], ref["residual"], rtol=tol, atol=1e-12)
        assert torch.equal(got["nnz_m"], ref["nnz_m"])
    got = ctx.reward_batch(t, 0.5, "ls", torch.float64)
    wls = orc.reward_batch_ls(p.n, p.edge_row, p.edge_col, p.a, acts[:2], 0.5, dtype=np.float64, baseline_dtype=np.float64)
    np.testing.assert_allclose(got["reward"].cpu().numpy()[:2], wls["reward"], rtol=1e-10, atol=1e-7)
    part, nnz = ctx.reward_rows(t, 100, 700, "ls", torch.float64)
    with _env(SPAI_K3_LUT="0"):
        part_ref, _ = ctx.reward_rows(t, 100, 700, "ls", torch.float64)
    assert torch.allclose(part, part_ref, rtol=1e-11, atol=1e-12)
    ctx.close()
