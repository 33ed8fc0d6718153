"""K3s, the deletion-driven copy kernel (few deletions per trajectory): against the
oracle, against the row-sweep kernel K3 on the same inputs, and on inputs that force
its rare paths (several deletions in a row, rows with more than 32 candidates)."""
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


class _force:
    """SPAI_K3_SPARSE=0/1 forces the row-sweep / deletion-driven kernel (read per call)."""
    def __init__(self, v):
        self.v = v

    def __enter__(self):
        self.old = os.environ.get("SPAI_K3_SPARSE")
        os.environ["SPAI_K3_SPARSE"] = self.v

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("SPAI_K3_SPARSE", None)
        else:
            os.environ["SPAI_K3_SPARSE"] = self.old


@pytest.mark.parametrize("cfg,scale", [("cfg2", 0.125), ("cfg3", 0.1875), ("cfg4", 0.046875), ("cfg5", 0.004)])
def test_short_trajectories_match_oracle_and_row_sweep(cfg, scale):
    p = synth.make_problem(cfg, scale)
    ctx = _ctx(p)
    acts = synth.make_trajectories(p.num_edges, 70, seed0=5, max_frac=0.02)     # T <= E/50: K3s is picked
    assert acts.shape[1] * 40 <= p.num_edges
    t = torch.from_numpy(acts).cuda()
    got32 = ctx.reward_batch(t, 0.5, "copy", torch.float32)
    got64 = ctx.reward_batch(t, 0.5, "copy", torch.float64)
    with _force("0"):
        ref32 = ctx.reward_batch(t, 0.5, "copy", torch.float32)
        ref64 = ctx.reward_batch(t, 0.5, "copy", torch.float64)
    assert torch.equal(got32["nnz_m"], ref32["nnz_m"])
    assert torch.allclose(got32["reward"], ref32["reward"], rtol=1e-5, atol=1e-3)
    assert torch.allclose(got64["residual"], ref64["residual"], rtol=1e-12, atol=1e-12)
    a32, v32 = p.a.astype(np.float32), p.edge_val.astype(np.float32)
    want32 = orc.reward_batch_copy(p.n, p.edge_row, p.edge_col, v32, a32, acts[:3], 0.5, dtype=np.float32)
    np.testing.assert_allclose(got32["reward"].cpu().numpy()[:3], want32["reward"], rtol=1e-4, atol=2e-2)
    want64 = orc.reward_batch_copy(p.n, p.edge_row, p.edge_col, p.edge_val, p.a, acts[:3], 0.5, dtype=np.float64)
    np.testing.assert_allclose(got64["reward"].cpu().numpy()[:3], want64["reward"], rtol=1e-10, atol=1e-8)
    # the host entry point takes the same kernel: identical numbers
    host = ctx.reward_batch(torch.from_numpy(acts), 0.5, "copy", torch.float32)
    assert torch.equal(host["reward"], got32["reward"].cpu())
    ctx.close()


@pytest.mark.parametrize("cfg,scale", [("cfg2", 0.125), ("cfg5", 0.004)])
def test_forced_on_dense_trajectories_every_row_takes_the_full_path(cfg, scale):
    """Half of the candidates deleted: almost every row has several deleted slots, so K3s
    re-evaluates whole rows (heavy list) — must still equal the row sweep."""
    p = synth.make_problem(cfg, scale)
    ctx = _ctx(p)
    acts = synth.make_trajectories(p.num_edges, 33, seed0=9)
    t = torch.from_numpy(acts).cuda()
    with _force("0"):
        ref32 = ctx.reward_batch(t, 0.5, "copy", torch.float32)
        ref64 = ctx.reward_batch(t, 0.5, "copy", torch.float64)
    with _force("1"):
        got32 = ctx.reward_batch(t, 0.5, "copy", torch.float32)
        got64 = ctx.reward_batch(t, 0.5, "copy", torch.float64)
        # duplicates / shuffles of the action list change nothing (the kernel walks the mask)
        shuf = acts[:, np.random.default_rng(0).permutation(acts.shape[1])]
        again = ctx.reward_batch(torch.from_numpy(shuf).cuda(), 0.5, "copy", torch.float32)
    assert torch.equal(got32["reward"], again["reward"])
    assert torch.allclose(got32["reward"], ref32["reward"], rtol=1e-5, atol=1e-3)
    assert torch.allclose(got64["residual"], ref64["residual"], rtol=1e-12, atol=1e-12)
    ctx.close()


def test_rows_with_more_than_32_candidates_and_duplicates():
    rng = np.random.default_rng(3)
    n = 96
    a = sp.random(n, n, density=0.06, random_state=2, format="csr") + sp.identity(n, format="csr") * 3.0
    a = sp.csr_matrix(a)
    a.sort_indices()
    rows, cols = [], []
    for i in range(n):
        k = 40 if i % 7 == 0 else 6                       # wide rows: kept bits come from the mask row
        cc = rng.choice(n, size=k, replace=False)
        if i % 5 == 0:
            cc = np.concatenate([cc, cc[:2]])             # repeated coordinates
        rows.append(np.full(cc.size, i))
        cols.append(cc)
    r = np.concatenate(rows).astype(np.int64)
    c = np.concatenate(cols).astype(np.int64)
    v = rng.uniform(-1, 1, r.size)
    perm = rng.permutation(r.size)                        # caller's (uncoalesced, unsorted) edge order
    r, c, v = r[perm], c[perm], v[perm]
    coo = a.tocoo()
    from gflownet_spai_b200.env import SpaiContext
    ctx = SpaiContext(n, r, c, v, coo.row, coo.col, coo.data)
    for frac, seed in ((0.03, 1), (0.5, 2)):
        acts = synth.make_trajectories(r.size, 40, seed0=seed, max_frac=frac)
        t = torch.from_numpy(acts).cuda()
        want = orc.reward_batch_copy(n, r, c, v, a, acts, 0.5, dtype=np.float64)
        with _force("1"):
            got = ctx.reward_batch(t, 0.5, "copy", torch.float64)
            got32 = ctx.reward_batch(t, 0.5, "copy", torch.float32)
        np.testing.assert_allclose(got["reward"].cpu().numpy(), want["reward"], rtol=1e-10, atol=1e-8)
        assert np.array_equal(got["nnz_m"].cpu().numpy(), want["nnz_m"])
        want32 = orc.reward_batch_copy(n, r, c, v.astype(np.float32), a.astype(np.float32), acts, 0.5, dtype=np.float32)
        np.testing.assert_allclose(got32["reward"].cpu().numpy(), want32["reward"], rtol=1e-4, atol=2e-2)
    ctx.close()


def test_row_ranges_with_the_deletion_driven_kernel():
    from gflownet_spai_b200.dist import shard_bounds
    p = synth.make_problem("cfg5", 0.004)
    ctx = _ctx(p)
    acts = torch.from_numpy(synth.make_trajectories(p.num_edges, 9, seed0=4, max_frac=0.02)).cuda()
    for dtype, tol in ((torch.float32, 1e-6), (torch.float64, 1e-12)):
        full = ctx.reward_batch(acts, 0.3, "copy", dtype)
        tot = torch.zeros(9, dtype=torch.float64, device="cuda")
        for r in range(3):
            lo, hi = shard_bounds(p.n, 3, r)
            part, nnz = ctx.reward_rows(acts, lo, hi, "copy", dtype)
            tot += part
        fin = ctx.finalize_rewards(tot, nnz, 0.3, dtype)
        assert torch.allclose(fin["reward"], full["reward"], rtol=tol, atol=1e-6)
        empty, _ = ctx.reward_rows(acts, 17, 17, "copy", dtype)
        assert float(empty.abs().max()) == 0.0
    ctx.close()


def test_taken_bitmask_entry_point_with_a_deletion_hint():
    """reward_from_taken sees no action list; spai_ctx_set_deletion_hint lets it take the
    deletion-driven kernel. The hint selects a kernel only: same numbers with and without."""
    p = synth.make_problem("cfg5", 0.004)
    ctx = _ctx(p)
    acts = synth.make_trajectories(p.num_edges, 12, seed0=8, max_frac=0.01)
    words = (p.num_edges + 1 + 31) // 32
    taken = np.zeros((12, words), dtype=np.uint32)
    for b in range(12):
        ids = acts[b][(acts[b] >= 0)]
        np.bitwise_or.at(taken[b], ids >> 5, (np.uint32(1) << (ids & 31).astype(np.uint32)))
    t = torch.from_numpy(taken.view(np.int32)).cuda()
    ctx.set_deletion_hint(0)
    plain = ctx.reward_from_taken(t, 0.5, "copy", torch.float64)
    ctx.set_deletion_hint(int(acts.shape[1]))
    hinted = ctx.reward_from_taken(t, 0.5, "copy", torch.float64)
    via_actions = ctx.reward_batch(torch.from_numpy(acts).cuda(), 0.5, "copy", torch.float64)
    assert torch.equal(hinted["nnz_m"], plain["nnz_m"])
    assert torch.allclose(hinted["reward"], plain["reward"], rtol=1e-12, atol=1e-9)
    assert torch.equal(hinted["reward"], via_actions["reward"])
    ctx.close()
