"""The named BASELINE.json configs: scaled-down instances against the oracle
(every ls kernel class gets exercised), and the full-size headline config
through size-independent properties plus an oracle spot check."""
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


@pytest.mark.parametrize("cfg,scale,ls_class", [
    ("cfg2", 0.125, 0),      # 32x32 grid, k<=8,  |I|<=18  -> row kernel (8,2,9)
    ("cfg3", 0.1875, 3),     # 12^3 grid,  k<=16, |I|<=49  -> column kernel W=16, QMAX=52
    ("cfg4", 0.046875, 5),   # 24x24 grid, k<=32, |I|<=50  -> column kernel W=32, QMAX=52
    ("cfg5", 0.0015, None),  # n=1500 banded + power law: mixed classes incl. generic
])
def test_scaled_configs_match_oracle(cfg, scale, ls_class):
    p = synth.make_problem(cfg, scale)
    ctx = _ctx(p)
    info = ctx.info()
    if ls_class is not None:
        assert info.ls_class_rows[ls_class] > 0
    acts = synth.make_trajectories(p.num_edges, 4, seed0=21)
    t = torch.from_numpy(acts).cuda()
    a32, v32 = p.a.astype(np.float32), p.edge_val.astype(np.float32)
    want32 = orc.reward_batch_copy(p.n, p.edge_row, p.edge_col, v32, a32, acts, 0.5, dtype=np.float32)
    got32 = ctx.reward_batch(t, 0.5, "copy", torch.float32)
    np.testing.assert_allclose(got32["reward"].cpu().numpy(), want32["reward"], rtol=1e-4, atol=2e-2)
    assert np.array_equal(got32["nnz_m"].cpu().numpy(), want32["nnz_m"])
    want64 = orc.reward_batch_copy(p.n, p.edge_row, p.edge_col, p.edge_val, p.a, acts, 0.5, dtype=np.float64)
    got64 = ctx.reward_batch(t, 0.5, "copy", torch.float64)
    np.testing.assert_allclose(got64["reward"].cpu().numpy(), want64["reward"], rtol=1e-10, atol=1e-8)
    wls = orc.reward_batch_ls(p.n, p.edge_row, p.edge_col, p.a, acts[:2], 0.5, dtype=np.float64,
                              baseline_dtype=np.float64)
    gls = ctx.reward_batch(t[:2], 0.5, "ls", torch.float64)
    np.testing.assert_allclose(gls["residual"].cpu().numpy(), wls["residual"], rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(gls["reward"].cpu().numpy(), wls["reward"], rtol=1e-10, atol=1e-7)
    wls32 = orc.reward_batch_ls(p.n, p.edge_row, p.edge_col, p.a, acts[:2], 0.5, dtype=np.float32,
                                baseline_dtype=np.float32)
    gls32 = ctx.reward_batch(t[:2], 0.5, "ls", torch.float32)
    np.testing.assert_allclose(gls32["reward"].cpu().numpy(), wls32["reward"], rtol=1e-4, atol=2e-2)
    ctx.close()


def test_full_size_cfg2_properties_and_spot_check():
    """BASELINE.json configs[1] at full size (n = 65 536, E = 524 280)."""
    p = synth.make_problem("cfg2")
    assert p.n == 65536 and p.num_edges == 524280
    ctx = _ctx(p)
    info = ctx.info()
    assert info.max_row_slots == 8 and info.max_row_union == 18 and info.has_duplicates == 0
    bsz = 48
    acts = synth.make_trajectories(p.num_edges, bsz, seed0=1000)
    t = torch.from_numpy(acts)
    dev = ctx.reward_batch(t.cuda(), 0.5, "copy", torch.float32)
    host = ctx.reward_batch(t, 0.5, "copy", torch.float32)            # host entry (trimmed H2D) == device entry
    assert torch.equal(dev["reward"].cpu(), host["reward"]) and torch.equal(dev["nnz_m"].cpu(), host["nnz_m"])
    # order of a trajectory's actions and repeated ids do not matter (set semantics, utils.py:318)
    rng = np.random.default_rng(0)
    shuf = acts.copy()
    for b in range(bsz):
        n_valid = int((acts[b] >= 0).sum())
        shuf[b, :n_valid] = rng.permutation(acts[b, :n_valid])
    dup = np.concatenate([acts, acts[:, :64]], axis=1)
    r_shuf = ctx.reward_batch(torch.from_numpy(shuf).cuda(), 0.5, "copy", torch.float32)["reward"]
    r_dup = ctx.reward_batch(torch.from_numpy(dup).cuda(), 0.5, "copy", torch.float32)["reward"]
    assert torch.equal(r_shuf, dev["reward"]) and torch.equal(r_dup, dev["reward"])
    # a shard of the batch scores exactly like the same rows inside the batch
    part = ctx.reward_batch(t[16:32].cuda(), 0.5, "copy", torch.float32)["reward"]
    assert torch.allclose(part, dev["reward"][16:32], rtol=1e-12, atol=1e-9)
    # nnz(M) = E - distinct deletions; terminal-only == nothing removed
    valid = [np.unique(acts[b][(acts[b] >= 0) & (acts[b] < p.num_edges)]).size for b in range(bsz)]
    assert dev["nnz_m"].cpu().tolist() == [p.num_edges - v for v in valid]
    term = ctx.reward_batch(torch.tensor([[p.num_edges, -1]]).cuda(), 0.5, "copy", torch.float64)
    assert int(term["nnz_m"][0]) == p.num_edges
    # ls never worse than copy, pattern by pattern
    ls = ctx.reward_batch(t[:8].cuda(), 0.5, "ls", torch.float64)
    cp = ctx.reward_batch(t[:8].cuda(), 0.5, "copy", torch.float64)
    assert torch.all(ls["residual"] <= cp["residual"] + 1e-9)
    # oracle spot check on three full-size patterns (copy fp32 / fp64)
    sel = [0, 17, 41]
    want = orc.reward_batch_copy(p.n, p.edge_row, p.edge_col, p.edge_val.astype(np.float32),
                                 p.a.astype(np.float32), acts[sel], 0.5, dtype=np.float32)
    np.testing.assert_allclose(dev["reward"].cpu().numpy()[sel], want["reward"], rtol=1e-4, atol=2e-2)
    want64 = orc.reward_batch_copy(p.n, p.edge_row, p.edge_col, p.edge_val, p.a, acts[sel[:1]], 0.5, dtype=np.float64)
    np.testing.assert_allclose(cp["reward"].cpu().numpy()[:1], want64["reward"], rtol=1e-10, atol=1e-8)
    ctx.close()


@pytest.mark.parametrize("cfg", ["cfg3", "cfg4"])
def test_full_size_cfg3_cfg4_spot_checks(cfg):
    """BASELINE.json configs[2] and [3] at full size (n = 262 144): device and
    host entry points agree, one pattern is re-scored by the oracle, the ls
    residual never exceeds the copy residual, fp32 and fp64 agree to fp32 accuracy."""
    p = synth.make_problem(cfg)
    assert p.n == 262144
    ctx = _ctx(p)
    acts = synth.make_trajectories(p.num_edges, 6, seed0=300)
    t = torch.from_numpy(acts)
    dev32 = ctx.reward_batch(t.cuda(), 0.5, "copy", torch.float32)
    host32 = ctx.reward_batch(t, 0.5, "copy", torch.float32)
    assert torch.equal(dev32["reward"].cpu(), host32["reward"])
    want = orc.reward_batch_copy(p.n, p.edge_row, p.edge_col, p.edge_val.astype(np.float32),
                                 p.a.astype(np.float32), acts[:1], 0.5, dtype=np.float32)
    np.testing.assert_allclose(dev32["reward"].cpu().numpy()[:1], want["reward"], rtol=1e-4, atol=2e-2)
    assert int(dev32["nnz_m"][0]) == int(want["nnz_m"][0])
    dev64 = ctx.reward_batch(t.cuda(), 0.5, "copy", torch.float64)
    # same pattern, two precisions: residuals agree to fp32 accuracy (baselines differ per dtype)
    assert torch.allclose(dev32["residual"], dev64["residual"], rtol=1e-5)
    ls = ctx.reward_batch(t[:3].cuda(), 0.5, "ls", torch.float64)
    assert torch.all(ls["residual"] <= dev64["residual"][:3] + 1e-9)
    ls32 = ctx.reward_batch(t[:3].cuda(), 0.5, "ls", torch.float32)
    assert torch.allclose(ls32["residual"], ls["residual"], rtol=1e-4)
    ctx.close()


@pytest.mark.parametrize("mode,dtype,tol", [("copy", torch.float32, 1e-6), ("copy", torch.float64, 1e-12),
                                            ("ls", torch.float64, 1e-12)])
def test_row_range_partials_sum_to_the_full_reward(mode, dtype, tol):
    """Row sharding (secondary): partial sums over disjoint row ranges,
    added and finalised, equal the full evaluation (emulates 3 ranks on one GPU)."""
    from gflownet_spai_b200.dist import shard_bounds, reward_row_sharded
    p = synth.make_problem("cfg5", 0.004)                 # n = 4000, mixed ls classes incl. generic rows
    ctx = _ctx(p)
    acts = torch.from_numpy(synth.make_trajectories(p.num_edges, 9, seed0=4)).cuda()
    full = ctx.reward_batch(acts, 0.3, mode, dtype)
    tot = torch.zeros(9, dtype=torch.float64, device="cuda")
    for r in range(3):
        lo, hi = shard_bounds(p.n, 3, r)
        part, nnz = ctx.reward_rows(acts, lo, hi, mode, dtype)
        tot += part
        assert torch.equal(nnz, full["nnz_m"])
    fin = ctx.finalize_rewards(tot, nnz, 0.3, dtype)
    assert torch.allclose(fin["residual"], full["residual"], rtol=tol, atol=1e-12)
    assert torch.allclose(fin["reward"], full["reward"], rtol=tol, atol=1e-6)
    one = reward_row_sharded(ctx, acts, 0.3, mode, dtype)                       # world size 1 path
    assert torch.allclose(one["reward"], full["reward"], rtol=tol, atol=1e-6)
    empty, _ = ctx.reward_rows(acts, 17, 17, mode, dtype)
    assert float(empty.abs().max()) == 0.0
    ctx.close()
