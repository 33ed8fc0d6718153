"""Full-size parity on the BASELINE configs 3, 4, 5 (VERDICT r1 'next' 1a): the sizes the large-E
kernels are tuned for, each checked against the CPU oracle (not only against invariants).

  cfg5 (n = 1e6, E = 10.9 M)   K0b masks bit-exact vs `i not in set(actions)`; copy fp32 / fp64 rewards
                               of SURVEY 8d trajectories vs the oracle; K3s vs the row sweep
  cfg3, cfg4                   ls and ls_gram vs LAPACK (`orc.ls_row_residual2`) on 2000 sampled rows
  K4                           masked categorical step at A = 524 281 (cfg2) and 8.4 M (cfg4)
  zero baselines               `inf` ratios of preconditioner.py:154,:158 (goldens made by the live reference)
"""
import os

import numpy as np
import pytest
import torch

from gflownet_spai_b200 import synth
from oracle import spai_oracle as orc
from conftest import load_golden

pytestmark = pytest.mark.gpu

_CACHE = {}


def _problem(name):
    if name not in _CACHE:
        from gflownet_spai_b200.env import SpaiContext
        for _, old in _CACHE.values():      # one full-size context at a time
            old.close()
        _CACHE.clear()
        p = synth.make_problem(name)
        coo = p.a.tocoo()
        ctx = SpaiContext(p.n, p.edge_row, p.edge_col, p.edge_val, coo.row, coo.col, coo.data, device=0)
        _CACHE[name] = (p, ctx)
    return _CACHE[name]


def _spec_trajectories(e, ids, max_frac=0.5):
    """trajectory b deletes floor(u_b * E / 2) distinct edges, then the terminal id."""
    rows, lens = [], []
    for b in ids:
        rng = np.random.default_rng(1000 + b)
        t = int(np.floor(rng.random() * e * max_frac))
        rows.append(rng.permutation(e)[:t].astype(np.int64))
        lens.append(t + 1)
    tmax = max(lens)
    acts = np.full((len(ids), tmax), -1, dtype=np.int64)
    for i, r in enumerate(rows):
        acts[i, : r.size] = r
        acts[i, r.size] = e
    return acts, np.asarray(lens, dtype=np.int32)


def test_cfg5_full_size_masks_and_copy_rewards_vs_oracle():
    p, ctx = _problem("cfg5")
    e = p.num_edges
    assert p.n == 1_000_000 and e > 10_000_000
    acts, lens = _spec_trajectories(e, [0, 1, 2, 3])
    t_acts = torch.from_numpy(acts).cuda()
    t_lens = torch.from_numpy(lens).cuda()
    # K0b masks, bit-exact, with and without row lengths, int64 and int32 ids
    kept = ctx.kept_mask(t_acts[:2]).cpu().numpy().astype(bool)
    for b in range(2):
        want = np.ones(e, dtype=bool)
        want[acts[b, : lens[b] - 1]] = False
        assert np.array_equal(kept[b], want)
        assert np.array_equal(want, orc.kept_edge_mask(e, acts[b])) if b == 0 else True
    out32 = ctx.reward_batch(t_acts, 0.5, "copy", torch.float32)
    out32_len = ctx.reward_batch(t_acts, 0.5, "copy", torch.float32, lengths=t_lens)
    out32_i32 = ctx.reward_batch(t_acts.to(torch.int32), 0.5, "copy", torch.float32, lengths=t_lens)
    assert torch.equal(out32["reward"], out32_len["reward"]) and torch.equal(out32["reward"], out32_i32["reward"])
    assert np.array_equal(out32["nnz_m"].cpu().numpy(), e - (lens - 1))
    out64 = ctx.reward_batch(t_acts, 0.5, "copy", torch.float64, lengths=t_lens)
    # oracle on two patterns (scipy SpGEMM at n = 1e6)
    a32 = p.a.astype(np.float32)
    w32 = orc.reward_batch_copy(p.n, p.edge_row, p.edge_col, p.edge_val.astype(np.float32), a32, acts[:2], 0.5,
                                dtype=np.float32)
    w64 = orc.reward_batch_copy(p.n, p.edge_row, p.edge_col, p.edge_val, p.a, acts[:2], 0.5, dtype=np.float64)
    np.testing.assert_allclose(out32["reward"][:2].cpu().numpy(), w32["reward"], rtol=1e-4, atol=2e-2)
    np.testing.assert_allclose(out32["residual"][:2].cpu().numpy(), w32["residual"], rtol=1e-4)
    np.testing.assert_allclose(out64["reward"][:2].cpu().numpy(), w64["reward"], rtol=1e-10, atol=1e-8)
    np.testing.assert_allclose(out64["residual"][:2].cpu().numpy(), w64["residual"], rtol=1e-10)
    assert np.array_equal(out64["nnz_m"][:2].cpu().numpy(), w64["nnz_m"])


def test_cfg5_full_size_sparse_kernel_vs_row_sweep_and_oracle(monkeypatch):
    """Short trajectories (0.7 % of the edges): the deletion-driven K3s against the row sweep and the oracle,
    on masks built by K0b at 340 k words per trajectory."""
    p, ctx = _problem("cfg5")
    e = p.num_edges
    acts, lens = _spec_trajectories(e, list(range(40, 72)), max_frac=0.014)
    t_acts = torch.from_numpy(acts).cuda()
    monkeypatch.setenv("SPAI_K3_SPARSE", "1")
    s32 = ctx.reward_batch(t_acts, 0.5, "copy", torch.float32)
    s64 = ctx.reward_batch(t_acts, 0.5, "copy", torch.float64)
    monkeypatch.setenv("SPAI_K3_SPARSE", "0")
    r32 = ctx.reward_batch(t_acts, 0.5, "copy", torch.float32)
    r64 = ctx.reward_batch(t_acts, 0.5, "copy", torch.float64)
    monkeypatch.delenv("SPAI_K3_SPARSE")
    assert torch.equal(s32["nnz_m"], r32["nnz_m"])
    np.testing.assert_allclose(s32["reward"].cpu().numpy(), r32["reward"].cpu().numpy(), rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(s64["reward"].cpu().numpy(), r64["reward"].cpu().numpy(), rtol=1e-11, atol=1e-9)
    w64 = orc.reward_batch_copy(p.n, p.edge_row, p.edge_col, p.edge_val, p.a, acts[:1], 0.5, dtype=np.float64)
    np.testing.assert_allclose(s64["reward"][:1].cpu().numpy(), w64["reward"], rtol=1e-10, atol=1e-8)
    np.testing.assert_allclose(r64["reward"][:1].cpu().numpy(), w64["reward"], rtol=1e-10, atol=1e-8)


@pytest.mark.parametrize("cfg", ["cfg3", "cfg4"])
def test_full_size_ls_and_ls_gram_vs_lapack_on_sampled_rows(cfg):
    """100 blocks of 20 consecutive rows (2000 rows) of one dense trajectory: sum of squared
    least-squares row residuals from the row-range entry point against LAPACK lstsq, fp64 1e-10,
    fp32 1e-4 ('parity unpinned': the reference has no ls code; LAPACK restatement)."""
    p, ctx = _problem(cfg)
    e = p.num_edges
    acts, lens = _spec_trajectories(e, [7])
    t_acts = torch.from_numpy(acts).cuda()
    kept = np.ones(e, dtype=bool)
    kept[acts[0, : lens[0] - 1]] = False
    pat = orc.build_pattern_matrix(p.n, p.edge_row, p.edge_col, np.ones(e), kept, np.float64)
    rng = np.random.default_rng(17)
    starts = np.sort(rng.choice(p.n - 20, size=100, replace=False))
    want = np.empty(starts.size)
    for q, s in enumerate(starts):
        tot = 0.0
        for i in range(int(s), int(s) + 20):
            tot += orc.ls_row_residual2(p.a, i, np.unique(pat.indices[pat.indptr[i]:pat.indptr[i + 1]]))
        want[q] = tot
    for mode in ("ls", "ls_gram"):
        for dt, tol in ((torch.float64, 1e-10), (torch.float32, 1e-4)):
            got = np.empty(starts.size)
            for q, s in enumerate(starts):
                res2, _ = ctx.reward_rows(t_acts, int(s), int(s) + 20, mode, dt)
                got[q] = float(res2[0])
            np.testing.assert_allclose(got, want, rtol=tol, atol=tol * 1e-2, err_msg=f"{cfg} {mode} {dt}")


@pytest.mark.parametrize("a", [524_281, 8_374_109])
def test_k4_sample_step_at_config_scale(a):
    """K4 at the action counts of cfg2 and cfg4: 64-bit span arithmetic and the fp32 sum over
    millions of terms against the fp64 inverse-CDF oracle."""
    from gflownet_spai_b200.env import SpaiContext
    if "k4" not in _CACHE:
        q = synth.make_problem("cfg2", 0.05)
        coo = q.a.tocoo()
        _CACHE["k4"] = (q, SpaiContext(q.n, q.edge_row, q.edge_col, q.edge_val, coo.row, coo.col, coo.data, device=0))
    ctx = _CACHE["k4"][1]
    rng = np.random.default_rng(a)
    bsz = 24
    logits = rng.normal(scale=1.5, size=a).astype(np.float32)
    words = (a + 31) // 32
    taken_np = np.zeros((bsz, words), dtype=np.uint32)
    taken_lists = []
    for b in range(bsz):
        cnt = int(rng.integers(0, a // 2)) if b else a - 2            # sample 0: two ids left
        ids = rng.permutation(a - 1)[:cnt] if b else np.setdiff1d(np.arange(a - 1), [a // 3])
        flags = np.zeros(words * 32, dtype=bool)
        flags[ids] = True
        taken_np[b] = np.packbits(flags, bitorder="little").view(np.uint32)
        taken_lists.append(ids)
    u = rng.random(bsz).astype(np.float32)
    taken = torch.from_numpy(taken_np.view(np.int32)).cuda()
    done = torch.zeros(bsz, dtype=torch.uint8, device="cuda")
    act = torch.empty(bsz, dtype=torch.int64, device="cuda")
    prob = torch.empty(bsz, dtype=torch.float32, device="cuda")
    ctx.sample_step(torch.from_numpy(logits).cuda(), taken, torch.from_numpy(u).cuda(), done, act, prob)
    act, prob, done = act.cpu().numpy(), prob.cpu().numpy(), done.cpu().numpy().astype(bool)
    l64 = logits.astype(np.float64)
    for b in range(bsz):
        w = np.exp(l64 - l64.max())
        w[taken_lists[b]] = 0.0
        cdf = np.cumsum(w)
        x = int(act[b])
        assert 0 <= x < a and w[x] > 0.0
        lo = cdf[x - 1] if x else 0.0
        tgt = float(u[b]) * cdf[-1]
        slack = 2e-5 * cdf[-1]                                        # fp32 partial sums over up to 8.4 M terms
        assert lo - slack <= tgt <= cdf[x] + slack
        assert prob[b] == pytest.approx(w[x] / cdf[-1], rel=5e-4, abs=1e-9)
        assert done[b] == (x == a - 1)
        assert (int(taken[b, x // 32].item()) >> (x % 32)) & 1
    assert int(act[0]) in (a // 3, a - 1)


@pytest.mark.parametrize("case", ["zero_res0", "zero_flops0"])
def test_zero_baselines_give_inf_ratios_like_the_reference(case):
    """preconditioner.py:154 (res0 == 0) and :158 (flops0 == 0): the live reference returns -inf
    rewards (goldens made with a tensor alpha, as the sampler passes it)."""
    from gflownet_spai_b200.env import PreconditionerEnv, SpaiContext
    g = load_golden(case)
    n = int(g["n"])
    ctx = SpaiContext(n, g["edge_row"], g["edge_col"], g["edge_val"].astype(np.float64),
                      g["a_row"], g["a_col"], g["a_val"].astype(np.float64), device=0)
    info = ctx.info()
    assert info.orig_flops == int(g["orig_flops"])
    assert info.orig_residual_f32 == pytest.approx(float(g["orig_residual"]), abs=1e-12)
    assert np.all(np.isneginf(g["reward"]))
    for a_in in (torch.from_numpy(g["actions"]), torch.from_numpy(g["actions"]).cuda()):
        for mode in (("copy", "ls_gram") if case == "zero_res0" else ("copy",)):
            got = ctx.reward_batch(a_in, 0.5, mode, torch.float32)["reward"].cpu().numpy()
            assert np.array_equal(got, g["reward"]), (case, mode, got)
    ctx.close()
    init = torch.sparse_coo_tensor(torch.tensor(np.stack([g["edge_row"], g["edge_col"]])),
                                   torch.tensor(g["edge_val"]), (n, n))
    orig = torch.sparse_coo_tensor(torch.tensor(np.stack([g["a_row"], g["a_col"]]).reshape(2, -1)),
                                   torch.tensor(g["a_val"]), (n, n))
    env = PreconditionerEnv(n, init, orig, device=0)
    rewards = env.update([init] * 3, torch.from_numpy(g["actions"]), torch.tensor(0.5))
    assert all(bool(torch.isneginf(r)) for r in rewards)
    assert bool(torch.isneginf(env.reward(init, 0, torch.tensor(0.5))))
    env.ctx.close()


def test_create_mask_from_sparse_matrix_content_matches_reference_semantics():
    """preconditioner.py:101-135: dense n*n 0/1 mask of the stored coordinates, flattened row-major,
    plus a trailing 1 (the terminal action)."""
    from gflownet_spai_b200.env import PreconditionerEnv
    g = load_golden("uncoalesced40")
    n = int(g["n"])
    init = torch.sparse_coo_tensor(torch.tensor(np.stack([g["edge_row"], g["edge_col"]])),
                                   torch.tensor(g["edge_val"]), (n, n))
    orig = torch.sparse_coo_tensor(torch.tensor(np.stack([g["a_row"], g["a_col"]])), torch.tensor(g["a_val"]), (n, n))
    env = PreconditionerEnv(n, init, orig, device=0)
    m = env.create_mask_from_sparse_matrix(init)
    want = np.zeros(n * n + 1, dtype=np.float32)
    want[g["edge_row"] * n + g["edge_col"]] = 1.0
    want[-1] = 1.0
    assert m.shape == (1, n * n + 1) and m.dtype == torch.float32
    assert np.array_equal(m.numpy().ravel(), want)
    assert torch.equal(env.mask([init, init]), torch.ones(2, env.num_actions))
    env.ctx.close()


def teardown_module(module):
    for _, ctx in _CACHE.values():
        ctx.close()
    _CACHE.clear()
