"""Two-pass mask build (K0b, k0b_bucket.cuh), caller-supplied row lengths and int32 ids.

Bit-exactness bar: kept-edge masks equal `i not in set(actions)` (gflownet/utils.py:315-323)
for every trajectory; rewards equal the ones of the plain entry points."""
import numpy as np
import pytest
import torch

from gflownet_spai_b200 import synth
from oracle import spai_oracle as orc

pytestmark = pytest.mark.gpu


def _ctx(p, perm=None):
    from gflownet_spai_b200.env import SpaiContext
    coo = p.a.tocoo()
    r, c, v = p.edge_row, p.edge_col, p.edge_val
    if perm is not None:
        r, c, v = r[perm], c[perm], v[perm]
    return SpaiContext(p.n, r, c, v, coo.row, coo.col, coo.data, device=0)


def _ragged(num_edges, batch, tmax, seed, with_noise=True):
    """actions with duplicates, -1 in the middle, ids >= E, the terminal id, ragged lengths."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, tmax, size=batch)
    lens[0] = 0
    lens[-1] = tmax - 1
    acts = np.full((batch, tmax), -1, dtype=np.int64)
    for b in range(batch):
        t = int(lens[b])
        row = rng.integers(0, num_edges, size=t)
        if with_noise and t > 8:
            row[rng.integers(0, t, size=3)] = -1
            row[rng.integers(0, t, size=2)] = num_edges + 5
            row[rng.integers(0, t)] = row[0]
        acts[b, :t] = row
        acts[b, t] = num_edges
    return acts, (lens + 1).astype(np.int32)


@pytest.mark.parametrize("group", ["0", "1", "3"])
def test_bucket_path_forced_on_goldens_is_bit_exact(golden, monkeypatch, group):
    from gflownet_spai_b200.env import SpaiContext
    g = golden
    ctx = SpaiContext(int(g["n"]), g["edge_row"], g["edge_col"], g["edge_val"].astype(np.float64),
                      g["a_row"], g["a_col"], g["a_val"].astype(np.float64), device=0)
    acts = torch.from_numpy(g["actions"]).cuda()
    ref_mask = ctx.kept_mask(acts).cpu().numpy()
    ref = ctx.reward_batch(acts, float(g["alpha"]), "copy", torch.float32)
    monkeypatch.setenv("SPAI_K0_VARIANT", "bucket")
    if group != "0":
        monkeypatch.setenv("SPAI_K0B_GROUP", group)
    got_mask = ctx.kept_mask(acts).cpu().numpy()
    got = ctx.reward_batch(acts, float(g["alpha"]), "copy", torch.float32)
    host = ctx.reward_batch(torch.from_numpy(g["actions"]), float(g["alpha"]), "copy", torch.float32)
    assert np.array_equal(got_mask, ref_mask)
    for b in range(ref_mask.shape[0]):
        assert np.array_equal(got_mask[b].astype(bool), orc.kept_edge_mask(g["edge_row"].size, g["actions"][b]))
    assert torch.equal(got["nnz_m"], ref["nnz_m"])
    assert torch.equal(got["reward"], ref["reward"])
    assert np.array_equal(host["reward"].numpy(), ref["reward"].cpu().numpy())
    ctx.close()


@pytest.mark.parametrize("shuffle", [False, True])
def test_bucket_path_medium_pattern(shuffle):
    """3-D Poisson 40^3, <= 16 candidates per row: E = 1.0 M > 819 200 slots, so the library takes
    the two-pass build on its own (16 segments, 2 build CTAs per trajectory); with a shuffled
    edge order the ids go through the edge -> slot map."""
    p = synth.make_problem("cfg3", 40 / 64)
    e = p.num_edges
    assert e > 819200
    perm = np.random.default_rng(3).permutation(e) if shuffle else None
    ctx = _ctx(p, perm)
    acts, lens = _ragged(e, 12, 3 * 8192 + 77, seed=5)
    acts[1, :20000] = np.random.default_rng(9).permutation(e)[:20000]       # a long distinct run
    t_acts = torch.from_numpy(acts).cuda()
    kept = ctx.kept_mask(t_acts).cpu().numpy().astype(bool)
    for b in range(acts.shape[0]):
        assert np.array_equal(kept[b], orc.kept_edge_mask(e, acts[b])), f"trajectory {b}"
    out = ctx.reward_batch(t_acts, 0.5, "copy", torch.float32)
    assert np.array_equal(out["nnz_m"].cpu().numpy(), kept.sum(axis=1))
    # the RED path (round 1) must agree bit for bit
    import os
    os.environ["SPAI_K0_VARIANT"] = "red"
    try:
        red = ctx.reward_batch(t_acts, 0.5, "copy", torch.float32)
    finally:
        del os.environ["SPAI_K0_VARIANT"]
    assert torch.equal(red["nnz_m"], out["nnz_m"]) and torch.equal(red["reward"], out["reward"])
    # oracle on two trajectories
    if not shuffle:
        a32 = p.a.astype(np.float32)
        want = orc.reward_batch_copy(p.n, p.edge_row, p.edge_col, p.edge_val.astype(np.float32), a32, acts[:2], 0.5,
                                     dtype=np.float32)
        np.testing.assert_allclose(out["reward"][:2].cpu().numpy(), want["reward"], rtol=1e-4, atol=2e-2)
    ctx.close()


@pytest.mark.parametrize("cfg,scale", [("cfg2", 0.25), ("cfg3", 40 / 64)])
def test_row_lengths_and_int32_ids(cfg, scale):
    """Entries beyond row_len[b] are never read: fill them with VALID ids that would change the
    result. int32 ids give the same masks as int64 ids. Host and device entry points."""
    p = synth.make_problem(cfg, scale)
    e = p.num_edges
    ctx = _ctx(p)
    acts, lens = _ragged(e, 9, 20001, seed=11)
    clean = torch.from_numpy(acts).cuda()
    want = ctx.reward_batch(clean, 0.5, "copy", torch.float32)
    want_mask = ctx.kept_mask(clean).cpu().numpy()
    dirty = acts.copy()
    rng = np.random.default_rng(2)
    for b in range(acts.shape[0]):
        dirty[b, lens[b]:] = rng.integers(0, e, size=acts.shape[1] - lens[b])
    assert not np.array_equal(ctx.kept_mask(torch.from_numpy(dirty).cuda()).cpu().numpy(), want_mask)
    lens_t = torch.from_numpy(lens)
    for ids in (torch.from_numpy(dirty), torch.from_numpy(dirty.astype(np.int32))):
        got_d = ctx.reward_batch(ids.cuda(), 0.5, "copy", torch.float32, lengths=lens_t.cuda())
        got_h = ctx.reward_batch(ids, 0.5, "copy", torch.float32, lengths=lens_t)
        pin = ids.pin_memory()
        got_p = ctx.reward_batch(pin, 0.5, "copy", torch.float32, lengths=lens_t)
        for got in (got_d, got_h, got_p):
            assert np.array_equal(got["nnz_m"].cpu().numpy(), want["nnz_m"].cpu().numpy())
            assert np.array_equal(got["reward"].cpu().numpy(), want["reward"].cpu().numpy())
    # int32 ids without lengths on the device (padding is read and ignored)
    got = ctx.reward_batch(torch.from_numpy(acts.astype(np.int32)).cuda(), 0.5, "copy", torch.float32)
    assert np.array_equal(got["reward"].cpu().numpy(), want["reward"].cpu().numpy())
    with pytest.raises(ValueError):
        ctx.reward_batch(torch.from_numpy(acts.astype(np.int32)), 0.5, "copy", torch.float32)
    ctx.close()


def test_long_rows_pinned_host_with_and_without_lengths():
    """T >= 4096 on pinned memory: the zero-copy route, trimmed by the host scan or by caller lengths."""
    p = synth.make_problem("cfg2", 0.25)
    e = p.num_edges
    ctx = _ctx(p)
    acts = synth.make_trajectories(e, 24, seed0=77)
    lens = torch.from_numpy(((acts >= 0).sum(axis=1)).astype(np.int32))
    dev = ctx.reward_batch(torch.from_numpy(acts).cuda(), 0.5, "copy", torch.float32)["reward"].cpu().numpy()
    pin = torch.from_numpy(acts).pin_memory()
    a = ctx.reward_batch(pin, 0.5, "copy", torch.float32)["reward"].numpy()
    h2d_scan = ctx.last_timing().h2d_bytes
    b = ctx.reward_batch(pin, 0.5, "copy", torch.float32, lengths=lens)["reward"].numpy()
    h2d_len = ctx.last_timing().h2d_bytes
    assert np.array_equal(a, dev) and np.array_equal(b, dev)
    assert h2d_len == h2d_scan == 8.0 * float(lens.sum()) + 4.0 * acts.shape[0]
    ctx.close()


@pytest.mark.parametrize("build,lanes", [("2", "4"), ("2", "2"), ("2", "1"), ("1", "0"), ("3", "0")])
def test_build_pass_versions_on_sorted_and_ragged_rows(build, lanes, monkeypatch):
    """Every version of the K0b build pass (default: 2 with lanes per run from the mean run length) gives the oracle's
    masks on rows that stress the run walk: sorted ids (one segment gets the whole 4096-id chunk: runs far beyond the
    words a lane keeps in flight), sorted with duplicates, descending, ragged noise rows, an empty row."""
    p = synth.make_problem("cfg3", 40 / 64)
    e = p.num_edges
    ctx = _ctx(p)
    tmax = 5 * 8192 + 19
    acts, _ = _ragged(e, 10, tmax, seed=21)
    rng = np.random.default_rng(4)
    acts[1, :] = -1
    acts[1, :tmax - 1] = np.arange(tmax - 1) * 5
    acts[2, :] = -1
    acts[2, :30000] = np.sort(rng.integers(0, e, size=30000))
    acts[3, :] = -1
    acts[3, :25000] = e - 1 - 2 * np.arange(25000)
    acts[4, :] = -1
    monkeypatch.setenv("SPAI_K0_VARIANT", "bucket")
    monkeypatch.setenv("SPAI_K0B_BUILD", build)
    if lanes != "0":
        monkeypatch.setenv("SPAI_K0B_LANES", lanes)
    t_acts = torch.from_numpy(acts).cuda()
    kept = ctx.kept_mask(t_acts).cpu().numpy().astype(bool)
    for b in range(acts.shape[0]):
        assert np.array_equal(kept[b], orc.kept_edge_mask(e, acts[b])), f"trajectory {b}"
    out = ctx.reward_batch(t_acts, 0.5, "copy", torch.float32)
    assert np.array_equal(out["nnz_m"].cpu().numpy(), kept.sum(axis=1))
    ctx.close()
