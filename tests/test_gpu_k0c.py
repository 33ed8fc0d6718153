"""One-pass cluster mask build (K0c, k0c_cluster.cuh): the bitmask of a trajectory sliced over the
shared memory of a thread-block cluster, ids exchanged in bulk through distributed shared memory.

Bit-exactness bar: kept-edge masks equal `i not in set(actions)` (gflownet/utils.py:315-323)
for every trajectory, for every cluster size / group count, including inputs that overflow an
inbox slot (sorted trajectories: every id of a chunk goes to ONE owner, the excess takes the
remote-atomic path)."""
import numpy as np
import pytest
import torch

from gflownet_spai_b200 import synth
from oracle import spai_oracle as orc

from test_gpu_k0b import _ctx, _ragged

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cs,groups", [("2", "2"), ("8", "1"), ("5", "3")])
def test_cluster_path_forced_on_goldens_is_bit_exact(golden, monkeypatch, cs, groups):
    from gflownet_spai_b200.env import SpaiContext
    g = golden
    ctx = SpaiContext(int(g["n"]), g["edge_row"], g["edge_col"], g["edge_val"].astype(np.float64),
                      g["a_row"], g["a_col"], g["a_val"].astype(np.float64), device=0)
    acts = torch.from_numpy(g["actions"]).cuda()
    ref_mask = ctx.kept_mask(acts).cpu().numpy()
    ref = ctx.reward_batch(acts, float(g["alpha"]), "copy", torch.float32)
    monkeypatch.setenv("SPAI_K0_VARIANT", "cluster")
    monkeypatch.setenv("SPAI_K0C_CS", cs)
    monkeypatch.setenv("SPAI_K0C_GROUPS", groups)
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


@pytest.mark.parametrize("shuffle,cs,groups", [(False, "0", "2"), (True, "0", "2"), (False, "8", "3"), (True, "3", "1"),
                                               (False, "7", "2")])
def test_cluster_path_medium_pattern(shuffle, cs, groups, monkeypatch):
    """3-D Poisson 40^3, <= 16 candidates per row: E = 1.0 M > 819 200 slots = 16 slices of 64 K slots; with
    SPAI_K0_VARIANT=cluster the library picks the cluster geometry itself (cs = 0: smallest cluster that fits). Ragged rows with
    duplicates / -1 / ids >= E / the terminal id, one long distinct run, one long SORTED run (every chunk
    lands on one owner: inbox slots overflow into the remote-atomic path), one sorted-descending run, an
    empty row. With a shuffled edge order the ids go through the edge -> slot map."""
    p = synth.make_problem("cfg3", 40 / 64)
    e = p.num_edges
    assert e > 819200
    perm = np.random.default_rng(3).permutation(e) if shuffle else None
    ctx = _ctx(p, perm)
    tmax = 9 * 8192 + 77
    acts, lens = _ragged(e, 14, tmax, seed=5)
    rng = np.random.default_rng(9)
    acts[1, :] = -1
    acts[1, :60000] = rng.permutation(e)[:60000]                    # long distinct run
    acts[2, :] = -1
    acts[2, :tmax - 1] = np.arange(tmax - 1) * 3                    # sorted: chunks of 2048 ids hit one owner
    acts[3, :] = -1
    acts[3, :50000] = e - 1 - np.arange(50000)                      # descending, the last slice
    acts[4, :] = -1                                                 # empty
    acts[5, :] = -1
    acts[5, :70000] = np.sort(rng.integers(0, e, size=70000))       # sorted with duplicates
    t_acts = torch.from_numpy(acts).cuda()
    want = [orc.kept_edge_mask(e, acts[b]) for b in range(acts.shape[0])]
    monkeypatch.setenv("SPAI_K0_VARIANT", "bucket")
    ref = ctx.reward_batch(t_acts, 0.5, "copy", torch.float32)
    monkeypatch.delenv("SPAI_K0_VARIANT")
    monkeypatch.setenv("SPAI_K0_VARIANT", "cluster")
    if cs != "0":
        monkeypatch.setenv("SPAI_K0C_CS", cs)
    monkeypatch.setenv("SPAI_K0C_GROUPS", groups)
    kept = ctx.kept_mask(t_acts).cpu().numpy().astype(bool)
    for b in range(acts.shape[0]):
        assert np.array_equal(kept[b], want[b]), f"trajectory {b}"
    out = ctx.reward_batch(t_acts, 0.5, "copy", torch.float32)
    assert np.array_equal(out["nnz_m"].cpu().numpy(), kept.sum(axis=1))
    assert torch.equal(ref["nnz_m"], out["nnz_m"]) and torch.equal(ref["reward"], out["reward"])
    # int32 ids + row lengths, device and pinned-host (zero-copy) entry points
    lens_all = torch.from_numpy(np.where(acts >= 0, np.arange(tmax)[None, :] + 1, 0).max(axis=1).astype(np.int32))
    got32 = ctx.reward_batch(t_acts.to(torch.int32), 0.5, "copy", torch.float32, lengths=lens_all.cuda())
    assert torch.equal(got32["reward"], out["reward"])
    pin = torch.from_numpy(acts).pin_memory()
    goth = ctx.reward_batch(pin, 0.5, "copy", torch.float32, lengths=lens_all)
    assert np.array_equal(goth["reward"].numpy(), out["reward"].cpu().numpy())
    ctx.close()
