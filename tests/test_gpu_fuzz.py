"""Differential test over random small problems: whatever kernel the library picks
(row sweep, table lookup, deletion-driven; Gram / Householder) must agree with the
oracle. Random sparse A, random candidate supersets with 1..40 candidates per row,
repeated coordinates, caller-order edges, short and dense trajectories, batch sizes on
both sides of the table threshold."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import spai_oracle as orc

pytestmark = pytest.mark.gpu


def _problem(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(40, 260))
    dens = float(rng.uniform(0.02, 0.08))
    a = sp.random(n, n, density=dens, random_state=int(rng.integers(1 << 30)), format="csr")
    a = sp.csr_matrix(a + sp.identity(n) * float(rng.uniform(1.5, 4.0)))
    a.sort_indices()
    kmax = int(rng.choice([3, 8, 8, 16, 40]))
    rows, cols = [], []
    for i in range(n):
        k = int(rng.integers(1, kmax + 1))
        cc = rng.choice(n, size=min(k, n), replace=False)
        if rng.random() < 0.05 and cc.size > 1:
            cc = np.concatenate([cc, cc[:1]])                 # a repeated coordinate
        rows.append(np.full(cc.size, i))
        cols.append(cc)
    r = np.concatenate(rows).astype(np.int64)
    c = np.concatenate(cols).astype(np.int64)
    v = rng.uniform(-1, 1, r.size)
    perm = rng.permutation(r.size)
    return n, a, r[perm], c[perm], v[perm], rng


def _trajectories(rng, e, bsz, frac):
    t = max(2, int(e * frac) + 1)
    acts = np.full((bsz, t), -1, dtype=np.int64)
    for b in range(bsz):
        ln = int(rng.integers(0, t))
        ids = rng.integers(0, e + 3, size=ln)                # duplicates and ids >= E on purpose
        acts[b, :ln] = ids
        if ln < t:
            acts[b, ln] = e                                   # terminal id
    return acts


@pytest.mark.parametrize("seed", range(12))
def test_random_problems_match_the_oracle(seed):
    from gflownet_spai_b200.env import SpaiContext
    n, a, r, c, v, rng = _problem(1000 + seed)
    coo = a.tocoo()
    ctx = SpaiContext(n, r, c, v, coo.row, coo.col, coo.data)
    e = r.size
    for bsz, frac in ((5, 0.5), (70, 0.5), (70, 0.02), (9, 0.02)):
        acts = _trajectories(rng, e, bsz, frac)
        t = torch.from_numpy(acts).cuda()
        sel = [0, bsz // 2, bsz - 1]
        want = orc.reward_batch_copy(n, r, c, v, a, acts[sel], 0.4, dtype=np.float64)
        got = ctx.reward_batch(t, 0.4, "copy", torch.float64)
        np.testing.assert_allclose(got["reward"].cpu().numpy()[sel], want["reward"], rtol=1e-10, atol=1e-8)
        assert np.array_equal(got["nnz_m"].cpu().numpy()[sel], want["nnz_m"])
        want32 = orc.reward_batch_copy(n, r, c, v.astype(np.float32), a.astype(np.float32), acts[sel], 0.4, dtype=np.float32)
        got32 = ctx.reward_batch(t, 0.4, "copy", torch.float32)
        np.testing.assert_allclose(got32["reward"].cpu().numpy()[sel], want32["reward"], rtol=1e-4, atol=5e-2)
        host = ctx.reward_batch(torch.from_numpy(acts), 0.4, "copy", torch.float64)
        assert torch.equal(host["reward"], got["reward"].cpu())
        # least squares: Gram route against the Householder route, and one pattern against LAPACK
        g = ctx.reward_batch(t, 0.4, "ls_gram", torch.float64)
        q = ctx.reward_batch(t, 0.4, "ls", torch.float64)
        assert torch.allclose(g["residual"], q["residual"], rtol=1e-8, atol=1e-9)
        wls = orc.reward_batch_ls(n, r, c, a, acts[:1], 0.4, dtype=np.float64, baseline_dtype=np.float64)
        np.testing.assert_allclose(g["residual"].cpu().numpy()[:1], wls["residual"], rtol=1e-8, atol=1e-9)
        assert torch.all(g["residual"] <= got["residual"] + 1e-8)        # LS optimality
    ctx.close()
