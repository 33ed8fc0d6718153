"""ls_gram mode (K2g, semi-normal equations) against the ls oracle (LAPACK lstsq on
the gathered tile) and against the Householder kernels: same least-squares
residual, so the same tolerances as ls mode apply (fp64 1e-10, fp32 1e-4)."""
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


@pytest.mark.parametrize("name", ["poisson10", "convdiff16", "uncoalesced40", "tiny3", "poisson32_k8"])
def test_gram_matches_oracle_on_goldens(name):
    import conftest
    from gflownet_spai_b200.env import SpaiContext
    g = conftest.load_golden(name)
    n = int(g["n"])
    ctx = SpaiContext(n, g["edge_row"], g["edge_col"], g["edge_val"].astype(np.float64),
                      g["a_row"], g["a_col"], g["a_val"].astype(np.float64), device=0)
    a64 = sp.coo_matrix((g["a_val"].astype(np.float64), (g["a_row"], g["a_col"])), shape=(n, n)).tocsr()
    a64.sum_duplicates()
    a64.sort_indices()
    acts = g["actions"][:6]
    t = torch.from_numpy(acts).cuda()
    alpha = float(g["alpha"])
    want = orc.reward_batch_ls(n, g["edge_row"], g["edge_col"], a64, acts, alpha, dtype=np.float64,
                               a_stored_nnz=g["a_val"].size, baseline_dtype=np.float64)
    out = ctx.reward_batch(t, alpha, "ls_gram", torch.float64)
    np.testing.assert_allclose(out["residual"].cpu().numpy(), want["residual"], rtol=1e-10, atol=1e-8)
    np.testing.assert_allclose(out["reward"].cpu().numpy(), want["reward"], rtol=1e-10, atol=1e-7)
    assert np.array_equal(out["nnz_m"].cpu().numpy(), want["nnz_m"])
    want32 = orc.reward_batch_ls(n, g["edge_row"], g["edge_col"], a64, acts, alpha, dtype=np.float32,
                                 a_stored_nnz=g["a_val"].size, baseline_dtype=np.float32)
    out32 = ctx.reward_batch(t, alpha, "ls_gram", torch.float32)
    np.testing.assert_allclose(out32["reward"].cpu().numpy(), want32["reward"], rtol=1e-4, atol=2e-2)
    ctx.close()


@pytest.mark.parametrize("cfg,scale", [("cfg2", 0.125), ("cfg3", 0.1875), ("cfg4", 0.046875), ("cfg5", 0.0015)])
def test_gram_matches_oracle_and_qr_on_scaled_configs(cfg, scale):
    """k <= 8 rows (cfg2) are all Gram rows; cfg3 (k <= 16) uses the 16-wide Gram class in
    fp32 and the QR column kernel in fp64; cfg4 (k <= 32) stays on QR except boundary rows;
    cfg5 mixes everything incl. generic rows."""
    p = synth.make_problem(cfg, scale)
    ctx = _ctx(p)
    acts = synth.make_trajectories(p.num_edges, 40, seed0=77)
    t = torch.from_numpy(acts).cuda()
    wls = orc.reward_batch_ls(p.n, p.edge_row, p.edge_col, p.a, acts[:2], 0.5, dtype=np.float64,
                              baseline_dtype=np.float64)
    g64 = ctx.reward_batch(t, 0.5, "ls_gram", torch.float64)
    np.testing.assert_allclose(g64["residual"].cpu().numpy()[:2], wls["residual"], rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(g64["reward"].cpu().numpy()[:2], wls["reward"], rtol=1e-10, atol=1e-7)
    q64 = ctx.reward_batch(t, 0.5, "ls", torch.float64)
    assert torch.allclose(g64["residual"], q64["residual"], rtol=1e-10, atol=1e-9)
    assert torch.equal(g64["nnz_m"], q64["nnz_m"])
    wls32 = orc.reward_batch_ls(p.n, p.edge_row, p.edge_col, p.a, acts[:2], 0.5, dtype=np.float32,
                                baseline_dtype=np.float32)
    g32 = ctx.reward_batch(t, 0.5, "ls_gram", torch.float32)
    np.testing.assert_allclose(g32["reward"].cpu().numpy()[:2], wls32["reward"], rtol=1e-4, atol=2e-2)
    assert torch.allclose(g32["residual"], q64["residual"], rtol=1e-4)
    ctx.close()


def test_gram_hands_dependent_columns_to_householder():
    """Two identical rows of A make some tiles rank-deficient: the relative pivot test
    must route those (row, trajectory) pairs to the generic Householder kernel."""
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
    acts = synth.make_trajectories(r.size, 6, seed0=2, max_frac=0.2)
    acts[0, :] = -1
    want = orc.reward_batch_ls(n, r, c, a, acts, 0.5, dtype=np.float64, baseline_dtype=np.float64)
    got = ctx.reward_batch(torch.from_numpy(acts).cuda(), 0.5, "ls_gram", torch.float64)
    np.testing.assert_allclose(got["residual"].cpu().numpy(), want["residual"], rtol=1e-9, atol=1e-9)
    want32 = orc.reward_batch_ls(n, r, c, a, acts, 0.5, dtype=np.float32, baseline_dtype=np.float32)
    got32 = ctx.reward_batch(torch.from_numpy(acts).cuda(), 0.5, "ls_gram", torch.float32)
    np.testing.assert_allclose(got32["reward"].cpu().numpy(), want32["reward"], rtol=1e-4, atol=5e-2)
    ctx.close()


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-6), (torch.float64, 1e-12)])
def test_gram_row_ranges_and_entry_points_agree(dtype, tol):
    from gflownet_spai_b200.dist import shard_bounds
    p = synth.make_problem("cfg5", 0.004)
    ctx = _ctx(p)
    acts_h = torch.from_numpy(synth.make_trajectories(p.num_edges, 9, seed0=4))
    acts = acts_h.cuda()
    full = ctx.reward_batch(acts, 0.3, "ls_gram", dtype)
    tot = torch.zeros(9, dtype=torch.float64, device="cuda")
    for r in range(3):
        lo, hi = shard_bounds(p.n, 3, r)
        part, nnz = ctx.reward_rows(acts, lo, hi, "ls_gram", dtype)
        tot += part
    fin = ctx.finalize_rewards(tot, nnz, 0.3, dtype)
    assert torch.allclose(fin["reward"], full["reward"], rtol=tol, atol=1e-6)
    host = ctx.reward_batch(acts_h, 0.3, "ls_gram", dtype)
    assert torch.equal(host["reward"], full["reward"].cpu())
    ctx.close()


def test_gram_short_trajectories_and_full_size_cfg2():
    """Full-size headline matrix: Gram and Householder agree pattern by pattern (fp64 1e-10),
    the ls residual never exceeds the copy residual, and the untouched-row path (few
    deletions) gives the same numbers as the solve path."""
    p = synth.make_problem("cfg2")
    ctx = _ctx(p)
    acts = synth.make_trajectories(p.num_edges, 64, seed0=900)
    t = torch.from_numpy(acts).cuda()
    g64 = ctx.reward_batch(t, 0.5, "ls_gram", torch.float64)
    q64 = ctx.reward_batch(t, 0.5, "ls", torch.float64)
    assert torch.allclose(g64["residual"], q64["residual"], rtol=1e-10, atol=1e-10)
    assert torch.allclose(g64["reward"], q64["reward"], rtol=1e-10, atol=1e-7)
    cp = ctx.reward_batch(t, 0.5, "copy", torch.float64)
    assert torch.all(g64["residual"] <= cp["residual"] + 1e-9)
    g32 = ctx.reward_batch(t, 0.5, "ls_gram", torch.float32)
    q32 = ctx.reward_batch(t, 0.5, "ls", torch.float32)
    assert torch.allclose(g32["reward"], q32["reward"], rtol=1e-4, atol=2e-2)
    short = torch.from_numpy(synth.make_trajectories(p.num_edges, 64, seed0=5, max_frac=0.0002)).cuda()
    gs = ctx.reward_batch(short, 0.5, "ls_gram", torch.float64)
    qs = ctx.reward_batch(short, 0.5, "ls", torch.float64)
    assert torch.allclose(gs["residual"], qs["residual"], rtol=1e-10, atol=1e-10)
    ctx.close()
