"""Parity tests proper: the CUDA path, called through the C ABI, against the
reference's golden outputs and the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north star): masks / index sets bit-exact; fp64 reward
1e-10 relative; fp32 reward 1e-4 relative (atol covers rewards that are ~0 on a
scale of 1000)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from gflownet_spai_b200 import synth
from oracle import spai_oracle as orc

pytestmark = pytest.mark.gpu

RTOL32, ATOL32 = 1e-4, 2e-2
RTOL64, ATOL64 = 1e-10, 1e-8


def _ctx_from_golden(g):
    from gflownet_spai_b200.env import SpaiContext
    return SpaiContext(int(g["n"]), g["edge_row"], g["edge_col"], g["edge_val"].astype(np.float64),
                       g["a_row"], g["a_col"], g["a_val"].astype(np.float64), device=0)


def _a_csr(g, dtype):
    n = int(g["n"])
    a = sp.coo_matrix((g["a_val"].astype(dtype), (g["a_row"], g["a_col"])), shape=(n, n)).tocsr()
    a.sum_duplicates()
    a.sort_indices()
    return a


def test_copy_fp32_matches_reference_golden(golden):
    g = golden
    ctx = _ctx_from_golden(g)
    info = ctx.info()
    assert info.init_nnz == int(g["init_nnz"])
    assert info.num_actions == int(g["num_actions"])
    assert info.orig_flops == int(g["orig_flops"])
    assert info.orig_residual_f32 == pytest.approx(float(g["orig_residual"]), rel=1e-6)
    acts = torch.from_numpy(g["actions"])
    for a_in in (acts, acts.cuda()):
        out = ctx.reward_batch(a_in, float(g["alpha"]), "copy", torch.float32)
        np.testing.assert_allclose(out["reward"].cpu().numpy(), g["reward"], rtol=RTOL32, atol=ATOL32)
    ctx.close()


def test_kept_masks_and_pattern_indices_bit_exact(golden):
    g = golden
    ctx = _ctx_from_golden(g)
    n = int(g["n"])
    kept = ctx.kept_mask(torch.from_numpy(g["actions"])).cpu().numpy().astype(bool)
    nnz = ctx.reward_batch(torch.from_numpy(g["actions"]).cuda(), 0.5)["nnz_m"].cpu().numpy()
    for b in range(kept.shape[0]):
        assert np.array_equal(kept[b], orc.kept_edge_mask(g["edge_row"].size, g["actions"][b]))
        # coalesced (row, col) of M implied by the mask == the reference's resize_sparse_tensor output
        key = np.unique(g["edge_row"][kept[b]] * n + g["edge_col"][kept[b]])
        lo, hi = int(g["m_ptr"][b]), int(g["m_ptr"][b + 1])
        assert np.array_equal(key // n, g["m_row"][lo:hi])
        assert np.array_equal(key % n, g["m_col"][lo:hi])
        assert nnz[b] == hi - lo
    ctx.close()


def test_copy_fp64_matches_oracle(golden):
    g = golden
    ctx = _ctx_from_golden(g)
    n = int(g["n"])
    want = orc.reward_batch_copy(n, g["edge_row"], g["edge_col"], g["edge_val"].astype(np.float64),
                                 _a_csr(g, np.float64), g["actions"], float(g["alpha"]), dtype=np.float64,
                                 a_stored_nnz=g["a_val"].size)
    out = ctx.reward_batch(torch.from_numpy(g["actions"]).cuda(), float(g["alpha"]), "copy", torch.float64)
    np.testing.assert_allclose(out["residual"].cpu().numpy(), want["residual"], rtol=RTOL64, atol=ATOL64)
    np.testing.assert_allclose(out["reward"].cpu().numpy(), want["reward"], rtol=RTOL64, atol=ATOL64)
    assert np.array_equal(out["nnz_m"].cpu().numpy(), want["nnz_m"])
    assert ctx.info().orig_residual_f64 == pytest.approx(want["orig_residual"], rel=1e-12)
    ctx.close()


@pytest.mark.parametrize("name", ["poisson10", "convdiff16", "uncoalesced40", "tiny3"])
def test_ls_matches_oracle(name):
    import conftest
    g = conftest.load_golden(name)
    ctx = _ctx_from_golden(g)
    n = int(g["n"])
    a64 = _a_csr(g, np.float64)
    acts = g["actions"][:6]
    want = orc.reward_batch_ls(n, g["edge_row"], g["edge_col"], a64, acts, float(g["alpha"]), dtype=np.float64,
                               a_stored_nnz=g["a_val"].size, baseline_dtype=np.float64)
    out = ctx.reward_batch(torch.from_numpy(acts).cuda(), float(g["alpha"]), "ls", torch.float64)
    np.testing.assert_allclose(out["residual"].cpu().numpy(), want["residual"], rtol=RTOL64, atol=ATOL64)
    np.testing.assert_allclose(out["reward"].cpu().numpy(), want["reward"], rtol=RTOL64, atol=1e-7)
    want32 = orc.reward_batch_ls(n, g["edge_row"], g["edge_col"], a64, acts, float(g["alpha"]), dtype=np.float32,
                                 a_stored_nnz=g["a_val"].size, baseline_dtype=np.float32)
    out32 = ctx.reward_batch(torch.from_numpy(acts).cuda(), float(g["alpha"]), "ls", torch.float32)
    np.testing.assert_allclose(out32["reward"].cpu().numpy(), want32["reward"], rtol=RTOL32, atol=ATOL32)
    # LS optimality: never worse than keeping the values
    cp = ctx.reward_batch(torch.from_numpy(acts).cuda(), float(g["alpha"]), "copy", torch.float64)
    assert torch.all(out["residual"] <= cp["residual"] + 1e-9)
    ctx.close()


def test_gathered_index_sets_bit_exact():
    a = synth.convdiff2d(12)
    n = a.shape[0]
    r, c = synth.superset_pattern(a, 8, max_power=2)
    v = synth.neumann_values(a, r, c)
    coo = a.tocoo()
    from gflownet_spai_b200.env import SpaiContext
    ctx = SpaiContext(n, r, c, v, coo.row, coo.col, coo.data)
    pat = sp.csr_matrix((np.ones(r.size), (r, c)), shape=(n, n))
    for i in list(range(0, n, 7)) + [n - 1]:
        j, iset = ctx.row_index_sets(i)
        wj, wi = orc.row_index_sets(pat, a, i)
        assert np.array_equal(j, wj) and np.array_equal(iset, wi)
    info = ctx.info()
    assert info.max_row_slots == 8
    ctx.close()


def _random_problem(n, row_nnz, a_row_nnz, seed):
    rng = np.random.default_rng(seed)
    a = sp.random(n, n, density=a_row_nnz / n, random_state=seed, format="csr") + sp.identity(n) * 3.0
    a = sp.csr_matrix(a)
    a.sort_indices()
    rows, cols = [], []
    for i in range(n):
        k = int(rng.integers(0, row_nnz + 1))
        cc = rng.choice(n, size=min(k, n), replace=False)
        rows.append(np.full(cc.size, i))
        cols.append(cc)
    r = np.concatenate(rows).astype(np.int64)
    c = np.concatenate(cols).astype(np.int64)
    perm = rng.permutation(r.size)
    r, c = r[perm], c[perm]
    v = rng.uniform(-1, 1, r.size)
    return a, r, c, v


@pytest.mark.parametrize("row_nnz,a_row_nnz", [(48, 6), (20, 30), (6, 3)])
def test_wide_rows_and_generic_ls_fallback(row_nnz, a_row_nnz):
    """Rows with > 32 candidates (wide copy path) and union sets beyond the
    register kernels' classes (generic ls kernel); empty rows included."""
    n = 96
    a, r, c, v = _random_problem(n, row_nnz, a_row_nnz, seed=row_nnz)
    coo = a.tocoo()
    from gflownet_spai_b200.env import SpaiContext
    ctx = SpaiContext(n, r, c, v, coo.row, coo.col, coo.data)
    acts = synth.make_trajectories(r.size, 5, seed0=9)
    t = torch.from_numpy(acts).cuda()
    want = orc.reward_batch_copy(n, r, c, v, a, acts, 0.5, dtype=np.float64)
    got = ctx.reward_batch(t, 0.5, "copy", torch.float64)
    np.testing.assert_allclose(got["residual"].cpu().numpy(), want["residual"], rtol=RTOL64, atol=ATOL64)
    assert np.array_equal(got["nnz_m"].cpu().numpy(), want["nnz_m"])
    want32 = orc.reward_batch_copy(n, r, c, v.astype(np.float32), a.astype(np.float32), acts, 0.5, dtype=np.float32)
    got32 = ctx.reward_batch(t, 0.5, "copy", torch.float32)
    np.testing.assert_allclose(got32["reward"].cpu().numpy(), want32["reward"], rtol=RTOL32, atol=ATOL32)
    wls = orc.reward_batch_ls(n, r, c, a, acts, 0.5, dtype=np.float64, baseline_dtype=np.float64)
    gls = ctx.reward_batch(t, 0.5, "ls", torch.float64)
    np.testing.assert_allclose(gls["residual"].cpu().numpy(), wls["residual"], rtol=1e-9, atol=1e-8)
    ctx.close()


def test_taken_bitmask_entry_point_equals_action_entry_point():
    p = synth.make_problem("cfg2", scale=0.125)          # 32 x 32 grid
    coo = p.a.tocoo()
    from gflownet_spai_b200.env import SpaiContext
    ctx = SpaiContext(p.n, p.edge_row, p.edge_col, p.edge_val, coo.row, coo.col, coo.data)
    acts = synth.make_trajectories(p.num_edges, 40, seed0=1)
    e = p.num_edges
    words = (e + 1 + 31) // 32
    taken = np.zeros((acts.shape[0], words), dtype=np.uint32)
    for b in range(acts.shape[0]):
        ids = acts[b][(acts[b] >= 0) & (acts[b] < e)]
        np.bitwise_or.at(taken[b], ids // 32, (np.uint32(1) << (ids % 32).astype(np.uint32)))
    t_taken = torch.from_numpy(taken.view(np.int32)).cuda()
    one = ctx.reward_batch(torch.from_numpy(acts).cuda(), 0.25, "copy", torch.float64)
    two = ctx.reward_from_taken(t_taken, 0.25, "copy", torch.float64)
    assert torch.equal(one["nnz_m"], two["nnz_m"])
    assert torch.allclose(one["reward"], two["reward"], rtol=1e-13, atol=1e-10)
    ctx.close()


def test_chunked_batches_match_single_pass_and_edge_cases():
    p = synth.make_problem("cfg2", scale=0.0625)         # 16 x 16 grid
    coo = p.a.tocoo()
    from gflownet_spai_b200.env import SpaiContext
    ctx = SpaiContext(p.n, p.edge_row, p.edge_col, p.edge_val, coo.row, coo.col, coo.data)
    acts = synth.make_trajectories(p.num_edges, 300, seed0=3)
    full = ctx.reward_batch(torch.from_numpy(acts), 0.5, "copy", torch.float32)
    ctx.set_workspace_limit(1 << 20)                     # forces several trajectory chunks
    chunked = ctx.reward_batch(torch.from_numpy(acts), 0.5, "copy", torch.float32)
    assert torch.allclose(full["reward"], chunked["reward"], rtol=1e-12, atol=1e-9)
    # empty batch / zero-length trajectories / single trajectory
    empty = ctx.reward_batch(torch.zeros((0, 4), dtype=torch.int64), 0.5)
    assert empty["reward"].numel() == 0
    none = ctx.reward_batch(torch.zeros((3, 0), dtype=torch.int64).cuda(), 0.5, "copy", torch.float64)
    allkept = ctx.reward_batch(torch.full((1, 2), -1, dtype=torch.int64).cuda(), 0.5, "copy", torch.float64)
    assert torch.allclose(none["reward"], allkept["reward"].expand(3))
    assert int(allkept["nnz_m"][0]) == p.num_edges
    with pytest.raises(ValueError):
        ctx.reward_batch(torch.zeros((2, 2), dtype=torch.int32), 0.5)
    ctx.close()


def test_drop_in_env_matches_reference_outputs(golden):
    """The PreconditionerEnv mirror: constructor attributes, update(), reward(),
    calculate_residual() read like the reference's (preconditioner.py)."""
    from gflownet_spai_b200.env import PreconditionerEnv
    g = golden
    n = int(g["n"])
    init = torch.sparse_coo_tensor(torch.tensor(np.stack([g["edge_row"], g["edge_col"]])),
                                   torch.tensor(g["edge_val"]), (n, n))
    orig = torch.sparse_coo_tensor(torch.tensor(np.stack([g["a_row"], g["a_col"]])),
                                   torch.tensor(g["a_val"]), (n, n))
    env = PreconditionerEnv(n, init, orig)
    assert env.init_nnz == int(g["init_nnz"]) and env.num_actions == int(g["num_actions"])
    assert env.state_dim == env.init_nnz and env.orig_flops == int(g["orig_flops"])
    assert float(env.orig_residual) == pytest.approx(float(g["orig_residual"]), rel=1e-6)
    assert env.data.edge_index.shape == (2, g["edge_row"].size) and "edge_attr" in env.data
    rewards = env.update([init] * g["actions"].shape[0], torch.from_numpy(g["actions"]), torch.tensor(float(g["alpha"])))
    assert isinstance(rewards, list) and rewards[0].dtype == torch.float64 and rewards[0].dim() == 0
    got = torch.tensor(rewards, dtype=torch.float32).numpy()
    np.testing.assert_allclose(got, g["reward"].astype(np.float32), rtol=RTOL32, atol=ATOL32)
    # reward(s, ...) on an explicit matrix == update() on the trajectory that produces it
    b = 0
    kept = orc.kept_edge_mask(g["edge_row"].size, g["actions"][b])
    s = torch.sparse_coo_tensor(torch.tensor(np.stack([g["edge_row"][kept], g["edge_col"][kept]])),
                                torch.tensor(g["edge_val"][kept]), (n, n)).coalesce()
    r1 = env.reward(s, int(kept.sum()), float(g["alpha"]))
    assert float(r1) == pytest.approx(float(g["reward"][b]), rel=RTOL32, abs=ATOL32)
    assert env.mask([0, 1]).shape == (2, env.num_actions)
    with pytest.raises(ValueError):
        env.create_mask_from_sparse_matrix(torch.eye(3))


@pytest.mark.parametrize("name", ["convdiff16", "uncoalesced40", "poisson10"])
def test_ls_solve_values_reproduce_the_ls_residual(name):
    """spai_ls_solve_values_host: the re-solved M, used as explicit values
    (copy semantics), gives the ls residual; rows agree with LAPACK."""
    import conftest
    g = conftest.load_golden(name)
    ctx = _ctx_from_golden(g)
    n = int(g["n"])
    a64 = _a_csr(g, np.float64)
    b = min(3, g["actions"].shape[0] - 1)
    acts = g["actions"][b]
    vals = ctx.ls_solve_values(acts, torch.float64)
    kept = orc.kept_edge_mask(g["edge_row"].size, acts)
    assert np.all(vals[~kept] == 0.0)
    m = orc.build_pattern_matrix(n, g["edge_row"], g["edge_col"], vals, kept, dtype=np.float64)
    res_from_values = orc.residual_copy(m, a64, dtype=np.float64)
    ls = ctx.reward_batch(torch.from_numpy(acts[None, :]).cuda(), 0.5, "ls", torch.float64)
    assert res_from_values == pytest.approx(float(ls["residual"][0]), rel=1e-9, abs=1e-9)
    # row-wise against LAPACK where the row problem has full column rank
    pat = orc.build_pattern_matrix(n, g["edge_row"], g["edge_col"], np.ones(g["edge_row"].size), kept, np.float64)
    mm = m.tocsr()
    checked = 0
    for i in range(0, n, max(1, n // 25)):
        j = np.unique(pat.indices[pat.indptr[i]:pat.indptr[i + 1]])
        if j.size == 0:
            continue
        sub = a64[j, :]
        iset = np.unique(sub.indices)
        hat = np.asarray(sub[:, iset].todense()).T
        if np.linalg.matrix_rank(hat) < j.size:
            continue
        _, sol = orc.ls_row_residual2(a64, i, j, return_m=True)
        got = np.asarray(mm[i, j].todense()).ravel()
        np.testing.assert_allclose(got, sol, rtol=1e-8, atol=1e-10)
        checked += 1
    assert checked > 0
    ctx.close()


def test_host_entry_trims_padding_exactly():
    """spai_reward_batch_host copies only the prefix that still holds ids (rows
    are scanned back over the -1 padding). -1 in the middle of a row, all-padding
    rows, rows without padding and a leading dimension > T must all score like
    the device entry point."""
    p = synth.make_problem("cfg2", scale=0.25)           # 64 x 64 grid, E = 32 760
    coo = p.a.tocoo()
    from gflownet_spai_b200.env import SpaiContext
    ctx = SpaiContext(p.n, p.edge_row, p.edge_col, p.edge_val, coo.row, coo.col, coo.data)
    rng = np.random.default_rng(5)
    bsz, t = 37, 9000                                    # T >= 4096 -> trimmed path
    acts = np.full((bsz, t), -1, dtype=np.int64)
    for b in range(bsz):
        n_valid = int(rng.integers(0, t))
        acts[b, :n_valid] = rng.integers(0, p.num_edges, n_valid)
        holes = rng.integers(0, max(n_valid, 1), 20)
        acts[b, holes] = -1                              # -1 anywhere is legal and ignored
    acts[3, :] = -1                                      # nothing removed
    acts[4, :] = rng.integers(0, p.num_edges, t)         # no padding at all
    acts[5, -1] = 7                                      # a real id in the very last slot
    wide = torch.full((bsz, t + 13), -1, dtype=torch.int64)
    wide[:, :t] = torch.from_numpy(acts)
    dev = ctx.reward_batch(torch.from_numpy(acts).cuda(), 0.5, "copy", torch.float64)
    for host_in in (torch.from_numpy(acts), wide[:, :t]):            # contiguous and ld > T
        host = ctx.reward_batch(host_in, 0.5, "copy", torch.float64)
        assert torch.equal(host["nnz_m"], dev["nnz_m"].cpu())
        assert torch.equal(host["reward"], dev["reward"].cpu())
    kept = orc.kept_edge_mask(p.num_edges, acts[5])
    assert not kept[7] and int(dev["nnz_m"][3]) == p.num_edges
    ctx.close()


def test_rank_deficient_tiles_are_handed_to_the_generic_kernel():
    """The column-per-lane ls kernel assumes full column rank; a singular A
    (two identical rows) makes some tiles rank-deficient: those (row, trajectory)
    pairs must be redone by the generic kernel and still match LAPACK."""
    rng = np.random.default_rng(11)
    n = 64
    a = sp.random(n, n, density=0.08, random_state=4, format="lil") + sp.identity(n, format="lil") * 2.0
    a = sp.lil_matrix(a)
    a[5, :] = a[3, :]                                   # rows 3 and 5 identical -> dependent columns
    a = sp.csr_matrix(a)
    a.sort_indices()
    rows, cols = [], []
    for i in range(n):
        base = [3, 5] if i % 3 == 0 else []
        extra = rng.choice([c for c in range(n) if c not in (3, 5)], size=10, replace=False)
        cc = np.array(base + list(extra))
        rows.append(np.full(cc.size, i))
        cols.append(cc)
    r = np.concatenate(rows).astype(np.int64)
    c = np.concatenate(cols).astype(np.int64)
    v = rng.uniform(-1, 1, r.size)
    coo = a.tocoo()
    from gflownet_spai_b200.env import SpaiContext
    ctx = SpaiContext(n, r, c, v, coo.row, coo.col, coo.data)
    info = ctx.info()
    assert info.ls_class_rows[3] + info.ls_class_rows[4] > 0        # column-per-lane classes in use
    acts = synth.make_trajectories(r.size, 6, seed0=2, max_frac=0.2)
    acts[0, :] = -1                                                  # everything kept: rows 0,3,6.. are deficient
    want = orc.reward_batch_ls(n, r, c, a, acts, 0.5, dtype=np.float64, baseline_dtype=np.float64)
    got = ctx.reward_batch(torch.from_numpy(acts).cuda(), 0.5, "ls", torch.float64)
    np.testing.assert_allclose(got["residual"].cpu().numpy(), want["residual"], rtol=1e-9, atol=1e-9)
    want32 = orc.reward_batch_ls(n, r, c, a, acts, 0.5, dtype=np.float32, baseline_dtype=np.float32)
    got32 = ctx.reward_batch(torch.from_numpy(acts).cuda(), 0.5, "ls", torch.float32)
    np.testing.assert_allclose(got32["reward"].cpu().numpy(), want32["reward"], rtol=1e-4, atol=5e-2)
    ctx.close()


def test_short_trajectories_take_the_untouched_row_path():
    """Incremental evaluation: rows none of whose slots was removed by any
    trajectory of a warp are not re-evaluated (their all-kept residual is cached
    per context). Few deletions per trajectory => almost every row is skipped."""
    p = synth.make_problem("cfg2", scale=0.25)           # 64 x 64 grid
    coo = p.a.tocoo()
    from gflownet_spai_b200.env import SpaiContext
    ctx = SpaiContext(p.n, p.edge_row, p.edge_col, p.edge_val, coo.row, coo.col, coo.data)
    rng = np.random.default_rng(8)
    bsz = 40
    acts = np.full((bsz, 9), -1, dtype=np.int64)
    for b in range(bsz):
        n_del = int(rng.integers(0, 6))
        acts[b, :n_del] = rng.integers(0, p.num_edges, n_del)
        acts[b, n_del] = p.num_edges
    t = torch.from_numpy(acts).cuda()
    want64 = orc.reward_batch_copy(p.n, p.edge_row, p.edge_col, p.edge_val, p.a, acts, 0.5, dtype=np.float64)
    got64 = ctx.reward_batch(t, 0.5, "copy", torch.float64)
    np.testing.assert_allclose(got64["residual"].cpu().numpy(), want64["residual"], rtol=RTOL64, atol=ATOL64)
    want32 = orc.reward_batch_copy(p.n, p.edge_row, p.edge_col, p.edge_val.astype(np.float32),
                                   p.a.astype(np.float32), acts, 0.5, dtype=np.float32)
    got32 = ctx.reward_batch(t, 0.5, "copy", torch.float32)
    np.testing.assert_allclose(got32["reward"].cpu().numpy(), want32["reward"], rtol=RTOL32, atol=ATOL32)
    # the ls kernels skip untouched (row, warp) tiles the same way
    wls = orc.reward_batch_ls(p.n, p.edge_row, p.edge_col, p.a, acts[:5], 0.5, dtype=np.float64,
                              baseline_dtype=np.float64)
    gls = ctx.reward_batch(t, 0.5, "ls", torch.float64)
    np.testing.assert_allclose(gls["residual"].cpu().numpy()[:5], wls["residual"], rtol=1e-10, atol=1e-9)
    gls32 = ctx.reward_batch(t, 0.5, "ls", torch.float32)
    np.testing.assert_allclose(gls32["residual"].cpu().numpy()[:5], wls["residual"], rtol=1e-4)
    ctx.close()


def test_degenerate_shapes():
    """n = 1, an empty initial matrix (E = 0), empty rows of A, B = 1 and T = 1."""
    from gflownet_spai_b200.env import SpaiContext
    one = np.array([0], dtype=np.int64)
    ctx = SpaiContext(1, one, one, np.array([2.0]), one, one, np.array([3.0]))
    out = ctx.reward_batch(torch.tensor([[1], [0]], dtype=torch.int64).cuda(), 0.5, "copy", torch.float64)
    assert out["residual"].tolist() == [5.0, 1.0] and out["nnz_m"].tolist() == [1, 0]       # |2*3 - 1|, |0 - 1|
    ls = ctx.reward_batch(torch.tensor([[1], [0]], dtype=torch.int64).cuda(), 0.5, "ls", torch.float64)
    assert float(ls["residual"][0]) == pytest.approx(0.0, abs=1e-12) and float(ls["residual"][1]) == pytest.approx(1.0)
    assert ctx.info().orig_residual_f64 == pytest.approx(8.0)                                 # |3*3 - 1|
    ctx.close()
    # E = 0: nothing to remove, M = 0, ||0 - I||_F = sqrt(n)
    n = 7
    empty = np.zeros(0, dtype=np.int64)
    a = sp.identity(n, format="coo") * 2.0
    ctx = SpaiContext(n, empty, empty, np.zeros(0), a.row, a.col, a.data)
    for mode in ("copy", "ls"):
        out = ctx.reward_batch(torch.tensor([[0, -1], [5, 3]], dtype=torch.int64).cuda(), 0.25, mode, torch.float64)
        assert torch.allclose(out["residual"], torch.full((2,), float(np.sqrt(n)), dtype=torch.float64, device="cuda"))
        assert out["nnz_m"].tolist() == [0, 0]
    host = ctx.reward_batch(torch.tensor([[0, -1]], dtype=torch.int64), 0.25, "copy", torch.float32)
    assert float(host["residual"][0]) == pytest.approx(np.sqrt(n))
    ctx.close()
    # A with empty rows: candidates pointing at them gather nothing
    rng = np.random.default_rng(3)
    n = 30
    a = sp.random(n, n, density=0.1, random_state=1, format="lil")
    a[4, :] = 0
    a[17, :] = 0
    a = sp.csr_matrix(a)
    a.eliminate_zeros()
    r = np.repeat(np.arange(n), 4).astype(np.int64)
    c = rng.integers(0, n, r.size).astype(np.int64)
    c[:8] = [4, 17, 4, 17, 4, 4, 17, 17]
    key = np.unique(r * n + c)
    r, c = key // n, key % n
    v = rng.uniform(-1, 1, r.size)
    coo = a.tocoo()
    ctx = SpaiContext(n, r, c, v, coo.row, coo.col, coo.data)
    acts = synth.make_trajectories(r.size, 4, seed0=6)
    want = orc.reward_batch_copy(n, r, c, v, a, acts, 0.5, dtype=np.float64)
    got = ctx.reward_batch(torch.from_numpy(acts).cuda(), 0.5, "copy", torch.float64)
    np.testing.assert_allclose(got["residual"].cpu().numpy(), want["residual"], rtol=RTOL64, atol=ATOL64)
    wls = orc.reward_batch_ls(n, r, c, a, acts, 0.5, dtype=np.float64, baseline_dtype=np.float64)
    gls = ctx.reward_batch(torch.from_numpy(acts).cuda(), 0.5, "ls", torch.float64)
    np.testing.assert_allclose(gls["residual"].cpu().numpy(), wls["residual"], rtol=1e-9, atol=1e-9)
    ctx.close()
