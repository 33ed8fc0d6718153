"""The CPU oracle against the reference's own outputs (tests/golden, generated
by tests/golden/make_golden.py from the live reference) and, when the checkout
is present, against the live reference."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import ref_shim
from oracle import spai_oracle as orc
from gflownet_spai_b200 import synth


def _a_csr(g, dtype=np.float32):
    n = int(g["n"])
    a = sp.coo_matrix((g["a_val"].astype(dtype), (g["a_row"], g["a_col"])), shape=(n, n)).tocsr()
    a.sum_duplicates()
    a.sort_indices()
    return a


def test_oracle_copy_rewards_match_reference_golden(golden):
    g = golden
    n = int(g["n"])
    out = orc.reward_batch_copy(n, g["edge_row"], g["edge_col"], g["edge_val"], _a_csr(g),
                                g["actions"], float(g["alpha"]), dtype=np.float32,
                                a_stored_nnz=g["a_val"].size)
    assert out["orig_flops"] == int(g["orig_flops"])
    assert out["orig_residual"] == pytest.approx(float(g["orig_residual"]), rel=1e-12)
    np.testing.assert_allclose(out["reward"], g["reward"], rtol=1e-10, atol=1e-9)


def test_oracle_pattern_indices_bit_exact(golden):
    g = golden
    n = int(g["n"])
    for b in range(g["actions"].shape[0]):
        kept = orc.kept_edge_mask(g["edge_row"].size, g["actions"][b])
        m = orc.build_pattern_matrix(n, g["edge_row"], g["edge_col"], g["edge_val"], kept).tocoo()
        lo, hi = int(g["m_ptr"][b]), int(g["m_ptr"][b + 1])
        assert np.array_equal(m.row, g["m_row"][lo:hi])
        assert np.array_equal(m.col, g["m_col"][lo:hi])
        np.testing.assert_allclose(m.data, g["m_val"][lo:hi], rtol=1e-6, atol=1e-7)


def test_known_answers_from_survey():
    g = dict(np.load(__import__("os").path.join(__import__("conftest").GOLDEN_DIR, "tiny3.npz")))
    assert float(g["orig_residual"]) == pytest.approx(19.28730152198591, rel=1e-14)
    assert int(g["orig_flops"]) == 36
    assert g["reward"][0] == 0.0
    assert g["reward"][1] == pytest.approx(955.0987, rel=1e-6)
    assert g["reward"][2] == pytest.approx(105.3222, rel=1e-6)
    p = dict(np.load(__import__("os").path.join(__import__("conftest").GOLDEN_DIR, "poisson10.npz")))
    assert float(p["orig_residual"]) == pytest.approx(243.48305895893455, rel=1e-14)
    assert int(p["orig_flops"]) == 92000


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference checkout absent")
def test_oracle_matches_live_reference_random():
    rng = np.random.default_rng(7)
    n = 24
    a = sp.random(n, n, density=0.15, random_state=3, format="csr") + sp.identity(n) * 2
    a = sp.csr_matrix(a)
    r, c = synth.superset_pattern(a, 6, max_power=2)
    v = rng.uniform(-1, 1, r.size)
    acts = synth.make_trajectories(r.size, 5, seed0=3)
    coo = a.tocoo()
    ref = ref_shim.reference_update(n, r, c, v, coo.row, coo.col, coo.data, acts, 0.4)
    out = orc.reward_batch_copy(n, r, c, v.astype(np.float32), sp.csr_matrix(a, dtype=np.float32),
                                acts, 0.4, dtype=np.float32)
    np.testing.assert_allclose(out["reward"], ref["reward"], rtol=1e-10, atol=1e-9)


def test_ls_never_worse_than_copy_and_solvers_agree():
    a = synth.convdiff2d(8)
    n = a.shape[0]
    r, c = synth.superset_pattern(a, 8, max_power=2)
    v = synth.neumann_values(a, r, c)
    acts = synth.make_trajectories(r.size, 3, seed0=11)
    cp = orc.reward_batch_copy(n, r, c, v, a, acts, 0.5, dtype=np.float64)
    ls = orc.reward_batch_ls(n, r, c, a, acts, 0.5, dtype=np.float64, baseline_dtype=np.float64)
    assert np.all(ls["residual"] <= cp["residual"] + 1e-12)
    assert np.array_equal(ls["nnz_m"], cp["nnz_m"])
    # independent check of one row via normal equations
    kept = orc.kept_edge_mask(r.size, acts[0])
    pat = orc.build_pattern_matrix(n, r, c, np.ones(r.size), kept, dtype=np.float64)
    i = n // 2
    j, iset = orc.row_index_sets(pat, a, i)
    hat = np.asarray(a[j, :][:, iset].todense()).T
    rhs = (iset == i).astype(float)
    x = np.linalg.solve(hat.T @ hat, hat.T @ rhs)
    r2 = np.sum((hat @ x - rhs) ** 2) + (0.0 if np.any(iset == i) else 1.0)
    assert orc.ls_row_residual2(a, i, j) == pytest.approx(r2, rel=1e-9)


def test_sample_step_known_answers():
    logits = np.log(np.array([0.1, 0.2, 0.3, 0.4], dtype=np.float32))
    act, prob, done = orc.sample_step(logits, [[], [3], [0, 1], []], [0.05, 0.99, 0.5, 0.0],
                                      [False, False, False, True])
    assert act.tolist() == [0, 2, 3, -1]
    assert done.tolist() == [False, False, True, True]
    assert prob[3] == 1.0
    assert prob[1] == pytest.approx(0.5, rel=1e-5)          # 0.3 / (0.1+0.2+0.3)
    assert prob[2] == pytest.approx(4.0 / 7.0, rel=1e-5)


def test_philox_known_answers_and_race_oracle():
    """Philox4x32-10 against the Random123 known-answer vectors (kat_vectors of the reference
    implementation), then the race oracle's invariants."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        got = orc.philox4x32_10(np.array(c, dtype=np.uint64), np.array(k, dtype=np.uint64))
        assert tuple(int(v) for v in got) == want
    lg = np.random.default_rng(0).normal(size=301)
    lq = orc.race_keys(lg, seed=9, sample=4)
    tr = orc.race_trajectory(lq)
    assert tr[-1] == 300 and len(set(tr.tolist())) == tr.size and lq[300] == 0.0
    assert np.all(np.diff(lq[tr[:-1]]) >= 0) and np.all(lq[tr[:-1]] < 0)
    assert not np.array_equal(lq, orc.race_keys(lg, seed=9, sample=5))
    # first arrival ~ softmax (exponential race), 20000 samples, 5 categories
    lg5 = np.array([0.0, 1.0, -1.0, 0.5, 0.2])
    p = np.exp(lg5) / np.exp(lg5).sum()
    first = np.array([int(np.argmin(orc.race_keys(lg5, 1, s))) for s in range(20000)])
    cnt = np.bincount(first, minlength=5)
    assert float(((cnt - 20000 * p) ** 2 / (20000 * p)).sum()) < 23.5        # chi2(4 dof) 99.99th percentile
