"""Device-side ingest against scipy / synth on the host: COO -> CSR (coalesce),
SpGEMM (the drivers' `L @ U`, GFlowNet100.py:139-141), candidate supersets (bit-exact edge lists vs
`synth.superset_pattern`) and their initial values (<= 1e-12 vs `synth.neumann_values`)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from gflownet_spai_b200 import synth

pytestmark = pytest.mark.gpu


def _dev(a):
    from gflownet_spai_b200 import ingest
    return ingest.CsrDev.from_scipy(a, 0)


def test_coo_to_csr_coalesces_like_scipy():
    from gflownet_spai_b200 import ingest
    rng = np.random.default_rng(0)
    n, nnz = 3000, 40000
    r = rng.integers(0, n, nnz)
    c = rng.integers(0, n, nnz)
    r[:500], c[:500] = r[500:1000], c[500:1000]                    # repeated coordinates
    v = rng.uniform(-1, 1, nnz)
    got = ingest.coo_to_csr(n, torch.from_numpy(r).cuda(), torch.from_numpy(c).cuda(), torch.from_numpy(v).cuda())
    want = sp.coo_matrix((v, (r, c)), shape=(n, n)).tocsr()
    want.sum_duplicates()
    want.sort_indices()
    assert np.array_equal(got.ptr.cpu().numpy(), want.indptr)
    assert np.array_equal(got.col.cpu().numpy(), want.indices)
    np.testing.assert_allclose(got.val.cpu().numpy(), want.data, rtol=1e-14, atol=1e-15)
    # one long row (> 512 entries): the serial path
    r2 = np.concatenate([np.zeros(700, dtype=np.int64), r[:100]])
    c2 = np.concatenate([rng.permutation(n)[:700], c[:100]])
    v2 = rng.uniform(-1, 1, r2.size)
    got = ingest.coo_to_csr(n, torch.from_numpy(r2).cuda(), torch.from_numpy(c2).cuda(), torch.from_numpy(v2).cuda())
    want = sp.coo_matrix((v2, (r2, c2)), shape=(n, n)).tocsr()
    want.sum_duplicates()
    want.sort_indices()
    assert np.array_equal(got.col.cpu().numpy(), want.indices)
    np.testing.assert_allclose(got.val.cpu().numpy(), want.data, rtol=1e-14, atol=1e-15)
    with pytest.raises(ValueError):
        ingest.coo_to_csr(10, torch.tensor([0, 11]).cuda(), torch.tensor([0, 1]).cuda(), torch.ones(2, dtype=torch.float64).cuda())


def test_spgemm_matches_scipy_lu_product():
    """The drivers' initial matrix: LU = L @ U of an incomplete factorisation (GFlowNet100.py:126-141)."""
    from gflownet_spai_b200 import ingest
    import scipy.sparse.linalg as spla
    a = sp.csc_matrix(synth.convdiff2d(24))
    ilu = spla.spilu(a)
    lo = sp.tril(ilu.L, format="csr")
    up = sp.triu(ilu.U, format="csr")
    lo.sort_indices()
    up.sort_indices()
    want = sp.csr_matrix(lo @ up)                                      # scipy prunes results that are exactly 0
    want.sort_indices()
    got = ingest.spgemm(_dev(lo), _dev(up))
    g = got.to_scipy()
    assert np.array_equal(g.indptr, want.indptr) and np.array_equal(g.indices, want.indices)
    assert np.array_equal(g.data, want.data)                           # same products, same order, no FMA: bit-exact


@pytest.mark.parametrize("cfg,scale", [("cfg2", 0.25), ("cfg3", 0.4), ("cfg4", 0.12), ("cfg5", 0.02)])
def test_superset_and_initial_values_match_synth(cfg, scale):
    from gflownet_spai_b200 import ingest
    p = synth.make_problem(cfg, scale)
    a = _dev(p.a)
    k, power, order, terms = {"cfg2": (8, 2, "distance", 3), "cfg3": (16, 2, "distance", 3),
                              "cfg4": (32, 4, "distance", 4), "cfg5": (32, 1, "band", 2)}[cfg]
    sptr, srow, scol = ingest.superset_pattern(a, k, power, order)
    assert np.array_equal(srow.cpu().numpy(), p.edge_row), "edge rows differ"
    assert np.array_equal(scol.cpu().numpy(), p.edge_col), "edge columns differ"
    assert int(sptr[-1]) == p.num_edges
    vals = ingest.neumann_values(a, sptr, scol, terms)
    np.testing.assert_allclose(vals.cpu().numpy(), p.edge_val, rtol=1e-12, atol=1e-15)


def test_full_size_cfg5_ingest_and_setup_time():
    """n = 1e6: superset + values on the device equal the host construction; reports the times."""
    import time
    from gflownet_spai_b200 import ingest
    t0 = time.perf_counter()
    p = synth.make_problem("cfg5")
    t_host = time.perf_counter() - t0
    a = _dev(p.a)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sptr, srow, scol = ingest.superset_pattern(a, 32, 1, "band")
    vals = ingest.neumann_values(a, sptr, scol, 2)
    torch.cuda.synchronize()
    t_dev = time.perf_counter() - t0
    assert np.array_equal(scol.cpu().numpy(), p.edge_col) and np.array_equal(srow.cpu().numpy(), p.edge_row)
    np.testing.assert_allclose(vals.cpu().numpy(), p.edge_val, rtol=1e-12, atol=1e-15)
    print(f"cfg5 superset + values: device {t_dev * 1e3:.1f} ms, host synth.make_problem (matrix + superset + values) {t_host:.1f} s")


def test_read_matrix_market_symmetric(tmp_path):
    from gflownet_spai_b200 import ingest
    import scipy.io
    a = synth.poisson2d(9)
    path = os.path.join(tmp_path, "p.mtx")
    scipy.io.mmwrite(path, sp.coo_matrix(a), symmetry="symmetric")
    got = ingest.read_matrix_market(path).to_scipy()
    want = sp.csr_matrix(scipy.io.mmread(path))
    want.sort_indices()
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    assert np.array_equal(got.data, want.data)
