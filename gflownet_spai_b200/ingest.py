"""Device-side ingest: COO -> CSR, sparse products, candidate supersets and their
initial values on the GPU, through the C ABI (`spai_ingest_*`, include/spai_b200.h).

Reference pointers: `market_matrix_to_sparse_tensor` (gflownet/utils.py:54-63) yields the COO this
module starts from; the drivers form the initial matrix as a product of sparse factors
(GFlowNet100.py:126-153, `L @ U`). `superset_pattern` / `neumann_values` are the device versions of
`synth.superset_pattern` / `synth.neumann_values` (bit-exact edge lists; values to 1e-12).
PyTorch is used for device memory only; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check

__all__ = ["CsrDev", "coo_to_csr", "spgemm", "drop_zeros", "superset_pattern", "neumann_values", "read_matrix_market"]


class CsrDev:
    """CSR on the device: ptr int32[n+1], col int32[nnz] (ascending per row), val float64[nnz]."""

    def __init__(self, n, ptr, col, val):
        self.n, self.ptr, self.col, self.val = int(n), ptr, col, val

    @property
    def nnz(self):
        return int(self.col.numel())

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.val.cpu().numpy(), self.col.cpu().numpy(), self.ptr.cpu().numpy()), shape=(self.n, self.n))

    @staticmethod
    def from_scipy(a, device=0):
        a = a.tocsr()
        a.sort_indices()
        dev = torch.device("cuda", device)
        return CsrDev(a.shape[0], torch.from_numpy(a.indptr.astype("int32")).to(dev),
                      torch.from_numpy(a.indices.astype("int32")).to(dev), torch.from_numpy(a.data.astype("float64")).to(dev))


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def coo_to_csr(n: int, row: torch.Tensor, col: torch.Tensor, val: torch.Tensor) -> CsrDev:
    """COO (CUDA int64 row/col, float64 val, any order, repeated coordinates summed) -> coalesced CSR."""
    lib = _lib.load()
    dev = row.device
    if not row.is_cuda:
        raise ValueError("coo_to_csr: inputs must be CUDA tensors (no CPU fallback)")
    row, col = row.to(torch.int64).contiguous(), col.to(torch.int64).contiguous()
    val = val.to(torch.float64).contiguous()
    nnz = row.numel()
    ptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    ocol = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    oval = torch.empty(max(nnz, 1), dtype=torch.float64, device=dev)
    total = C.c_int64(0)
    check(lib.spai_ingest_coo_to_csr_dev(dev.index, n, nnz, _p(row), _p(col), _p(val), _p(ptr), _p(ocol), _p(oval),
                                         C.byref(total), _stream(dev)), "spai_ingest_coo_to_csr_dev")
    return CsrDev(n, ptr, ocol[: total.value], oval[: total.value])


def drop_zeros(a: CsrDev) -> CsrDev:
    """Remove stored entries that are exactly 0 (what scipy's sparse product does)."""
    lib = _lib.load()
    dev = a.ptr.device
    optr = torch.empty(a.n + 1, dtype=torch.int32, device=dev)
    ocol = torch.empty(max(a.nnz, 1), dtype=torch.int32, device=dev)
    oval = torch.empty(max(a.nnz, 1), dtype=torch.float64, device=dev)
    total = C.c_int64(0)
    check(lib.spai_ingest_csr_drop_zeros_dev(dev.index, a.n, _p(a.ptr), _p(a.col), _p(a.val), _p(optr), _p(ocol), _p(oval),
                                             C.byref(total), _stream(dev)), "spai_ingest_csr_drop_zeros_dev")
    return CsrDev(a.n, optr, ocol[: total.value], oval[: total.value])


def spgemm(a: CsrDev, b: CsrDev, prune_zeros: bool = True) -> CsrDev:
    """C = A @ B (CSR, columns ascending, deterministic summation order). prune_zeros=True drops
    results that are exactly 0, as scipy's product (the reference's `L @ U`) does."""
    lib = _lib.load()
    dev = a.ptr.device
    cptr = torch.empty(a.n + 1, dtype=torch.int32, device=dev)
    total = C.c_int64(0)
    check(lib.spai_ingest_spgemm_count_dev(dev.index, a.n, _p(a.ptr), _p(a.col), _p(b.ptr), _p(b.col), _p(cptr),
                                           C.byref(total), _stream(dev)), "spai_ingest_spgemm_count_dev")
    ccol = torch.empty(max(total.value, 1), dtype=torch.int32, device=dev)
    cval = torch.empty(max(total.value, 1), dtype=torch.float64, device=dev)
    check(lib.spai_ingest_spgemm_fill_dev(dev.index, a.n, _p(a.ptr), _p(a.col), _p(a.val), _p(b.ptr), _p(b.col), _p(b.val),
                                          _p(cptr), _p(ccol), _p(cval), _stream(dev)), "spai_ingest_spgemm_fill_dev")
    out = CsrDev(a.n, cptr, ccol[: total.value], cval[: total.value])
    return drop_zeros(out) if prune_zeros else out


def superset_pattern(a: CsrDev, k: int, max_power: int = 4, order: str = "distance"):
    """(s_ptr int64[n+1], s_row int64[E], s_col int64[E]) of the candidate superset S, row-major with
    ascending columns — `synth.superset_pattern` (order="distance") or the cfg5 rule (order="band")."""
    lib = _lib.load()
    dev = a.ptr.device
    sptr = torch.empty(a.n + 1, dtype=torch.int64, device=dev)
    srow = torch.empty(a.n * k, dtype=torch.int64, device=dev)
    scol = torch.empty(a.n * k, dtype=torch.int64, device=dev)
    total = C.c_int64(0)
    check(lib.spai_ingest_superset_dev(dev.index, a.n, _p(a.ptr), _p(a.col), int(k), int(max_power),
                                       {"distance": 0, "band": 1}[order], _p(sptr), _p(srow), _p(scol), C.byref(total),
                                       _stream(dev)), "spai_ingest_superset_dev")
    return sptr, srow[: total.value], scol[: total.value]


def neumann_values(a: CsrDev, s_ptr: torch.Tensor, s_col: torch.Tensor, terms: int = 3) -> torch.Tensor:
    """float64[E] initial values on S (device version of `synth.neumann_values`)."""
    lib = _lib.load()
    dev = a.ptr.device
    out = torch.empty(max(s_col.numel(), 1), dtype=torch.float64, device=dev)
    omega = C.c_double(0.0)
    check(lib.spai_ingest_neumann_dev(dev.index, a.n, _p(a.ptr), _p(a.col), _p(a.val), _p(s_ptr), _p(s_col), int(terms),
                                      C.byref(omega), _p(out), _stream(dev)), "spai_ingest_neumann_dev")
    return out[: s_col.numel()]


def read_matrix_market(path: str, device: int = 0) -> CsrDev:
    """gflownet/utils.py:54-63: a MatrixMarket coordinate file -> (here) coalesced CSR on the device.
    The text is tokenised on the host (numpy); expansion of `symmetric` / `skew-symmetric` / `pattern`
    files, sorting and coalescing run on the GPU."""
    import numpy as np
    with open(path, "rb") as f:
        head = f.readline().decode().lower().split()
        if len(head) < 5 or head[0] != "%%matrixmarket" or head[1] != "matrix" or head[2] != "coordinate":
            raise ValueError("read_matrix_market: only `matrix coordinate` files are supported")
        field, symm = head[3], head[4]
        line = f.readline()
        while line.startswith(b"%") or not line.strip():
            line = f.readline()
        nrow, ncol, nent = (int(x) for x in line.split()[:3])
        if nrow != ncol:
            raise ValueError("read_matrix_market: the matrix must be square")
        body = np.loadtxt(f, dtype=np.float64, ndmin=2) if nent else np.zeros((0, 3))
    if body.shape[0] != nent:
        raise ValueError("read_matrix_market: entry count does not match the size line")
    dev = torch.device("cuda", device)
    r = torch.from_numpy(body[:, 0].astype(np.int64) - 1).to(dev)
    c = torch.from_numpy(body[:, 1].astype(np.int64) - 1).to(dev)
    v = torch.ones(nent, dtype=torch.float64, device=dev) if field == "pattern" else torch.from_numpy(body[:, 2].copy()).to(dev)
    if symm in ("symmetric", "skew-symmetric", "hermitian"):
        off = r != c
        sign = -1.0 if symm == "skew-symmetric" else 1.0
        r, c, v = torch.cat([r, c[off]]), torch.cat([c, r[off]]), torch.cat([v, sign * v[off]])
    return coo_to_csr(nrow, r, c, v)
