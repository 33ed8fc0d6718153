"""Deterministic synthetic inputs for the SPAI reward path.

Matrices (scipy CSR, fp64 values; callers cast), candidate superset patterns
``S`` (<= k entries per row, ordered by (graph distance, column id)), initial
values on ``S`` and trajectory batches in the reference's ``actions[B, T]``
layout (-1 padded, last valid id = terminal id ``E``; gflownet/gflownet.py:181,
gflownet/log.py:84-87 in the reference).

Host-side numpy/scipy only: this module prepares *inputs*; it never computes a
reward.
"""
from __future__ import annotations

import dataclasses

import numpy as np
import scipy.sparse as sp

__all__ = [
    "poisson2d", "poisson3d", "convdiff2d", "banded_powerlaw",
    "superset_pattern", "neumann_values", "make_trajectories", "Problem",
    "make_problem", "CONFIGS",
]


# --------------------------------------------------------------------------
# matrices
# --------------------------------------------------------------------------
def _tridiag(m: int, lo: float, di: float, up: float) -> sp.csr_matrix:
    return sp.diags([np.full(m - 1, lo), np.full(m, di), np.full(m - 1, up)],
                    [-1, 0, 1], format="csr")


def poisson2d(nx: int, ny: int | None = None) -> sp.csr_matrix:
    """2-D 5-point Poisson, kron(I,T)+kron(T,I), T=tridiag(-1,2,-1)."""
    ny = nx if ny is None else ny
    a = sp.kron(sp.identity(ny), _tridiag(nx, -1.0, 2.0, -1.0)) + \
        sp.kron(_tridiag(ny, -1.0, 2.0, -1.0), sp.identity(nx))
    a = sp.csr_matrix(a)
    a.eliminate_zeros()
    a.sort_indices()
    return a


def poisson3d(m: int) -> sp.csr_matrix:
    """3-D 7-point Poisson on an m^3 grid."""
    t = _tridiag(m, -1.0, 2.0, -1.0)
    i = sp.identity(m)
    a = sp.kron(sp.kron(i, i), t) + sp.kron(sp.kron(i, t), i) + sp.kron(sp.kron(t, i), i)
    a = sp.csr_matrix(a)
    a.eliminate_zeros()
    a.sort_indices()
    return a


def convdiff2d(m: int, beta: float = 10.0) -> sp.csr_matrix:
    """Non-symmetric 2-D convection-diffusion, central differences.

    -Laplace(u) + beta * du/dx on an m x m grid with h = 1/(m+1); scaled by h^2
    so the diffusion stencil is (-1, 4, -1) and the convection adds
    -+ beta*h/2 on the x neighbours (non-integer entries => fp32 rounding is
    visible, unlike Poisson).
    """
    h = 1.0 / (m + 1)
    c = 0.5 * beta * h
    tx = _tridiag(m, -1.0 - c, 2.0, -1.0 + c)
    ty = _tridiag(m, -1.0, 2.0, -1.0)
    a = sp.kron(sp.identity(m), tx) + sp.kron(ty, sp.identity(m))
    a = sp.csr_matrix(a)
    a.eliminate_zeros()
    a.sort_indices()
    return a


def banded_powerlaw(n: int, seed: int = 12345, half_bw: int = 4,
                    zipf_a: float = 2.1, cap: int = 64) -> sp.csr_matrix:
    """Band (2*half_bw+1 diagonals, U(-1,1)) + power-law extra entries.

    Row degree of the extras ~ Zipf(zipf_a) capped at ``cap``; columns uniform;
    diagonal += row abs sum (diagonal dominance).
    """
    rng = np.random.default_rng(seed)
    rows, cols, vals = [], [], []
    for d in range(-half_bw, half_bw + 1):
        lo, hi = max(0, -d), min(n, n - d)
        r = np.arange(lo, hi, dtype=np.int64)
        rows.append(r)
        cols.append(r + d)
        vals.append(rng.uniform(-1.0, 1.0, size=r.size))
    deg = np.minimum(rng.zipf(zipf_a, size=n), cap).astype(np.int64)
    deg[deg == 1] = 0  # Zipf mode (1) means "no extras"
    tot = int(deg.sum())
    r = np.repeat(np.arange(n, dtype=np.int64), deg)
    rows.append(r)
    cols.append(rng.integers(0, n, size=tot, dtype=np.int64))
    vals.append(rng.uniform(-1.0, 1.0, size=tot))
    a = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(n, n)).tocsr()
    a.sum_duplicates()
    rowabs = np.asarray(abs(a).sum(axis=1)).ravel()
    a = a + sp.diags(rowabs, 0, format="csr")
    a = sp.csr_matrix(a)
    a.sort_indices()
    return a


# --------------------------------------------------------------------------
# candidate superset S and initial values
# --------------------------------------------------------------------------
def superset_pattern(a: sp.csr_matrix, k: int, max_power: int = 4):
    """Rows/cols (row-major, per row sorted by column) of the superset S.

    S(i) = the first k entries of pattern(A) U pattern(A^2) U ... ordered by
    (graph distance from i, column id). Powers are added until every row has k
    candidates or ``max_power`` is reached. Returns (row int64[E], col int64[E]).
    """
    n = a.shape[0]
    pat = sp.csr_matrix((np.ones(a.nnz, dtype=np.float32), a.indices, a.indptr), shape=a.shape)
    eye = sp.identity(n, dtype=np.float32, format="csr")
    reach = sp.csr_matrix(eye)           # distance <= 0
    rows = [np.arange(n, dtype=np.int64)]
    cols = [np.arange(n, dtype=np.int64)]
    dist = [np.zeros(n, dtype=np.int64)]
    for d in range(1, max_power + 1):
        nxt = sp.csr_matrix(reach @ pat + reach)
        nxt.data[:] = 1.0
        new = sp.csr_matrix(nxt - reach)
        new.eliminate_zeros()
        new = new.tocoo()
        rows.append(new.row.astype(np.int64))
        cols.append(new.col.astype(np.int64))
        dist.append(np.full(new.nnz, d, dtype=np.int64))
        reach = nxt
        if np.diff(reach.indptr).min() >= k:
            break
    r = np.concatenate(rows)
    c = np.concatenate(cols)
    dd = np.concatenate(dist)
    order = np.lexsort((c, dd, r))
    r, c = r[order], c[order]
    start = np.searchsorted(r, np.arange(n))
    rank = np.arange(r.size) - start[r]
    keep = rank < k
    r, c = r[keep], c[keep]
    order = np.lexsort((c, r))           # row-major, columns ascending
    return r[order], c[order]


def neumann_values(a: sp.csr_matrix, s_row: np.ndarray, s_col: np.ndarray,
                   terms: int = 3) -> np.ndarray:
    """Initial values on S: omega * sum_{j<terms} (I - omega*A)^j restricted to S.

    A deterministic, non-integer stand-in for "some approximate inverse on S"
    (the reference's drivers use spilu L@U, GFlowNet100.py:126-153). fp64.
    """
    n = a.shape[0]
    omega = 1.0 / float(abs(a).sum(axis=1).max())
    g = sp.identity(n, format="csr") - omega * a
    acc = sp.identity(n, format="csr")
    p = sp.identity(n, format="csr")
    for _ in range(1, terms):
        p = sp.csr_matrix(p @ g)
        acc = acc + p
    acc = sp.csr_matrix(omega * acc)
    vals = np.asarray(acc[s_row, s_col]).ravel().astype(np.float64)
    return vals


# --------------------------------------------------------------------------
# trajectories
# --------------------------------------------------------------------------
def make_trajectories(num_edges: int, batch: int, seed0: int = 1000,
                      max_frac: float = 0.5, first: int = 0) -> np.ndarray:
    """actions int64[B, T]: trajectory b deletes floor(u_b * E * max_frac)
    distinct edges, then the terminal id E, -1 padded to max T_b + 1.

    Seeded per *global* trajectory index (``first`` + b) so a shard of a batch
    equals the same rows of the full batch.
    """
    e = int(num_edges)
    lens = np.empty(batch, dtype=np.int64)
    picks = []
    for b in range(batch):
        rng = np.random.default_rng(seed0 + first + b)
        t = int(np.floor(rng.random() * e * max_frac))
        # partial Fisher-Yates via choice without replacement
        picks.append(rng.choice(e, size=t, replace=False).astype(np.int64) if t else
                     np.empty(0, dtype=np.int64))
        lens[b] = t
    tmax = int(lens.max()) + 1 if batch else 1
    out = np.full((batch, tmax), -1, dtype=np.int64)
    for b in range(batch):
        t = int(lens[b])
        out[b, :t] = picks[b]
        out[b, t] = e
    return out


# --------------------------------------------------------------------------
# named configs (BASELINE.json "configs")
# --------------------------------------------------------------------------
@dataclasses.dataclass
class Problem:
    name: str
    n: int
    a: sp.csr_matrix            # original matrix A (fp64)
    edge_row: np.ndarray        # int64[E]  initial-matrix entries in action order
    edge_col: np.ndarray        # int64[E]
    edge_val: np.ndarray        # float64[E]
    k: int                      # max candidates per row
    batch: int                  # default trajectory batch

    @property
    def num_edges(self) -> int:
        return int(self.edge_row.size)


CONFIGS = {
    # name: (builder, k, default batch)
    "cfg1": ("poisson2d-10x10 initial=A", None, 32),
    "cfg2": ("poisson2d-256x256 S<=8/row", 8, 4096),
    "cfg3": ("poisson3d-64^3 S<=16/row", 16, 4096),
    "cfg4": ("convdiff2d-512x512 S<=32/row", 32, 1024),
    "cfg5": ("banded+powerlaw n=1M S=pattern(A)<=32/row", 32, 16384),
}


def _from_superset(name, a, k, batch, max_power):
    r, c = superset_pattern(a, k, max_power=max_power)
    v = neumann_values(a, r, c, terms=min(max_power, 3) + 1)
    return Problem(name, a.shape[0], a, r, c, v, k, batch)


def make_problem(name: str, scale: float = 1.0) -> Problem:
    """Build a named config; ``scale`` < 1 shrinks the grid for tests."""
    if name == "cfg1":
        a = poisson2d(10)
        coo = a.tocoo()
        return Problem(name, 100, a, coo.row.astype(np.int64), coo.col.astype(np.int64),
                       coo.data.astype(np.float64), 5, 32)
    if name == "cfg2":
        m = max(4, int(round(256 * scale)))
        return _from_superset(name, poisson2d(m), 8, 4096, 2)
    if name == "cfg3":
        m = max(3, int(round(64 * scale)))
        return _from_superset(name, poisson3d(m), 16, 4096, 2)
    if name == "cfg4":
        m = max(6, int(round(512 * scale)))
        return _from_superset(name, convdiff2d(m), 32, 1024, 4)
    if name == "cfg5":
        n = max(64, int(round(1_000_000 * scale)))
        a = banded_powerlaw(n)
        # S = pattern(A) capped at 32/row (first 32 by column id distance order
        # is not defined for a random graph: keep the 32 smallest |col - row|).
        coo = a.tocoo()
        r = coo.row.astype(np.int64)
        c = coo.col.astype(np.int64)
        order = np.lexsort((c, np.abs(c - r), r))
        r, c = r[order], c[order]
        start = np.searchsorted(r, np.arange(n))
        keep = (np.arange(r.size) - start[r]) < 32
        r, c = r[keep], c[keep]
        order = np.lexsort((c, r))
        r, c = r[order], c[order]
        v = neumann_values(a, r, c, terms=2)
        return Problem(name, n, a, r, c, v, 32, 16384)
    raise KeyError(name)
