"""CPU oracle for the SPAI reward path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; the product path
(``gflownet_spai_b200``) never does and fails loudly without its CUDA library.

A numpy/scipy restatement of the reference's algorithm (tonylizza/gflownet-spai).
Citations are ``file:line`` relative to the reference checkout.

Parity pinning
--------------
* ``copy`` mode (what the reference computes): PINNED. ``tests/golden/*.npz`` hold
  inputs and outputs produced by importing the reference itself
  (``tests/golden/make_golden.py`` via ``oracle/ref_shim.py``) and
  ``tests/test_oracle.py`` checks this restatement against them (and against the
  live reference when ``/root/reference`` is present).
* ``ls`` mode (true SPAI re-solve named by BASELINE.json's north star): the
  reference has no such code and no tests, goldens or fixtures for it —
  **parity unpinned**. The restatement below is the published SPAI row
  least-squares (Grote & Huckle 1997) solved with LAPACK (``numpy.linalg``);
  it is cross-checked by invariants only (``ls`` residual <= ``copy`` residual,
  two independent solvers agree).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

__all__ = [
    "strip_actions", "kept_edge_mask", "build_pattern_matrix", "residual_copy",
    "matrix_flops", "evaluate_preconditioner", "reward_from_residual",
    "baseline_constants", "reward_batch_copy", "row_index_sets",
    "ls_row_residual2", "residual_ls", "reward_batch_ls", "masked_softmax_probs",
    "sample_step", "philox4x32_10", "race_keys", "race_trajectory",
]


# --------------------------------------------------------------------------
# actions -> kept-edge mask -> M
# --------------------------------------------------------------------------
def strip_actions(actions_row) -> list[int]:
    """preconditioner.py:38-39 — drop the -1 padding, keep everything else
    (the terminal id E stays in the list)."""
    return [int(x) for x in np.asarray(actions_row).ravel() if int(x) != -1]


def kept_edge_mask(num_edges: int, actions_row) -> np.ndarray:
    """bool[E]: utils.py:315-323 — ``i not in set(actions)`` for i in range(E).

    Duplicates collapse; ids >= E (the terminal id) and negative ids other than
    the stripped -1 never match an edge index and are ignored.
    """
    kept = np.ones(num_edges, dtype=bool)
    acts = np.asarray(strip_actions(actions_row), dtype=np.int64)
    acts = acts[(acts >= 0) & (acts < num_edges)]
    kept[acts] = False
    return kept


def build_pattern_matrix(n: int, edge_row, edge_col, edge_val, kept, dtype=np.float32) -> sp.csr_matrix:
    """M as the reference builds it: kept edges, flattened key row*n+col,
    coalesced (duplicates summed, keys sorted), un-flattened, coalesced again
    (utils.py:331-353, utils.py:114-124). Values are cast to fp32 by the
    reference (utils.py:350-352); ``dtype=np.float64`` is the lifted variant.

    Explicit zeros are kept as stored entries (they count in nnz(M))."""
    er = np.asarray(edge_row, dtype=np.int64)[kept]
    ec = np.asarray(edge_col, dtype=np.int64)[kept]
    ev = np.asarray(edge_val)[kept].astype(dtype)
    key = er * n + ec
    order = np.argsort(key, kind="stable")
    key, ev = key[order], ev[order]
    if key.size:
        first = np.ones(key.size, dtype=bool)
        first[1:] = key[1:] != key[:-1]
        grp = np.cumsum(first) - 1
        vals = np.zeros(int(grp[-1]) + 1, dtype=dtype)
        np.add.at(vals, grp, ev)           # sequential, stable order
        ukey = key[first]
    else:
        vals = np.zeros(0, dtype=dtype)
        ukey = key
    rows = ukey // n
    cols = ukey % n
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(indptr, rows + 1, 1)
    indptr = np.cumsum(indptr)
    m = sp.csr_matrix((vals, cols.astype(np.int64), indptr), shape=(n, n))
    return m


# --------------------------------------------------------------------------
# residual and reward (copy mode)
# --------------------------------------------------------------------------
def residual_copy(m: sp.csr_matrix, a: sp.csr_matrix, dtype=np.float32) -> float:
    """|| M @ A - I ||_F: product in ``dtype`` (reference: fp32 SpGEMM,
    preconditioner.py:88), subtraction of the fp64 identity and the norm in fp64
    (preconditioner.py:82-85, :90)."""
    n = a.shape[0]
    mm = sp.csr_matrix(m, dtype=dtype)
    aa = sp.csr_matrix(a, dtype=dtype)
    prod = sp.csr_matrix(mm @ aa)
    diff = prod.astype(np.float64) - sp.identity(n, dtype=np.float64, format="csr")
    diff = sp.csr_matrix(diff)
    return float(np.sqrt(np.sum(diff.data.astype(np.float64) ** 2)))


def matrix_flops(nnz: int, n: int) -> int:
    """preconditioner.py:68-72 sparse branch: 2 * stored entries * n."""
    return 2 * int(nnz) * int(n)


def evaluate_preconditioner(residual, flops, orig_residual, orig_flops, alpha) -> float:
    """preconditioner.py:154-163 (inf ratios when a baseline is zero)."""
    rr = residual / orig_residual if orig_residual != 0 else float("inf")
    cr = flops / orig_flops if orig_flops != 0 else float("inf")
    return alpha * (1 - rr) + (1 - alpha) * (1 - cr)


def reward_from_residual(residual, nnz_m, n, orig_residual, orig_flops, alpha) -> float:
    """preconditioner.py:59-64: metric -> float64 * 1000."""
    metric = evaluate_preconditioner(float(residual), matrix_flops(nnz_m, n),
                                     float(orig_residual), int(orig_flops), float(alpha))
    return float(np.float64(metric) * 1000.0)


def baseline_constants(a: sp.csr_matrix, a_stored_nnz: int | None = None, dtype=np.float32):
    """preconditioner.py:28-29: res0 = ||A0 @ A0 - I||_F, flops0 = 2*nnz(A0)*n
    where nnz counts the original's stored values (uncoalesced count)."""
    n = a.shape[0]
    res0 = residual_copy(a, a, dtype=dtype)
    nnz = a.nnz if a_stored_nnz is None else a_stored_nnz
    return res0, matrix_flops(nnz, n)


def reward_batch_copy(n, edge_row, edge_col, edge_val, a: sp.csr_matrix, actions, alpha,
                      dtype=np.float32, a_stored_nnz=None, baseline_dtype=None):
    """preconditioner.py:32-52 over a batch. Returns dict of arrays
    (reward f64[B], residual f64[B], nnz_m i64[B], kept bool[B,E])."""
    actions = np.asarray(actions)
    if actions.ndim == 1:
        actions = actions[None, :]
    b = actions.shape[0]
    e = len(edge_row)
    res0, flops0 = baseline_constants(a, a_stored_nnz, dtype=baseline_dtype or dtype)
    out_r = np.zeros(b)
    out_res = np.zeros(b)
    out_nnz = np.zeros(b, dtype=np.int64)
    kept_all = np.zeros((b, e), dtype=bool)
    for i in range(b):
        kept = kept_edge_mask(e, actions[i])
        m = build_pattern_matrix(n, edge_row, edge_col, edge_val, kept, dtype=dtype)
        res = residual_copy(m, a, dtype=dtype)
        out_res[i] = res
        out_nnz[i] = m.nnz
        out_r[i] = reward_from_residual(res, m.nnz, n, res0, flops0, alpha)
        kept_all[i] = kept
    return {"reward": out_r, "residual": out_res, "nnz_m": out_nnz, "kept": kept_all,
            "orig_residual": res0, "orig_flops": flops0}


# --------------------------------------------------------------------------
# ls mode (north-star SPAI; no reference code — parity unpinned)
# --------------------------------------------------------------------------
def row_index_sets(m_pattern: sp.csr_matrix, a: sp.csr_matrix, i: int):
    """J = kept columns of row i of M (sorted); I = union of cols(A[c,:]), c in J
    (sorted). Row-oriented form of SPAI (SURVEY.md §7 'Left vs right')."""
    j = m_pattern.indices[m_pattern.indptr[i]:m_pattern.indptr[i + 1]]
    j = np.unique(j)
    if j.size == 0:
        return j, np.zeros(0, dtype=np.int64)
    cols = [a.indices[a.indptr[c]:a.indptr[c + 1]] for c in j]
    return j, np.unique(np.concatenate(cols))


def ls_row_residual2(a: sp.csr_matrix, i: int, j_cols: np.ndarray, dtype=np.float64,
                     return_m: bool = False):
    """min_m || A(J, I)^T m - e_i(I) ||^2 (+1 when i is not in I, the uncovered
    diagonal of -I). Solved by LAPACK least squares."""
    j_cols = np.asarray(j_cols, dtype=np.int64)
    if j_cols.size == 0:
        return (1.0, np.zeros(0, dtype=dtype)) if return_m else 1.0
    sub = a[j_cols, :]
    iset = np.unique(sub.indices)
    hat = np.asarray(sub[:, iset].todense(), dtype=dtype).T        # |I| x |J|
    rhs = (iset == i).astype(dtype)
    sol, *_ = np.linalg.lstsq(hat, rhs, rcond=None)
    r = hat @ sol - rhs
    r2 = float(np.sum(r.astype(np.float64) ** 2))
    if not np.any(iset == i):
        r2 += 1.0
    return (r2, sol) if return_m else r2


def residual_ls(n, m_pattern: sp.csr_matrix, a: sp.csr_matrix, dtype=np.float64) -> float:
    """|| M_ls @ A - I ||_F with every row of M re-solved on its pattern."""
    aa = sp.csr_matrix(a, dtype=dtype)
    tot = 0.0
    for i in range(n):
        j = np.unique(m_pattern.indices[m_pattern.indptr[i]:m_pattern.indptr[i + 1]])
        tot += ls_row_residual2(aa, i, j, dtype=dtype)
    return float(np.sqrt(tot))


def reward_batch_ls(n, edge_row, edge_col, a: sp.csr_matrix, actions, alpha, dtype=np.float64,
                    a_stored_nnz=None, baseline_dtype=np.float32):
    """ls-mode batch: same masks, same nnz(M), same mix formula
    (preconditioner.py:154-163, :64); only the residual differs. The baseline
    constants stay the reference's (preconditioner.py:28-29)."""
    actions = np.asarray(actions)
    if actions.ndim == 1:
        actions = actions[None, :]
    b = actions.shape[0]
    e = len(edge_row)
    res0, flops0 = baseline_constants(a, a_stored_nnz, dtype=baseline_dtype)
    out_r = np.zeros(b)
    out_res = np.zeros(b)
    out_nnz = np.zeros(b, dtype=np.int64)
    ones = np.ones(e)
    for i in range(b):
        kept = kept_edge_mask(e, actions[i])
        pat = build_pattern_matrix(n, edge_row, edge_col, ones, kept, dtype=np.float64)
        res = residual_ls(n, pat, a, dtype=dtype)
        out_res[i] = res
        out_nnz[i] = pat.nnz
        out_r[i] = reward_from_residual(res, pat.nnz, n, res0, flops0, alpha)
    return {"reward": out_r, "residual": out_res, "nnz_m": out_nnz,
            "orig_residual": res0, "orig_flops": flops0}


# --------------------------------------------------------------------------
# masked categorical step (K4)
# --------------------------------------------------------------------------
def masked_softmax_probs(logits: np.ndarray, taken_ids) -> np.ndarray:
    """policy.py:64-73: logits[:A] with -inf at already-taken ids (index -1
    wraps to the terminal logit), softmax in fp32."""
    x = np.asarray(logits, dtype=np.float32).copy()
    ids = np.asarray(list(taken_ids), dtype=np.int64)
    if ids.size:
        x[ids] = -np.inf
    mx = np.max(x)
    ex = np.exp((x - mx).astype(np.float32)).astype(np.float32)
    return (ex / np.float32(ex.sum(dtype=np.float32))).astype(np.float32)


def sample_step(logits: np.ndarray, taken_lists, uniforms, done):
    """One environment step for B samples by inverse CDF with injected uniforms
    (gflownet.py:145-148 uses torch.multinomial, which cannot be reproduced
    bit-for-bit; the inverse-CDF rule is the comparable known-answer form).

    Returns (action i64[B] with -1 for finished rows — log.py:84-86,
             prob f32[B] with 1.0 for finished rows — log.py:67,78,
             done_out bool[B] — gflownet.py:177-179)."""
    logits = np.asarray(logits, dtype=np.float32)
    a = logits.shape[-1]
    bsz = len(uniforms)
    act = np.full(bsz, -1, dtype=np.int64)
    prob = np.ones(bsz, dtype=np.float32)
    done_out = np.asarray(done, dtype=bool).copy()
    for b in range(bsz):
        if done[b]:
            continue
        lg = logits if logits.ndim == 1 else logits[b]
        p = masked_softmax_probs(lg, taken_lists[b]).astype(np.float64)
        cdf = np.cumsum(p)
        target = float(uniforms[b]) * cdf[-1]
        idx = int(np.searchsorted(cdf, target, side="right"))
        idx = min(idx, a - 1)
        while p[idx] == 0.0 and idx > 0:     # never land on a masked id
            idx -= 1
        act[b] = idx
        prob[b] = np.float32(p[idx] / cdf[-1])
        done_out[b] = idx == a - 1
    return act, prob, done_out


# --------------------------------------------------------------------------
# whole trajectories by the exponential race (K4g)
# --------------------------------------------------------------------------
def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11;
    the Random123 reference implementation, the generator behind torch.cuda / curand). Not in the
    reference repository (it draws with torch.multinomial, gflownet.py:148, whose stream cannot be
    reproduced); this is the published algorithm, pinned by the Random123 known-answer vectors in
    tests/test_oracle.py. counter uint32[..., 4], key uint32[..., 2] -> uint32[..., 4]."""
    c = np.array(counter, dtype=np.uint64, copy=True)
    k = np.array(np.broadcast_to(np.asarray(key, dtype=np.uint64), c.shape[:-1] + (2,)), copy=True)
    m0, m1, mask = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = m0 * c[..., 0], m1 * c[..., 2]
        c = np.stack([(p1 >> np.uint64(32)) ^ c[..., 1] ^ k[..., 0], p1 & mask,
                      (p0 >> np.uint64(32)) ^ c[..., 3] ^ k[..., 1], p0 & mask], axis=-1)
        k = np.stack([(k[..., 0] + np.uint64(0x9E3779B9)) & mask, (k[..., 1] + np.uint64(0xBB67AE85)) & mask], axis=-1)
    return c.astype(np.uint32)


def race_keys(logits, seed: int, sample: int) -> np.ndarray:
    """Keys of the exponential race that is equal in distribution to the reference's step loop
    (gflownet.py:135-179 with the masked softmax of policy.py:64-73; Plackett-Luce / Gumbel-top-k):
    id i arrives at t_i = E_i / exp(logit_i); lq_i = log(t_i / t_terminal) in float64 (0 for the terminal
    id A-1). E_i = -log(u_i), u_i = (x_i + 1/2) / 2^32, x_i = word (i mod 4) of
    Philox4x32-10(counter = (i // 4, sample), key = seed)."""
    lg = np.asarray(logits, dtype=np.float64)
    a = lg.shape[0]
    groups = (a + 3) // 4
    g = np.arange(groups, dtype=np.uint64)
    ctr = np.stack([g & np.uint64(0xFFFFFFFF), g >> np.uint64(32),
                    np.full(groups, sample & 0xFFFFFFFF, dtype=np.uint64),
                    np.full(groups, (sample >> 32) & 0xFFFFFFFF, dtype=np.uint64)], axis=-1)
    x = philox4x32_10(ctr, np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint64)).reshape(-1)[:a]
    xf = x.astype(np.float64)
    upper = x >= np.uint32(0x80000000)
    e = np.where(upper, -np.log1p(-((4294967295.0 - xf) + 0.5) / 4294967296.0), -np.log((xf + 0.5) / 4294967296.0))
    arr = np.log(e) - lg
    lq = arr - arr[a - 1]
    lq[a - 1] = 0.0
    return lq


def race_trajectory(lq: np.ndarray) -> np.ndarray:
    """The trajectory the keys define: ids with lq < 0 in ascending (lq, id) order, then the terminal id."""
    lq = np.asarray(lq)
    a = lq.shape[0]
    ids = np.nonzero(lq[: a - 1] < 0)[0]
    order = ids[np.lexsort((ids, lq[ids]))]
    return np.concatenate([order, [a - 1]]).astype(np.int64)
