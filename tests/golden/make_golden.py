"""Generate golden vectors by running the REFERENCE ITSELF (read-only checkout
at /root/reference, imported through oracle/ref_shim.py). Run in the build
container only:  python tests/golden/make_golden.py

Each ``<case>.npz`` holds the inputs (fp32 COO of the initial and original
matrices in the caller's entry order, actions[B,T], alpha) and the reference's
outputs: rewards, baseline constants and, per trajectory, the coalesced
(row, col, value) of M (utils.py:124) as ragged arrays.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from gflownet_spai_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402


def run_case(name, n, er, ec, ev, ar, ac, av, actions, alpha, alpha_tensor=False):
    er, ec, ar, ac = (np.asarray(x, dtype=np.int64) for x in (er, ec, ar, ac))
    ev = np.asarray(ev, dtype=np.float32)
    av = np.asarray(av, dtype=np.float32)
    actions = np.asarray(actions, dtype=np.int64)
    out = ref_shim.reference_update(n, er, ec, ev, ar, ac, av, actions, alpha, alpha_tensor=alpha_tensor)
    m_ptr, m_row, m_col, m_val = [0], [], [], []
    for b in range(actions.shape[0]):
        _, r, c, v = ref_shim.reference_masks_and_indices(n, er, ec, ev, actions[b])
        m_row.append(r), m_col.append(c), m_val.append(v)
        m_ptr.append(m_ptr[-1] + r.size)
    np.savez_compressed(
        os.path.join(HERE, f"{name}.npz"),
        n=n, edge_row=er, edge_col=ec, edge_val=ev, a_row=ar, a_col=ac, a_val=av,
        actions=actions, alpha=np.float64(alpha), reward=out["reward"],
        orig_residual=np.float64(out["orig_residual"]), orig_flops=np.int64(out["orig_flops"]),
        init_nnz=np.int64(out["init_nnz"]), num_actions=np.int64(out["num_actions"]),
        m_ptr=np.asarray(m_ptr, dtype=np.int64),
        m_row=np.concatenate(m_row) if m_row else np.zeros(0, np.int64),
        m_col=np.concatenate(m_col) if m_col else np.zeros(0, np.int64),
        m_val=np.concatenate(m_val) if m_val else np.zeros(0, np.float32))
    print(f"{name}: B={actions.shape[0]} E={er.size} res0={out['orig_residual']!r} "
          f"flops0={out['orig_flops']} rewards[:3]={out['reward'][:3]}")


def pad(rows, fill=-1):
    t = max(len(r) for r in rows)
    return np.array([list(r) + [fill] * (t - len(r)) for r in rows], dtype=np.int64)


def inf_cases():
    """Zero baselines: `float('inf')` ratios of preconditioner.py:154 and :158."""
    rng = np.random.default_rng(99)
    n = 6
    er = rng.integers(0, n, 14)
    ec = rng.integers(0, n, 14)
    ev = rng.uniform(-1, 1, 14)
    acts = pad([[14], [0, 3, 14], list(range(14)) + [14]])
    # 6. A0 = identity: ||A0 A0 - I||_F = 0  ->  residual ratio = inf (:154)
    run_case("zero_res0", n, er, ec, ev, np.arange(n), np.arange(n), np.ones(n), acts, 0.5, alpha_tensor=True)
    # 7. A0 with no stored entry: orig_flops = 0  ->  computational ratio = inf (:158)
    run_case("zero_flops0", n, er, ec, ev, np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0), acts, 0.5,
             alpha_tensor=True)


def main():
    if "--inf-only" in sys.argv:
        inf_cases()
        return
    # 1. the 3x3 known-answer case of SURVEY.md §4 (explicit zero, unsorted order)
    er = [2, 0, 1, 0, 2, 1]
    ec = [2, 0, 1, 1, 0, 0]
    ev = [4, 2, 3, -1, 0, -1]
    acts = pad([[6], [0, 1, 2, 3, 4, 5, 6], [1, 1, 6], [3, 6], [6, -1, -1]])
    run_case("tiny3", 3, er, ec, ev, er, ec, ev, acts, 0.5)

    # 2. cfg1: 2-D Poisson 10x10, initial = original = A, B = 32
    p = synth.make_problem("cfg1")
    coo = p.a.tocoo()
    acts = synth.make_trajectories(p.num_edges, 32)
    run_case("poisson10", p.n, p.edge_row, p.edge_col, p.edge_val, coo.row, coo.col, coo.data,
             acts, 0.5)

    # 3. non-symmetric convection-diffusion 16x16, superset pattern k=8 with
    #    non-integer initial values (precision-sensitive)
    a = synth.convdiff2d(16)
    r, c = synth.superset_pattern(a, 8, max_power=2)
    v = synth.neumann_values(a, r, c, terms=3)
    coo = a.tocoo()
    acts = synth.make_trajectories(r.size, 8, seed0=77)
    run_case("convdiff16", a.shape[0], r, c, v, coo.row, coo.col, coo.data, acts, 0.3)

    # 4. uncoalesced caller order: shuffled entries, duplicate coordinates,
    #    explicit zeros; actions with duplicates, out-of-range ids, all edges
    #    removed, nothing removed
    rng = np.random.default_rng(2024)
    n = 40
    e0 = 150
    er = rng.integers(0, n, e0)
    ec = rng.integers(0, n, e0)
    ev = rng.uniform(-2, 2, e0)
    ev[rng.integers(0, e0, 6)] = 0.0
    dup = rng.integers(0, e0, 12)
    er = np.concatenate([er, er[dup]])
    ec = np.concatenate([ec, ec[dup]])
    ev = np.concatenate([ev, rng.uniform(-2, 2, dup.size)])
    perm = rng.permutation(er.size)
    er, ec, ev = er[perm], ec[perm], ev[perm]
    a0 = 120
    ar = np.concatenate([np.arange(n), rng.integers(0, n, a0)])
    ac = np.concatenate([np.arange(n), rng.integers(0, n, a0)])
    av = np.concatenate([rng.uniform(1, 3, n), rng.uniform(-1, 1, a0)])
    e = er.size
    # the reference's terminal id is init_nnz (coalesced count) < E here
    rows = [
        [e],
        list(range(e)) + [e],
        [5, 5, 9, 9, 9, e + 7, e],
        list(rng.permutation(e)[:60]) + [e],
        list(rng.permutation(e)[:10]) + [e],
        list(rng.permutation(e)[:140]) + [e],
        [0, e - 1, e],
        list(rng.permutation(e)[:90]),
    ]
    run_case("uncoalesced40", n, er, ec, ev, ar, ac, av, pad(rows), 0.7)

    # 5. 2-D Poisson 32x32 with the k=8 superset (the cfg2 shape, small)
    a = synth.poisson2d(32)
    r, c = synth.superset_pattern(a, 8, max_power=2)
    v = synth.neumann_values(a, r, c, terms=3)
    coo = a.tocoo()
    acts = synth.make_trajectories(r.size, 6, seed0=5)
    run_case("poisson32_k8", a.shape[0], r, c, v, coo.row, coo.col, coo.data, acts, 0.5)
    inf_cases()


if __name__ == "__main__":
    main()
