"""Import the live reference (read-only checkout) as an oracle — TEST
INFRASTRUCTURE. Uses ``/root/reference`` in the build container; on the GPU box
(where that path does not exist) the unmodified copy staged under ``baseline/_ref`` by
``__graft_entry__.build()`` (git-ignored, travels with the snapshot) — only ``bench.py
--impl reference`` reads it there, to time the reference itself.

Two shims:
  1. ``torch_geometric`` is not installed: inject a container-only
     ``torch_geometric.data.Data`` (attribute bag with ``__contains__``) and
     placeholder ``torch_geometric.nn`` names so ``preconditioner.py``,
     ``policy.py`` and ``gflownet/*.py`` import.
  2. ``PreconditionerEnv.evaluate_preconditioner`` reads ``self.alpha``
     (preconditioner.py:163) which is never assigned: callers set ``env.alpha``.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_STAGED = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")      # copy made by __graft_entry__.build()


def _default_root() -> str:
    env = os.environ.get("SPAI_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile("/root/reference/preconditioner.py"):
        return "/root/reference"
    return _STAGED


REFERENCE_ROOT = _default_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "preconditioner.py"))


def _install_pyg_stub() -> None:
    if "torch_geometric" in sys.modules:
        return

    class Data:  # container only; the reward path uses it as a struct
        def __init__(self, **kw):
            for k, v in kw.items():
                setattr(self, k, v)

        def __contains__(self, key):
            return hasattr(self, key) and getattr(self, key) is not None

    class _Unavailable:
        def __init__(self, *a, **k):
            raise RuntimeError("torch_geometric.nn is not installed in this container")

    def _global_mean_pool(*a, **k):
        raise RuntimeError("torch_geometric.nn is not installed in this container")

    tg = types.ModuleType("torch_geometric")
    tgd = types.ModuleType("torch_geometric.data")
    tgn = types.ModuleType("torch_geometric.nn")
    tgd.Data = Data
    tgn.GATv2Conv = _Unavailable
    tgn.global_mean_pool = _global_mean_pool
    tg.data = tgd
    tg.nn = tgn
    sys.modules["torch_geometric"] = tg
    sys.modules["torch_geometric.data"] = tgd
    sys.modules["torch_geometric.nn"] = tgn


def load_reference():
    """Returns a namespace with the reference's PreconditionerEnv, GFlowNet, Log,
    trajectory_balance_loss, BackwardPolicy and helper functions."""
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    _install_pyg_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import preconditioner as ref_pre          # noqa: E402
    import policy as ref_policy               # noqa: E402
    from gflownet import gflownet as ref_gfn  # noqa: E402
    from gflownet import log as ref_log       # noqa: E402
    from gflownet import utils as ref_utils   # noqa: E402
    ns = types.SimpleNamespace(
        PreconditionerEnv=ref_pre.PreconditionerEnv,
        GFlowNet=ref_gfn.GFlowNet,
        Log=ref_log.Log,
        BackwardPolicy=ref_policy.BackwardPolicy,
        trajectory_balance_loss=ref_utils.trajectory_balance_loss,
        update_edges_and_convert_to_sparse=ref_utils.update_edges_and_convert_to_sparse,
        resize_sparse_tensor=ref_utils.resize_sparse_tensor,
        modules=(ref_pre, ref_policy, ref_gfn, ref_log, ref_utils),
    )
    return ns


def reference_update(n, edge_row, edge_col, edge_val, a_row, a_col, a_val, actions, alpha,
                     quiet: bool = True, alpha_tensor: bool = False):
    """Run the reference's own PreconditionerEnv.update on fp32 COO inputs.

    Returns dict(reward f64[B], orig_residual, orig_flops, init_nnz, num_actions)."""
    import numpy as np
    import torch

    ref = load_reference()
    init = torch.sparse_coo_tensor(
        torch.tensor(np.stack([edge_row, edge_col]), dtype=torch.long),
        torch.tensor(np.asarray(edge_val), dtype=torch.float32), (n, n))
    orig = torch.sparse_coo_tensor(
        torch.tensor(np.stack([a_row, a_col]), dtype=torch.long),
        torch.tensor(np.asarray(a_val), dtype=torch.float32), (n, n))
    env = ref.PreconditionerEnv(n, init, orig)
    # alpha_tensor: keep alpha a 0-dim tensor as the live sampler passes it (gflownet.py:183); with
    # a Python float the reference's inf-ratio branch (preconditioner.py:154,:158) dies at :64
    # ('float' object has no attribute 'to')
    env.alpha = torch.tensor(float(alpha)) if alpha_tensor else float(alpha)
    acts = torch.tensor(np.asarray(actions), dtype=torch.long)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext():
        rewards = env.update([init] * acts.shape[0], acts, torch.tensor(float(alpha)))
    return {
        "reward": np.array([float(r) for r in rewards], dtype=np.float64),
        "orig_residual": float(env.orig_residual),
        "orig_flops": int(env.orig_flops),
        "init_nnz": int(env.init_nnz),
        "num_actions": int(env.num_actions),
    }


def reference_masks_and_indices(n, edge_row, edge_col, edge_val, actions_row):
    """Bit-exact artefacts from the reference helpers:
    kept-edge mask (utils.py:323), coalesced (row, col) of M (utils.py:124)."""
    import numpy as np
    import torch

    ref = load_reference()
    from torch_geometric.data import Data
    data = Data(edge_index=torch.tensor(np.stack([edge_row, edge_col]), dtype=torch.long),
                edge_attr=torch.tensor(np.asarray(edge_val), dtype=torch.float32))
    good = [x for x in torch.tensor(np.asarray(actions_row), dtype=torch.long) if x != -1]
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        flat = ref.update_edges_and_convert_to_sparse(data, good, n)
        m = ref.resize_sparse_tensor(flat, (n, n))
    acts = {int(a) for a in good}
    kept = np.array([i not in acts for i in range(len(edge_row))], dtype=bool)
    idx = m._indices().numpy()
    return kept, idx[0].copy(), idx[1].copy(), m._values().numpy().copy()
