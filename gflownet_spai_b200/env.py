"""Drop-in `PreconditionerEnv` backed by libspai_b200.so (B200, sm_100a).

Mirrors the reference's environment (tonylizza/gflownet-spai):
  * protocol            gflownet/env.py:3-38
  * PreconditionerEnv   preconditioner.py:11-165 (same method names, argument
                        meaning, return conventions and error behaviour)
The arithmetic of `update` / `reward` / `calculate_residual` runs in CUDA through
the C ABI of include/spai_b200.h; PyTorch is used for device memory and streams
only. There is no CPU fallback: without the shared library or without a CUDA
device construction raises.
"""
from __future__ import annotations

import ctypes as C
from abc import ABC, abstractmethod
from typing import Sequence

import numpy as np
import torch

from . import _lib
from ._lib import F32, F64, MODES, SpaiError, SpaiInfo, SpaiTiming, check

__all__ = ["Env", "Data", "SpaiContext", "PreconditionerEnv"]


class Env(ABC):
    """gflownet/env.py:3-38."""

    @abstractmethod
    def update(self, s, actions):
        pass

    @abstractmethod
    def mask(self, s):
        pass

    @abstractmethod
    def reward(self, s):
        pass


class Data:
    """Struct stand-in for torch_geometric.data.Data (the reward path uses it as
    an attribute bag only: preconditioner.py:25, gflownet/utils.py:310-311)."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def __contains__(self, key):
        return getattr(self, key, None) is not None


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return C.c_void_p(t.ctypes.data)
    return C.c_void_p(t.data_ptr())


def _coo_parts(m: torch.Tensor, what: str):
    """(rows, cols, values f64) of a sparse COO tensor in STORED order
    (`_indices()` / `_values()`, preconditioner.py:23-24: no coalescing)."""
    if not isinstance(m, torch.Tensor) or not m.is_sparse:
        raise ValueError(f"The {what} must be a sparse tensor.")
    idx = m._indices().detach().cpu()
    val = m._values().detach().cpu()
    if idx.shape[0] != 2:
        raise ValueError(f"The {what} must be a 2-D sparse tensor.")
    r = np.ascontiguousarray(idx[0].numpy().astype(np.int64))
    c = np.ascontiguousarray(idx[1].numpy().astype(np.int64))
    v = np.ascontiguousarray(val.numpy().astype(np.float64))
    return r, c, v


class SpaiContext:
    """Owns one `spai_ctx` (device-resident pattern, CSR(A), gather plans)."""

    def __init__(self, n: int, edge_row, edge_col, edge_val, a_row, a_col, a_val, device: int = 0):
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise SpaiError("no CUDA device: the SPAI reward path has no CPU fallback")
        self.device = int(device)
        self.n = int(n)
        er = np.ascontiguousarray(edge_row, dtype=np.int64)
        ec = np.ascontiguousarray(edge_col, dtype=np.int64)
        ev = np.ascontiguousarray(edge_val, dtype=np.float64)
        ar = np.ascontiguousarray(a_row, dtype=np.int64)
        ac = np.ascontiguousarray(a_col, dtype=np.int64)
        av = np.ascontiguousarray(a_val, dtype=np.float64)
        if not (er.size == ec.size == ev.size and ar.size == ac.size == av.size):
            raise ValueError("COO arrays of unequal length")
        h = C.c_void_p()
        check(self._lib.spai_ctx_create(self.device, self.n, er.size, _ptr(er), _ptr(ec), _ptr(ev),
                                        ar.size, _ptr(ar), _ptr(ac), _ptr(av), C.byref(h)),
              "spai_ctx_create")
        self._h = h
        self.num_edges = int(er.size)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.spai_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ info
    def info(self) -> SpaiInfo:
        out = SpaiInfo()
        check(self._lib.spai_ctx_info(self._h, C.byref(out)), "spai_ctx_info")
        return out

    def set_workspace_limit(self, nbytes: int):
        check(self._lib.spai_ctx_set_workspace_limit(self._h, int(nbytes)), "set_workspace_limit")

    def set_deletion_hint(self, max_deletions: int):
        """Upper bound on the edges one trajectory removes, for reward_from_taken (which sees
        no action list): short trajectories then take the deletion-driven kernel. 0 = unknown."""
        check(self._lib.spai_ctx_set_deletion_hint(self._h, int(max_deletions)), "set_deletion_hint")

    def k3m_rows(self) -> tuple:
        """(rows with <= 16 candidates, rows with 17..32) served by the tensor-core copy kernel; (0, 0) before its records
        exist or when it does not apply (spai_ctx_k3m_rows)."""
        a, b = C.c_int64(), C.c_int64()
        check(self._lib.spai_ctx_k3m_rows(self._h, C.byref(a), C.byref(b)), "spai_ctx_k3m_rows")
        return int(a.value), int(b.value)

    def enable_timing(self, on: bool = True):
        check(self._lib.spai_ctx_enable_timing(self._h, int(on)), "enable_timing")

    def last_timing(self) -> SpaiTiming:
        out = SpaiTiming()
        check(self._lib.spai_ctx_last_timing(self._h, C.byref(out)), "last_timing")
        return out

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ reward
    def reward_batch(self, actions: torch.Tensor, alpha: float, mode: str = "copy",
                     dtype: torch.dtype = torch.float32, want=("reward", "residual", "nnz_m"),
                     lengths: torch.Tensor | None = None):
        """actions int64 (or int32) [B, T] (CPU -> host entry point with copies inside; CUDA ->
        device entry point). `lengths` (optional int32[B], same side as `actions`): valid length
        of every row; entries beyond it are padding the caller vouches for and are never read
        (spai_reward_batch_*_len). Returns dict of tensors on the same side as `actions`."""
        if actions.dtype not in (torch.int64, torch.int32) or actions.dim() != 2:
            raise ValueError("actions must be an int64 (or int32) [B, T] tensor")
        if actions.stride(1) != 1 and actions.shape[1] > 1:
            actions = actions.contiguous()
        b, t = actions.shape
        ld = actions.stride(0) if b > 1 else max(t, 1)
        if ld < t:
            actions = actions.contiguous()
            ld = t
        md = MODES[mode]
        dt = F32 if dtype == torch.float32 else F64
        on_dev = actions.is_cuda
        dev = actions.device if on_dev else torch.device("cpu")
        if on_dev and actions.device.index != self.device:
            raise ValueError("actions live on a different CUDA device than the context")
        if lengths is not None:
            if lengths.dtype != torch.int32 or lengths.numel() != b or lengths.device != actions.device:
                raise ValueError("lengths must be an int32 [B] tensor on the same device as actions")
            lengths = lengths.contiguous()
        elif actions.dtype == torch.int32 and not on_dev:
            raise ValueError("int32 host actions need lengths")
        out = {}
        out["reward"] = torch.empty(b, dtype=torch.float64, device=dev) if "reward" in want else None
        out["residual"] = torch.empty(b, dtype=torch.float64, device=dev) if "residual" in want else None
        out["nnz_m"] = torch.empty(b, dtype=torch.int64, device=dev) if "nnz_m" in want else None
        if lengths is None and actions.dtype == torch.int64:
            fn = self._lib.spai_reward_batch_dev if on_dev else self._lib.spai_reward_batch_host
            st = fn(self._h, _ptr(actions), b, t, ld, float(alpha), md, dt, _ptr(out["reward"]),
                    _ptr(out["residual"]), _ptr(out["nnz_m"]), self._stream())
        else:
            fn = self._lib.spai_reward_batch_dev_len if on_dev else self._lib.spai_reward_batch_host_len
            st = fn(self._h, _ptr(actions), actions.element_size(), _ptr(lengths), b, t, ld, float(alpha), md, dt,
                    _ptr(out["reward"]), _ptr(out["residual"]), _ptr(out["nnz_m"]), self._stream())
        check(st, "spai_reward_batch")
        return {k: v for k, v in out.items() if v is not None}

    def reward_rows(self, actions: torch.Tensor, row_begin: int, row_end: int, mode: str = "copy",
                    dtype: torch.dtype = torch.float32):
        """Partial evaluation of rows [row_begin, row_end) (row sharding across GPUs):
        (sum of squared row residuals f64[B], nnz(M) i64[B]) for CUDA `actions`."""
        if not actions.is_cuda or actions.dtype != torch.int64 or actions.dim() != 2:
            raise ValueError("actions must be a CUDA int64 [B, T] tensor")
        actions = actions.contiguous()
        b, t = actions.shape
        res2 = torch.empty(b, dtype=torch.float64, device=actions.device)
        nnz = torch.empty(b, dtype=torch.int64, device=actions.device)
        check(self._lib.spai_reward_rows_dev(self._h, _ptr(actions), b, t, max(t, 1), MODES[mode],
                                             F32 if dtype == torch.float32 else F64, int(row_begin), int(row_end),
                                             _ptr(res2), _ptr(nnz), self._stream()), "spai_reward_rows_dev")
        return res2, nnz

    def finalize_rewards(self, res2: torch.Tensor, nnz: torch.Tensor, alpha: float,
                         dtype: torch.dtype = torch.float32):
        """sqrt + mix formula on summed partials (see reward_rows)."""
        b = res2.numel()
        reward = torch.empty(b, dtype=torch.float64, device=res2.device)
        residual = torch.empty_like(reward)
        check(self._lib.spai_finalize_rewards_dev(self._h, _ptr(res2), _ptr(nnz), b, float(alpha),
                                                  F32 if dtype == torch.float32 else F64, _ptr(reward),
                                                  _ptr(residual), self._stream()), "spai_finalize_rewards_dev")
        return {"reward": reward, "residual": residual, "nnz_m": nnz}

    def reward_from_taken(self, taken: torch.Tensor, alpha: float, mode: str = "copy",
                          dtype: torch.dtype = torch.float32):
        """taken int32[B, words] CUDA tensor, EDGE order, bit set = edge removed."""
        if not taken.is_cuda or taken.dtype != torch.int32 or taken.dim() != 2 or not taken.is_contiguous():
            raise ValueError("taken must be a contiguous CUDA int32 [B, words] tensor")
        b, words = taken.shape
        reward = torch.empty(b, dtype=torch.float64, device=taken.device)
        residual = torch.empty_like(reward)
        nnz = torch.empty(b, dtype=torch.int64, device=taken.device)
        check(self._lib.spai_reward_from_taken_dev(
            self._h, _ptr(taken), b, words, float(alpha), MODES[mode],
            F32 if dtype == torch.float32 else F64, _ptr(reward), _ptr(residual), _ptr(nnz),
            self._stream()), "spai_reward_from_taken_dev")
        return {"reward": reward, "residual": residual, "nnz_m": nnz}

    def kept_mask(self, actions: torch.Tensor) -> torch.Tensor:
        """uint8[B, E] kept-edge mask in the caller's edge order (utils.py:323)."""
        acts = actions.to(device=f"cuda:{self.device}", dtype=torch.int64).contiguous()
        b, t = acts.shape
        out = torch.empty((b, self.num_edges), dtype=torch.uint8, device=acts.device)
        check(self._lib.spai_kept_mask_dev(self._h, _ptr(acts), b, t, max(t, 1), _ptr(out), self._stream()),
              "spai_kept_mask_dev")
        return out

    def row_index_sets(self, row: int):
        nj, ni = C.c_int64(0), C.c_int64(0)
        check(self._lib.spai_row_index_sets(self._h, int(row), C.byref(nj), None, C.byref(ni), None),
              "spai_row_index_sets")
        j = np.empty(nj.value, dtype=np.int64)
        i = np.empty(ni.value, dtype=np.int64)
        check(self._lib.spai_row_index_sets(self._h, int(row), C.byref(nj), _ptr(j), C.byref(ni), _ptr(i)),
              "spai_row_index_sets")
        return j, i

    def ls_solve_values(self, actions_row, dtype: torch.dtype = torch.float64) -> np.ndarray:
        """Values of M re-solved (ls mode) on the pattern of ONE trajectory, float64[E]
        in the caller's edge order; 0 for removed edges."""
        acts = np.ascontiguousarray(np.asarray(actions_row, dtype=np.int64).ravel())
        out = np.zeros(self.num_edges, dtype=np.float64)
        check(self._lib.spai_ls_solve_values_host(self._h, _ptr(acts), acts.size,
                                                  F32 if dtype == torch.float32 else F64, _ptr(out),
                                                  self._stream()), "spai_ls_solve_values_host")
        return out

    def pack_taken(self, keys: torch.Tensor):
        """keys f32[B, A] (CUDA) -> (taken int32[B, words], length int32[B]); see spai_pack_taken_dev."""
        b, a = keys.shape
        words = (a + 31) // 32
        taken = torch.empty((b, words), dtype=torch.int32, device=keys.device)
        length = torch.empty(b, dtype=torch.int32, device=keys.device)
        check(self._lib.spai_pack_taken_dev(self._h, _ptr(keys), keys.stride(0), a, b, _ptr(taken), words,
                                            _ptr(length), self._stream()), "spai_pack_taken_dev")
        return taken, length

    def sample_taken(self, logits: torch.Tensor, bsz: int, seed: int, sample0: int = 0, export_keys: bool = False):
        """Whole trajectories by the exponential race (K4g, spai_sample_taken_dev): logits f32[A] (CUDA) ->
        (taken int32[B, words], length int32[B]) and, with export_keys (tests), the keys f32[B, A]."""
        a = logits.numel()
        words = (a + 31) // 32
        dev = logits.device
        taken = torch.empty((bsz, words), dtype=torch.int32, device=dev)
        length = torch.empty(bsz, dtype=torch.int32, device=dev)
        keys = torch.empty((bsz, a), dtype=torch.float32, device=dev) if export_keys else None
        check(self._lib.spai_sample_taken_dev(self.device, _ptr(logits), a, bsz, int(seed) & (2 ** 64 - 1), int(sample0),
                                              _ptr(taken), words, _ptr(length), _ptr(keys) if export_keys else None,
                                              a if export_keys else 0, self._stream()), "spai_sample_taken_dev")
        return (taken, length, keys) if export_keys else (taken, length)

    def sample_order(self, logits: torch.Tensor, length: torch.Tensor, seed: int, sample0: int = 0,
                     dtype: torch.dtype = torch.int32, ld: int = 0) -> torch.Tensor:
        """The drawn ids in draw order (K4g, spai_sample_order_dev): int32/int64 [B, ld] with the terminal id last
        and -1 padding; `length` from sample_taken with the same logits / seed / sample0."""
        bsz = length.numel()
        if ld <= 0:
            ld = int(length.max()) if bsz else 1
        out = torch.empty((bsz, ld), dtype=dtype, device=logits.device)
        check(self._lib.spai_sample_order_dev(self.device, _ptr(logits), logits.numel(), bsz, int(seed) & (2 ** 64 - 1),
                                              int(sample0), _ptr(length), _ptr(out), out.element_size(), ld,
                                              self._stream()), "spai_sample_order_dev")
        return out

    def sample_steps(self, logits: torch.Tensor, taken: torch.Tensor, done: torch.Tensor, nsteps: int,
                     uniforms: torch.Tensor | None = None, seed: int = 0, sample0: int = 0, step0: int = 0,
                     actions: torch.Tensor | None = None, probs: torch.Tensor | None = None,
                     dtype: torch.dtype = torch.int32, want_probs: bool = True):
        """Up to `nsteps` masked-categorical steps per sample in one launch (K4p, spai_sample_steps_dev).
        taken int32[B, words] / done uint8[B] are updated in place; returns (actions [B, ld], probs f32[B, ld] or
        None, steps_taken int32[B]); columns step0 .. step0 + nsteps - 1 are written."""
        bsz = taken.shape[0]
        dev = logits.device
        if actions is None:
            actions = torch.empty((bsz, step0 + nsteps), dtype=dtype, device=dev)
        if probs is None and want_probs:
            probs = torch.empty((bsz, actions.shape[1]), dtype=torch.float32, device=dev)
        steps = torch.empty(bsz, dtype=torch.int32, device=dev)
        check(self._lib.spai_sample_steps_dev(self.device, _ptr(logits), logits.numel(), bsz, _ptr(taken), taken.shape[1],
                                              _ptr(done), _ptr(uniforms) if uniforms is not None else None,
                                              int(seed) & (2 ** 64 - 1), int(sample0), int(step0), int(nsteps), _ptr(actions),
                                              actions.element_size(), _ptr(probs) if probs is not None else None,
                                              actions.shape[1], _ptr(steps), self._stream()), "spai_sample_steps_dev")
        return actions, probs, steps

    def sample_step(self, logits, taken, uniforms, done, action, prob):
        """In-place masked categorical step on CUDA tensors (include/spai_b200.h)."""
        a = logits.shape[-1]
        ld = 0 if logits.dim() == 1 else logits.stride(0)
        b = taken.shape[0]
        check(self._lib.spai_sample_step_dev(self._h, _ptr(logits), ld, a, _ptr(taken), taken.shape[1],
                                             _ptr(uniforms), _ptr(done), b, _ptr(action), _ptr(prob),
                                             self._stream()), "spai_sample_step_dev")


def residual_pair(n, m_row, m_col, m_val, a_row, a_col, a_val, dtype=torch.float32, device=0):
    """||M @ A - I||_F for arbitrary COO pairs (preconditioner.py:79-93) on the GPU."""
    lib = _lib.load()
    arrs = [np.ascontiguousarray(x, dtype=np.int64) for x in (m_row, m_col)]
    mv = np.ascontiguousarray(m_val, dtype=np.float64)
    brrs = [np.ascontiguousarray(x, dtype=np.int64) for x in (a_row, a_col)]
    av = np.ascontiguousarray(a_val, dtype=np.float64)
    res = C.c_double(0.0)
    nnz = C.c_int64(0)
    check(lib.spai_residual_pair_host(int(device), int(n), mv.size, _ptr(arrs[0]), _ptr(arrs[1]), _ptr(mv),
                                      av.size, _ptr(brrs[0]), _ptr(brrs[1]), _ptr(av),
                                      F32 if dtype == torch.float32 else F64, C.byref(res), C.byref(nnz)),
          "spai_residual_pair_host")
    return res.value, nnz.value


class PreconditionerEnv(Env):
    """preconditioner.py:11-165 with the reward evaluated on the GPU.

    Extra keyword arguments (not in the reference): ``device`` (CUDA index),
    ``mode`` ("copy" = the reference's semantics, "ls" = re-solve each row's
    least-squares problem by Householder QR, "ls_gram" = the same residual
    through per-row Gram matrices, much faster) and ``dtype`` (torch.float32 = the reference's
    precision, torch.float64).

    ``alpha``: the reference passes alpha to update()/reward() but reads the
    never-assigned ``self.alpha`` (preconditioner.py:163). Here ``self.alpha``
    wins when the caller has set it (the reference's behaviour once the
    attribute exists), otherwise the argument is used.
    """

    def __init__(self, matrix_size: int, initial_matrix: torch.Tensor, original_matrix: torch.Tensor,
                 device: int | None = None, mode: str = "copy", dtype: torch.dtype = torch.float32):
        if mode not in MODES:
            raise ValueError(f"mode must be one of {sorted(MODES)}")
        self.matrix_size = int(matrix_size)
        er, ec, ev = _coo_parts(initial_matrix, "initial matrix")
        ar, ac, av = _coo_parts(original_matrix, "original matrix")
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.device = int(device)
        self.mode = mode
        self.dtype = dtype
        self.alpha = None
        self.ctx = SpaiContext(self.matrix_size, er, ec, ev, ar, ac, av, device=self.device)
        info = self.ctx.info()
        self.init_nnz = int(info.init_nnz)                    # preconditioner.py:14
        self.state_dim = self.init_nnz                        # :15
        self.num_actions = self.init_nnz + 1                  # :16
        self.matrix = initial_matrix.clone()                  # :17
        self.original_matrix = original_matrix.clone()        # :18
        self.data = Data(edge_index=self.matrix._indices(),   # :23-25
                         edge_attr=self.matrix._values().float())
        res0 = info.orig_residual_f32 if dtype == torch.float32 else info.orig_residual_f64
        self.orig_residual = torch.tensor(res0, dtype=torch.float64)   # :28
        self.orig_flops = int(info.orig_flops)                          # :29
        self._a_coo = (ar, ac, av)

    # ------------------------------------------------------------------ hot call
    def _alpha(self, alpha) -> float:
        a = self.alpha if self.alpha is not None else alpha
        return float(a.detach()) if isinstance(a, torch.Tensor) else float(a)

    def update_tensor(self, actions, alpha, want=("reward", "residual", "nnz_m"), lengths=None):
        """Batch reward as tensors (no per-trajectory Python objects). `lengths` (optional
        int32[B]): valid length of every row, as the sampler's Log knows it (log.py:84-87);
        the -1 padding behind it is then never read by host or device."""
        if not isinstance(actions, torch.Tensor):
            rows = [list(map(int, r)) for r in actions]
            t = max((len(r) for r in rows), default=0)
            actions = torch.tensor([r + [-1] * (t - len(r)) for r in rows], dtype=torch.int64).reshape(len(rows), t)
        if actions.dim() == 1:
            actions = actions.unsqueeze(0)
        if actions.dtype != torch.int32:
            actions = actions.to(torch.int64)
        return self.ctx.reward_batch(actions, self._alpha(alpha), self.mode, self.dtype, want, lengths=lengths)

    def update_from_taken(self, taken: torch.Tensor, alpha, max_deletions: int = 0):
        """Batch reward straight from the sampler's device-resident taken-bitmask
        (int32 [B, words], edge order, bit set = edge removed): no action lists,
        no host round trip. `max_deletions` (optional) bounds the edges a trajectory
        removes; it only lets the library pick the kernel for short trajectories."""
        self.ctx.set_deletion_hint(int(max_deletions))
        return self.ctx.reward_from_taken(taken, self._alpha(alpha), self.mode, self.dtype)

    def update(self, sparse_matrices, actions, alpha) -> list:
        """preconditioner.py:32-52: rewards of a batch of trajectories.

        `sparse_matrices` is ignored, as in the reference. Returns a list of B
        0-dim float64 tensors (what `torch.tensor(rewards, dtype=float32)` at
        gflownet/gflownet.py:193 consumes)."""
        out = self.update_tensor(actions, alpha, want=("reward",))["reward"]
        return list(out.detach().cpu().unbind(0))

    # ------------------------------------------------------------------ secondary API
    def reward(self, s: torch.Tensor, traj_length: int, alpha) -> torch.Tensor:
        """preconditioner.py:55-66 for one explicit matrix `s`."""
        metric = self.evaluate_preconditioner(s, self.original_matrix, self.orig_residual, self.orig_flops, alpha)
        return torch.as_tensor(metric).to(torch.float64) * 1000

    def matrix_flops(self, matrix: torch.Tensor):
        """preconditioner.py:68-77."""
        if matrix.is_sparse:
            non_zeros = matrix._values().numel()
            flops = non_zeros * matrix.shape[1] * 2
        else:
            non_zeros = torch.nonzero(matrix).size(0)
            flops = 2 * non_zeros
        return flops, non_zeros

    def calculate_residual(self, updated_matrix: torch.Tensor, original_matrix: torch.Tensor) -> torch.Tensor:
        """preconditioner.py:79-93: ||M @ A - I||_F, 0-dim float64 tensor."""
        mr, mc, mv = _coo_parts(updated_matrix, "updated matrix")
        if original_matrix is self.original_matrix:
            ar, ac, av = self._a_coo
        else:
            ar, ac, av = _coo_parts(original_matrix, "original matrix")
        res, _ = residual_pair(self.matrix_size, mr, mc, mv, ar, ac, av, dtype=self.dtype, device=self.device)
        return torch.tensor(res, dtype=torch.float64)

    def mask(self, s) -> torch.Tensor:
        """preconditioner.py:95-98."""
        return torch.ones(len(s), self.num_actions)

    def create_mask_from_sparse_matrix(self, sparse_matrix: torch.Tensor) -> torch.Tensor:
        """preconditioner.py:101-135 (dense n*n mask; small n only)."""
        if not sparse_matrix.is_sparse:
            raise ValueError("The input tensor must be a sparse tensor.")
        size = sparse_matrix.size()
        mask = torch.zeros(size, dtype=torch.float32)
        indices = sparse_matrix._indices()
        mask[indices[0], indices[1]] = 1
        resized = mask.view(1, size[0] * size[1])
        return torch.cat((resized, torch.tensor([[1]], dtype=torch.float32)), dim=1)

    def evaluate_preconditioner(self, updated_matrix, original_matrix, orig_residual, orig_flops, alpha):
        """preconditioner.py:137-165."""
        residual = self.calculate_residual(updated_matrix, original_matrix)
        flops, _ = self.matrix_flops(updated_matrix)
        residual_ratio = residual / orig_residual if orig_residual != 0 else float("inf")
        computational_ratio = flops / orig_flops if orig_flops != 0 else float("inf")
        a = self._alpha(alpha)
        return a * (1 - residual_ratio) + (1 - a) * (1 - computational_ratio)
