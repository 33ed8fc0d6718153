"""Drop-in `GFlowNet` sampler, `Log` and trajectory-balance loss.

Mirrors the reference's
  * gflownet/gflownet.py:12-197  (GFlowNet, forward_probs, sample_states)
  * gflownet/log.py:10-164       (Log)
  * gflownet/utils.py:228-278    (trajectory_balance_loss)
with the masked-categorical environment step (policy.py:64-73,
gflownet.py:116-119,:148,:177-179, log.py:67-87) executed by CUDA kernel K4 and
the batch reward by the SPAI reward kernels.

Why one policy call per epoch is enough: inside `sample_states` the graph and
the weights are identical for every sample and every step (states are never
mutated, gflownet.py:164-172), so the policy's pre-mask distribution p is one
vector; masking taken ids and re-normalising gives, for an untaken id a,
    softmax(masked logits)[a] = p[a] / (1 - sum_{taken} p).
The sampler therefore evaluates the policy once (with autograd), draws all
steps on the device from log(p), and rebuilds the chosen-action probabilities
as a differentiable function of p for the loss.
"""
from __future__ import annotations

from typing import List

import torch
from torch import nn

from .env import Data

__all__ = ["GFlowNet", "Log", "trajectory_balance_loss", "BackwardPolicy"]


def trajectory_balance_loss(total_flow, rewards, fwd_probs, back_probs):
    """Mean trajectory-balance loss with the reference's stabilisation
    (gflownet/utils.py:228-278): eps = 1e-9 inside every log, per-side
    subtraction of the batch maximum of the summed log-probabilities."""
    eps = 1e-9
    dt, dev = fwd_probs.dtype, fwd_probs.device
    total_flow = total_flow.to(dev).to(dt)
    rewards = rewards.to(dev).to(dt)
    back_probs = back_probs.to(dev).to(dt)
    lf = torch.log(fwd_probs + eps).sum(dim=-1)
    lb = torch.log(back_probs + eps).sum(dim=-1)
    lf = lf - lf.max(dim=0, keepdim=True)[0]
    lb = lb - lb.max(dim=0, keepdim=True)[0]
    lhs = torch.log(total_flow + eps) + lf
    rhs = torch.log(rewards + eps) + lb
    return ((lhs - rhs) ** 2).mean()


class BackwardPolicy(nn.Module):
    """Drop-in for the reference's backward policy (policy.py:75-129) with the per-trajectory
    Python loop replaced by ONE packed LSTM call over the batch.

    Same constructor, same parameter names (`lstm`, `fc`: a reference state_dict loads
    as is), same output: for trajectory b with `len_b` valid (non -1) leading entries,
    softmax of the first `len_b` outputs of `fc(h_last)` followed by ones up to the
    maximum length; shape [B, 1, T] like the reference's stack of [1, T] rows.
    The reference feeds each action id as a 1-feature float sequence (`input_dim` must be 1).
    """

    def __init__(self, input_dim: int, hidden_dim: int, max_num_actions: int):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.max_num_actions = max_num_actions
        self.lstm = nn.LSTM(input_dim, hidden_dim, batch_first=True)
        self.fc = nn.Linear(hidden_dim, max_num_actions)

    def forward(self, trajectories: torch.Tensor) -> torch.Tensor:
        bsz, max_len = trajectories.shape
        lengths = (trajectories != -1).sum(dim=1)                     # policy.py:96-100 (valid prefix)
        if bool((lengths == 0).any()):
            raise RuntimeError("BackwardPolicy: a trajectory has no valid action "
                               "(the reference fails in pack_padded_sequence for length 0)")
        packed = nn.utils.rnn.pack_padded_sequence(trajectories.float().unsqueeze(-1), lengths.cpu(),
                                                   batch_first=True, enforce_sorted=False)
        out, _ = self.lstm(packed)
        out, _ = nn.utils.rnn.pad_packed_sequence(out, batch_first=True)      # [B, max(len), H]
        idx = (lengths - 1).to(out.device).view(-1, 1, 1).expand(bsz, 1, out.size(2))
        last = out.gather(1, idx).squeeze(1)                                  # hidden state at the last valid step
        logits = self.fc(last)                                                # [B, max_num_actions]
        width = min(max_len, self.max_num_actions)
        pos = torch.arange(width, device=logits.device)[None, :]
        valid = pos < lengths.to(logits.device)[:, None]                      # slice [:len_b] (policy.py:119)
        probs = torch.softmax(logits[:, :width].masked_fill(~valid, float("-inf")), dim=1)
        probs = torch.where(valid, probs, torch.ones_like(probs))             # pad with 1 (policy.py:125)
        if width < max_len:
            probs = torch.nn.functional.pad(probs, (0, max_len - width), value=1.0)
        return probs.unsqueeze(1)


class Log:
    """gflownet/log.py:10-164: per-step chosen-action probabilities and actions
    (-1 once a sample has finished), rewards, lazy backward probabilities."""

    def __init__(self, s0, backward_policy, total_flow, env):
        self._fwd_probs = []
        self._back_probs = None
        self._actions = []
        self.rewards = torch.zeros(len(s0))
        self.backward_policy = backward_policy
        self.total_flow = total_flow
        self.env = env
        self.num_samples = len(s0)

    def log(self, s, probs: torch.Tensor, actions: torch.Tensor, done: torch.Tensor):
        """log.py:24-89 (host tensors; the device sampler fills the log directly)."""
        chosen = probs.gather(2, actions.unsqueeze(1)).reshape(-1)
        active = ~done.flatten().bool()
        fwd = torch.ones(actions.shape[0], device=actions.device, dtype=chosen.dtype)
        fwd[active] = chosen[active]
        self._fwd_probs.append(fwd)
        rec = -torch.ones(self.num_samples, dtype=torch.long)
        rec[active] = actions.reshape(-1)[active]
        self._actions.append(rec)

    @property
    def fwd_probs(self):
        if isinstance(self._fwd_probs, list):
            self._fwd_probs = torch.stack(self._fwd_probs, dim=0).t()
        return self._fwd_probs

    @property
    def actions(self):
        if isinstance(self._actions, list):
            self._actions = torch.stack(self._actions, dim=0)
        return self._actions

    @property
    def back_probs(self):
        if self._back_probs is None:
            back = self.backward_policy(self.actions.t())
            self._back_probs = back.reshape(self.num_samples, -1)
        return self._back_probs


class GFlowNet(nn.Module):
    """gflownet/gflownet.py:12-197 with the step loop on the GPU."""

    def __init__(self, forward_policy, backward_policy, env):
        super().__init__()
        self.register_buffer("total_flow", torch.ones(1))      # gflownet.py:16 (not learned)
        self.forward_policy = forward_policy
        self.backward_policy = backward_policy
        self.env = env
        self.check_every = 8

    # ------------------------------------------------------------------ reference protocol
    def state_to_data(self, s: List[torch.Tensor]) -> list:
        """gflownet.py:223-257."""
        out = []
        for i, m in enumerate(s):
            if not m.is_sparse:
                raise ValueError(f"Tensor at index {i} is not a sparse tensor.")
            x = torch.ones((self.env.matrix_size * 2, 1))
            out.append(Data(x=x, edge_index=m._indices(), edge_attr=m._values().float()))
        return out

    def forward_probs(self, s, data_list, actions=None):
        """gflownet.py:47-123: ([B, 1, A] probabilities, mean alpha); one policy
        call per sample, renormalised only when B > 1."""
        if actions is None or len(actions) == 0:
            acts = torch.empty(0)
        else:
            acts = torch.tensor(list(zip(*actions)), dtype=torch.long)
        probs, alphas = [], []
        for i, data in enumerate(data_list):
            a_i = acts[i, :] if acts.numel() > 0 else torch.empty(0, dtype=torch.long)
            p, al = self.forward_policy(data, a_i)
            probs.append(p)
            alphas.append(al)
        probs = torch.stack(probs, dim=0)
        alpha = torch.stack(alphas, dim=0).mean()
        if probs.size(0) > 1:
            tot = probs.sum(2)
            tot = torch.where(tot == 0, torch.ones_like(tot), tot)
            probs = probs / tot.unsqueeze(1)
        return probs, alpha

    # ------------------------------------------------------------------ device sampler
    @staticmethod
    def chosen_probs(p: torch.Tensor, actions_bt: torch.Tensor, rest: torch.Tensor | None = None) -> torch.Tensor:
        """Differentiable probabilities of the drawn actions: p[a_t] divided by the
        mass not yet taken before step t (what the reference's per-step masked softmax
        returns, policy.py:64-73); 1.0 where the action is -1.

        The remaining mass is built WITHOUT cancellation: (mass of the ids the trajectory
        never takes) + (suffix sum of the drawn ids' masses from step t on). `1 - cumsum`
        loses every digit on long trajectories because an fp32 softmax sums to 1 only to
        ~1e-6; here the denominator is >= p[a_t] by construction, so the result is in (0, 1]
        for any policy. `rest` (optional, f64[B]) is the never-taken mass when the caller has
        it exactly (GFlowNet.untaken_mass over the taken-bitmask); otherwise it is
        sum(p) - sum(drawn), clamped at 0 (exact to 1e-16 * sum(p))."""
        valid = actions_bt >= 0
        idx = actions_bt.clamp(min=0)
        p64 = p.to(torch.float64)
        pa = p64[idx] * valid
        if rest is None:
            rest = (p64.sum() - pa.sum(dim=1)).clamp_min(0.0)
        suffix = torch.flip(torch.cumsum(torch.flip(pa, dims=[1]), dim=1), dims=[1])
        den = rest.to(pa.device)[:, None] + suffix
        out = pa / torch.where(valid, den, torch.ones_like(den)).clamp_min(1e-300)
        return torch.where(valid, out, torch.ones_like(out)).to(p.dtype)

    def _sample_gumbel(self, logits, bsz, dev, generator, sample0: int = 0, id_dtype=torch.int64):
        """Whole trajectories at once (exponential race / Gumbel-top-k, kernels K4g): equal in
        distribution to drawing one id per step from the re-normalised untaken mass (Plackett-Luce),
        O(A) per sample instead of O(A * T). The keys are a pure function of (seed, sample index, id)
        (Philox, generated inside the kernels): no B x A tensor, no library sort. Returns
        (actions [B, T] with the terminal last and -1 padding, taken int32 [B, words], length int32 [B])."""
        ctx = self.env.ctx
        seed = int(torch.randint(0, 2 ** 62, (1,), generator=generator,
                                 device=generator.device if generator is not None else "cpu"))
        self.last_seed = seed
        taken, length = ctx.sample_taken(logits, bsz, seed, sample0)
        actions = ctx.sample_order(logits, length, seed, sample0, dtype=id_dtype)
        return actions, taken, length

    def sample_states(self, s0, return_log=False, generator: torch.Generator | None = None,
                      method: str = "step"):
        """gflownet.py:125-197. Returns the Log (or None), as the reference does.

        method="step": one K4 masked-categorical kernel per environment step (the
        reference's loop shape); method="gumbel": whole trajectories at once."""
        bsz = len(s0)
        ctx = self.env.ctx
        dev = torch.device("cuda", ctx.device)
        log = Log(s0, self.backward_policy, self.total_flow, self.env) if return_log else None
        data = self.state_to_data(s0[:1])[0]
        p, alpha = self.forward_policy(data, torch.empty(0, dtype=torch.long))   # [1, A], with autograd
        p = p.reshape(-1)
        a = p.numel()
        logits = torch.log(p.detach().to(dev, torch.float32).clamp_min(1e-45)).contiguous()
        words = (a + 31) // 32
        taken = torch.zeros((bsz, words), dtype=torch.int32, device=dev)
        done = torch.zeros(bsz, dtype=torch.uint8, device=dev)
        if method == "gumbel":
            if not bool(torch.isfinite(logits).all()):
                raise RuntimeError("sample_states: the forward policy returned non-finite probabilities")
            complete_actions, taken, _ = self._sample_gumbel(logits, bsz, dev, generator)
            al = float(alpha.detach()) if isinstance(alpha, torch.Tensor) else float(alpha)
            rewards = self.env.update_from_taken(taken, al, max_deletions=complete_actions.shape[1])["reward"]
            if log is not None:
                log._actions = complete_actions.t().contiguous().cpu()
                log._fwd_probs = self.chosen_probs(p, complete_actions.to(p.device))
                log.rewards = rewards.to(torch.float32).cpu()
                log.alpha = alpha
            return log if return_log else None
        if method != "step":
            raise ValueError("method must be 'step' or 'gumbel'")
        if not bool(torch.isfinite(logits).all()):
            raise RuntimeError("sample_states: the forward policy returned non-finite probabilities")
        acts, probs = [], []
        step = 0
        while True:
            u = torch.rand(bsz, device=dev, generator=generator)
            act = torch.empty(bsz, dtype=torch.int64, device=dev)
            pr = torch.empty(bsz, dtype=torch.float32, device=dev)
            ctx.sample_step(logits, taken, u, done, act, pr)
            acts.append(act)
            probs.append(pr)
            step += 1
            if step % self.check_every == 0 or step >= a:
                if bool(done.all()):
                    break
                if step >= a + 1:       # every id incl. the terminal has been offered: cannot happen with finite logits
                    raise RuntimeError(f"sample_states: {int((done == 0).sum())} samples not finished after {step} steps")
        actions_tb = torch.stack(acts, dim=0)                     # [T', B]
        live = (actions_tb >= 0).any(dim=1)
        t_len = int(live.sum())                                   # drop all-finished trailing steps
        actions_tb = actions_tb[:t_len].contiguous()
        complete_actions = actions_tb.t().contiguous()            # [B, T] on the device
        al = float(alpha.detach()) if isinstance(alpha, torch.Tensor) else float(alpha)
        # the taken-bitmask IS the final state: score it directly (skips the actions -> mask kernel)
        rewards = self.env.update_from_taken(taken, al, max_deletions=t_len)["reward"]
        if log is not None:
            log._actions = actions_tb.cpu()
            log._fwd_probs = self.chosen_probs(p, complete_actions.to(p.device))
            log.sampled_probs = torch.stack(probs[:t_len], dim=0).t().cpu()   # K4's own fp32 values
            log.rewards = rewards.to(torch.float32).cpu()
            log.alpha = alpha
        return log if return_log else None
