"""Host-side logic of the sampler mirror (no GPU): loss, Log bookkeeping and the
differentiable chosen-probability reconstruction."""
import numpy as np
import pytest
import torch

from gflownet_spai_b200.sampler import GFlowNet, Log, trajectory_balance_loss
from oracle import ref_shim
from oracle import spai_oracle as orc


def _rand_inputs(seed=0, b=5, t=7):
    g = torch.Generator().manual_seed(seed)
    fwd = torch.rand(b, t, generator=g).clamp_min(1e-3).requires_grad_(True)
    back = torch.rand(b, t, generator=g).clamp_min(1e-3)
    rewards = torch.rand(b, generator=g) * 500 + 1
    return torch.ones(1), rewards, fwd, back


def test_tb_loss_formula():
    z, r, f, bk = _rand_inputs()
    got = trajectory_balance_loss(z, r, f, bk)
    eps = 1e-9
    lf = torch.log(f + eps).sum(-1)
    lb = torch.log(bk + eps).sum(-1)
    want = ((torch.log(z + eps) + lf - lf.max() - torch.log(r + eps) - (lb - lb.max())) ** 2).mean()
    assert float(got) == pytest.approx(float(want), rel=1e-6)
    got.backward()
    assert torch.isfinite(f.grad).all()


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference checkout absent")
def test_tb_loss_and_log_match_live_reference():
    ref = ref_shim.load_reference()
    for seed in range(3):
        z, r, f, bk = _rand_inputs(seed)
        assert float(trajectory_balance_loss(z, r, f, bk)) == pytest.approx(
            float(ref.trajectory_balance_loss(z, r, f.detach(), bk)), rel=1e-6)
    # Log.log bookkeeping: chosen prob / -1 padding for finished rows (log.py:24-89)
    s0 = [None] * 4
    mine, theirs = Log(s0, None, torch.ones(1), None), ref.Log(s0, None, torch.ones(1), None)
    g = torch.Generator().manual_seed(1)
    for step in range(3):
        probs = torch.softmax(torch.randn(4, 1, 6, generator=g), dim=-1)
        acts = torch.randint(0, 6, (4, 1), generator=g)
        done = torch.tensor([False, step > 0, False, step > 1])
        mine.log(s0, probs, acts, done)
        theirs.log(s0, probs, acts, done)
    assert torch.equal(mine.actions, theirs.actions)
    assert torch.allclose(mine.fwd_probs, theirs.fwd_probs)


def test_chosen_probs_equal_sequential_masked_softmax():
    rng = np.random.default_rng(3)
    a = 12
    logits = rng.normal(size=a).astype(np.float32)
    p = torch.softmax(torch.tensor(logits, dtype=torch.float64), dim=0).requires_grad_(True)
    acts = torch.tensor([[3, 7, 1, 11, -1], [11, -1, -1, -1, -1], [0, 1, 2, 3, 11]])
    got = GFlowNet.chosen_probs(p, acts)
    for b in range(acts.shape[0]):
        taken = []
        for t in range(acts.shape[1]):
            x = int(acts[b, t])
            if x < 0:
                assert float(got[b, t]) == 1.0
                continue
            want = orc.masked_softmax_probs(logits, taken)[x]
            assert float(got[b, t]) == pytest.approx(float(want), rel=2e-6)
            taken.append(x)
    got.log().sum().backward()
    assert torch.isfinite(p.grad).all()


def _loop_backward_policy(lstm, fc, traj):
    """Per-trajectory restatement of policy.py:87-129 (used when the reference checkout is absent)."""
    outs = []
    max_len = traj.shape[1]
    for i in range(traj.shape[0]):
        n_valid = int((traj[i] != -1).sum())
        seq = traj[i, :n_valid].float().view(1, n_valid, 1)
        h, _ = lstm(seq)
        o = torch.softmax(fc(h[:, -1, :])[:, :n_valid], dim=1)
        outs.append(torch.nn.functional.pad(o, (0, max_len - o.size(1)), value=1.0))
    return torch.stack(outs, dim=0)


def _trajectories_for_backward(seed, b=9, t=11, a=30):
    g = torch.Generator().manual_seed(seed)
    traj = torch.full((b, t), -1, dtype=torch.long)
    for i in range(b):
        n = int(torch.randint(1, t + 1, (1,), generator=g))
        traj[i, :n] = torch.randint(0, a, (n,), generator=g)
    traj[0, :] = torch.randint(0, a, (t,), generator=g)          # one full-length row
    return traj


def test_batched_backward_policy_equals_the_per_trajectory_loop():
    from gflownet_spai_b200.sampler import BackwardPolicy
    torch.manual_seed(0)
    for max_actions in (40, 6):                                   # 6 < T: fc narrower than the longest trajectory
        pol = BackwardPolicy(1, 16, max_actions)
        traj = _trajectories_for_backward(3)
        got = pol(traj)
        want = _loop_backward_policy(pol.lstm, pol.fc, traj)
        assert got.shape == want.shape == (traj.shape[0], 1, traj.shape[1])
        assert torch.allclose(got, want, atol=1e-6)
        got.sum().backward()                                      # differentiable end to end
        assert all(p.grad is not None for p in pol.parameters())
        pol.zero_grad()
    with pytest.raises(RuntimeError):
        pol(torch.full((2, 3), -1, dtype=torch.long))


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference checkout absent")
def test_batched_backward_policy_matches_live_reference():
    from gflownet_spai_b200.sampler import BackwardPolicy
    ref = ref_shim.load_reference()
    torch.manual_seed(1)
    theirs = ref.BackwardPolicy(1, 12, 25)
    mine = BackwardPolicy(1, 12, 25)
    mine.load_state_dict(theirs.state_dict())                    # same parameter names
    for seed in range(3):
        traj = _trajectories_for_backward(seed, b=7, t=9, a=25)
        with torch.no_grad():
            assert torch.allclose(mine(traj), theirs(traj), atol=1e-6)
    # through Log.back_probs (log.py:124-164): [B, T]
    log = Log([None] * 7, mine, torch.ones(1), None)
    log._actions = traj.t().contiguous()
    assert log.back_probs.shape == (7, 9)
