"""K4 (masked categorical step) and the device sampling loop."""
import numpy as np
import pytest
import torch

from gflownet_spai_b200 import synth
from oracle import spai_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small_env():
    from gflownet_spai_b200.env import PreconditionerEnv
    p = synth.make_problem("cfg1")
    init = torch.sparse_coo_tensor(torch.tensor(np.stack([p.edge_row, p.edge_col])),
                                   torch.tensor(p.edge_val, dtype=torch.float32), (p.n, p.n))
    env = PreconditionerEnv(p.n, init, init.clone())
    yield env, init
    env.ctx.close()


def _pack(taken_lists, a):
    words = (a + 31) // 32
    out = np.zeros((len(taken_lists), words), dtype=np.uint32)
    for b, ids in enumerate(taken_lists):
        for i in ids:
            out[b, i // 32] |= np.uint32(1) << np.uint32(i % 32)
    return torch.from_numpy(out.view(np.int32)).cuda()


@pytest.mark.parametrize("a", [5, 461, 4099])
def test_sample_step_matches_inverse_cdf_oracle(small_env, a):
    env, _ = small_env
    rng = np.random.default_rng(a)
    bsz = 96
    logits = rng.normal(scale=2.0, size=a).astype(np.float32)
    taken_lists = [list(rng.choice(a - 1, size=int(rng.integers(0, a - 1)), replace=False)) for _ in range(bsz)]
    done0 = rng.random(bsz) < 0.2
    u = rng.random(bsz).astype(np.float32)
    taken = _pack(taken_lists, a)
    before = taken.clone()
    done = torch.from_numpy(done0.astype(np.uint8)).cuda()
    act = torch.empty(bsz, dtype=torch.int64, device="cuda")
    prob = torch.empty(bsz, dtype=torch.float32, device="cuda")
    env.ctx.sample_step(torch.from_numpy(logits).cuda(), taken, torch.from_numpy(u).cuda(), done, act, prob)
    act, prob, done = act.cpu().numpy(), prob.cpu().numpy(), done.cpu().numpy().astype(bool)
    for b in range(bsz):
        if done0[b]:
            assert act[b] == -1 and prob[b] == 1.0 and done[b]
            assert torch.equal(taken[b], before[b])
            continue
        p = orc.masked_softmax_probs(logits, taken_lists[b]).astype(np.float64)
        cdf = np.cumsum(p)
        x = int(act[b])
        assert 0 <= x < a and x not in taken_lists[b]            # never re-select a taken id
        lo = cdf[x - 1] if x else 0.0
        tgt = float(u[b]) * cdf[-1]
        assert lo - 1e-5 <= tgt <= cdf[x] + 1e-5                  # inverse-CDF interval (fp32 sums)
        assert prob[b] == pytest.approx(p[x] / cdf[-1], rel=2e-4, abs=1e-7)
        assert done[b] == (x == a - 1)                            # terminal id = A-1
        assert (int(taken[b, x // 32]) >> (x % 32)) & 1          # state updated in place


def test_sample_step_distribution_chi_square(small_env):
    env, _ = small_env
    a, bsz = 9, 40000
    logits = torch.log(torch.tensor([0.05, 0.2, 0.05, 0.1, 0.1, 0.15, 0.05, 0.2, 0.1]))
    taken = _pack([[2, 6]] * bsz, a)
    done = torch.zeros(bsz, dtype=torch.uint8, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(7)
    u = torch.rand(bsz, device="cuda", generator=g)
    act = torch.empty(bsz, dtype=torch.int64, device="cuda")
    prob = torch.empty(bsz, dtype=torch.float32, device="cuda")
    env.ctx.sample_step(logits.cuda(), taken, u, done, act, prob)
    counts = np.bincount(act.cpu().numpy(), minlength=a).astype(np.float64)
    p = orc.masked_softmax_probs(logits.numpy(), [2, 6]).astype(np.float64)
    assert counts[2] == 0 and counts[6] == 0
    keep = p > 0
    chi2 = float(((counts[keep] - bsz * p[keep]) ** 2 / (bsz * p[keep])).sum())
    assert chi2 < 27.9                                            # chi2(6 dof) 99.99th percentile


class _StubForward(torch.nn.Module):
    """Stand-in for policy.ForwardPolicy (GATv2Conv is unavailable): same
    call signature and return convention (probs [1, A], sigmoid(alpha))."""

    def __init__(self, a):
        super().__init__()
        self.logit = torch.nn.Parameter(torch.linspace(-1.0, 1.0, a))
        self.alpha = torch.nn.Parameter(torch.tensor(0.0))

    def forward(self, data, actions):
        x = self.logit[None, : data.edge_attr.size(0) + 1]
        if actions.numel() > 0:
            m = torch.ones_like(x, dtype=torch.bool)
            m[:, actions] = 0
            x = x.masked_fill(~m, float("-inf"))
        return torch.softmax(x, dim=1), torch.sigmoid(self.alpha)


class _StubBackward(torch.nn.Module):
    def forward(self, trajectories):
        return torch.full(trajectories.shape, 0.5)


def test_sample_states_device_loop(small_env):
    from gflownet_spai_b200.sampler import GFlowNet, trajectory_balance_loss
    env, init = small_env
    a = env.num_actions
    fp = _StubForward(a)
    with torch.no_grad():
        fp.logit[-1] = 3.0                       # make termination likely so trajectories stay short
    model = GFlowNet(fp, _StubBackward(), env)
    bsz = 16
    s0 = [init.clone() for _ in range(bsz)]
    log = model.sample_states(s0, return_log=True, generator=torch.Generator(device="cuda").manual_seed(3))
    acts = log.actions                           # [T, B]
    assert acts.shape[1] == bsz and log.fwd_probs.shape == (bsz, acts.shape[0])
    assert len(log._actions) == acts.shape[0]
    for b in range(bsz):
        col = acts[:, b].tolist()
        valid = [x for x in col if x != -1]
        assert valid[-1] == a - 1 and len(set(valid)) == len(valid)        # ends at terminal, no repeats
        assert all(x == -1 for x in col[len(valid):])                      # -1 only after the terminal
    # rewards equal a fresh evaluation of the logged actions (gflownet.py:181-195)
    again = torch.tensor(env.update(s0, acts.t(), 0.5), dtype=torch.float32)
    assert torch.allclose(log.rewards, again, rtol=1e-5, atol=1e-3)
    # differentiable chosen probabilities agree with K4's own fp32 values
    assert torch.allclose(log.fwd_probs.detach().float(), log.sampled_probs, rtol=2e-3, atol=1e-6)
    loss = trajectory_balance_loss(log.total_flow, log.rewards.clamp_min(1e-3), log.fwd_probs, log.back_probs)
    loss.backward()
    assert torch.isfinite(fp.logit.grad).all() and float(fp.logit.grad.abs().sum()) > 0
    assert model.sample_states(s0, return_log=False) is None


def test_pack_taken_kernel_matches_threshold_rule(small_env):
    env, _ = small_env
    g = torch.Generator(device="cuda").manual_seed(5)
    for a in (5, 461, 4100):
        keys = torch.randn((7, a), device="cuda", generator=g)
        taken, length = env.ctx.pack_taken(keys)
        want = keys > keys[:, a - 1:a]
        want[:, a - 1] = True
        got = torch.zeros_like(want)
        tw = taken.cpu().numpy().view(np.uint32)
        for b in range(7):
            bits = np.unpackbits(tw[b].view(np.uint8), bitorder="little")[:a]
            got[b] = torch.from_numpy(bits.astype(bool)).cuda()
        assert torch.equal(got, want)
        assert torch.equal(length.to(torch.int64), want.sum(1))


def test_gumbel_whole_trajectory_sampler(small_env):
    """method='gumbel': same Log contract and, in distribution, the same
    trajectories as the step sampler (Plackett-Luce equivalence)."""
    from gflownet_spai_b200.sampler import GFlowNet, trajectory_balance_loss
    env, init = small_env
    a = env.num_actions
    fp = _StubForward(a)
    with torch.no_grad():
        fp.logit[:] = -25.0                       # mass on ids 0..5 and the terminal id
        fp.logit[:6] = torch.tensor([0.0, 0.5, 1.0, -0.5, 0.2, 0.8])
        fp.logit[-1] = 1.2
    model = GFlowNet(fp, _StubBackward(), env)
    bsz = 24000
    s0 = [init] * bsz
    gen = torch.Generator(device="cuda").manual_seed(11)
    log = model.sample_states(s0, return_log=True, generator=gen, method="gumbel")
    acts = log.actions.t()                        # [B, T]
    assert acts.shape[0] == bsz and log.fwd_probs.shape == acts.shape
    lens = (acts >= 0).sum(1)
    last = acts.gather(1, (lens - 1)[:, None]).squeeze(1)
    assert torch.all(last == a - 1)                                           # every trajectory ends at terminal
    for b in range(0, bsz, 997):
        row = acts[b, : int(lens[b])].tolist()
        assert len(set(row)) == len(row) and all(x == -1 for x in acts[b, int(lens[b]):].tolist())
    # first action ~ softmax(logits) (7 categories that carry the mass)
    p = torch.softmax(fp.logit.detach(), 0).double().numpy()
    cats = [0, 1, 2, 3, 4, 5, a - 1]
    first = acts[:, 0].numpy()
    counts = np.array([(first == c).sum() for c in cats], dtype=np.float64)
    expect = bsz * p[cats]
    chi2 = float(((counts - expect) ** 2 / expect).sum())
    assert counts.sum() >= bsz - 5 and chi2 < 27.9                            # chi2(6 dof) 99.99th percentile
    # second action given the first: p_j / (1 - p_i) for a fixed first id
    sel = first == 2
    second = acts[sel, 1].numpy()
    cats2 = [0, 1, 3, 4, 5, a - 1]
    c2 = np.array([(second == c).sum() for c in cats2], dtype=np.float64)
    e2 = sel.sum() * p[cats2] / (1 - p[2])
    assert float(((c2 - e2) ** 2 / e2).sum()) < 25.7                          # chi2(5 dof) 99.99th percentile
    # rewards equal a fresh evaluation of the logged actions; loss is differentiable
    again = env.update_tensor(acts[:64].contiguous(), 0.5, want=("reward",))["reward"].float()
    assert torch.allclose(log.rewards[:64], again, rtol=1e-5, atol=1e-3)
    loss = trajectory_balance_loss(log.total_flow, log.rewards.clamp_min(1e-3), log.fwd_probs, log.back_probs)
    loss.backward()
    assert torch.isfinite(fp.logit.grad).all()


def test_training_step_with_batched_backward_policy_and_ls_gram_rewards():
    """The loop of the reference's train.py / GFlowNet100.py:285-310 on the drop-in classes:
    sample a batch on the device, least-squares rewards, batched LSTM backward policy,
    trajectory-balance loss, one optimiser step on both policies."""
    from gflownet_spai_b200.env import PreconditionerEnv
    from gflownet_spai_b200.sampler import BackwardPolicy, GFlowNet, trajectory_balance_loss
    p = synth.make_problem("cfg1")
    init = torch.sparse_coo_tensor(torch.tensor(np.stack([p.edge_row, p.edge_col])),
                                   torch.tensor(p.edge_val, dtype=torch.float32), (p.n, p.n))
    env = PreconditionerEnv(p.n, init, init.clone(), mode="ls_gram", dtype=torch.float64)
    fp = _StubForward(env.num_actions)
    with torch.no_grad():
        fp.logit[-1] = 4.0
    bp = BackwardPolicy(1, 8, env.num_actions)
    model = GFlowNet(fp, bp, env)
    opt = torch.optim.Adam(list(fp.parameters()) + list(bp.parameters()), lr=1e-2)
    s0 = [init.clone() for _ in range(64)]
    losses = []
    for it in range(2):
        log = model.sample_states(s0, return_log=True, method="gumbel",
                                  generator=torch.Generator(device="cuda").manual_seed(10 + it))
        assert log.back_probs.shape == log.fwd_probs.shape
        loss = trajectory_balance_loss(log.total_flow, log.rewards.clamp_min(1e-3), log.fwd_probs, log.back_probs)
        opt.zero_grad()
        loss.backward()
        assert all(torch.isfinite(q.grad).all() for q in list(fp.parameters()) + list(bp.parameters()) if q.grad is not None)
        opt.step()
        losses.append(float(loss.detach()))
    assert all(np.isfinite(losses))
    # ls rewards never below copy rewards for the same trajectories (smaller residual)
    acts = log.actions.t().contiguous()
    r_ls = torch.tensor(env.update(s0, acts, 0.5), dtype=torch.float64)
    env_copy = PreconditionerEnv(p.n, init, init.clone(), mode="copy", dtype=torch.float64)
    r_cp = torch.tensor(env_copy.update(s0, acts, 0.5), dtype=torch.float64)
    assert torch.all(r_ls >= r_cp - 1e-6)
    env.ctx.close()
    env_copy.ctx.close()
