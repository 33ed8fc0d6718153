"""K4g: whole trajectories by the exponential race (spai_sample_taken_dev / spai_sample_order_dev)
against the oracle's restatement (Philox4x32-10 known answers + float64 keys) through the C ABI."""
import numpy as np
import pytest
import torch

from gflownet_spai_b200 import synth
from oracle import spai_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from gflownet_spai_b200.env import SpaiContext
    p = synth.make_problem("cfg1")
    coo = p.a.tocoo()
    c = SpaiContext(p.n, p.edge_row, p.edge_col, p.edge_val, coo.row, coo.col, coo.data, device=0)
    yield c
    c.close()


def _unpack(taken, a):
    t = taken.cpu().numpy().view(np.uint32)
    bits = ((t[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(t.shape[0], -1)[:, :a]
    return bits.astype(bool)


def _check_against_keys(keys, taken, length, acts, a):
    """Everything the kernels emit must follow from the keys THEY computed: taken = {lq < 0} + terminal,
    order = ascending (lq, id), terminal last, -1 padding."""
    keys = keys.cpu().numpy()
    bits = _unpack(taken, a)
    ln = length.cpu().numpy()
    acts = acts.cpu().numpy()
    for b in range(keys.shape[0]):
        want = orc.race_trajectory(keys[b])
        assert ln[b] == want.size
        wbits = np.zeros(a, dtype=bool)
        wbits[want] = True
        assert np.array_equal(bits[b], wbits)
        assert np.array_equal(acts[b, :ln[b]], want)
        assert np.all(acts[b, ln[b]:] == -1)


@pytest.mark.parametrize("a,scale", [(2, 1.0), (5, 1.0), (461, 2.0), (4099, 0.5), (70001, 3.0)])
@pytest.mark.parametrize("id_dtype", [torch.int32, torch.int64])
def test_race_keys_match_oracle_and_outputs_follow_keys(ctx, a, scale, id_dtype):
    rng = np.random.default_rng(a)
    logits = (rng.normal(size=a) * scale).astype(np.float32)
    lg = torch.from_numpy(logits).cuda()
    bsz, seed, s0 = 12, 0x1234_5678_9ABC_DEF0 + a, 7
    taken, length, keys = ctx.sample_taken(lg, bsz, seed, s0, export_keys=True)
    acts = ctx.sample_order(lg, length, seed, s0, dtype=id_dtype)
    assert acts.dtype == id_dtype and acts.shape[1] == int(length.max())
    _check_against_keys(keys, taken, length, acts, a)
    # the keys themselves: Philox bits exact, fp32 transform vs float64 restatement
    k = keys.cpu().numpy()
    for b in (0, bsz - 1):
        want = orc.race_keys(logits, seed, s0 + b)
        np.testing.assert_allclose(k[b], want, rtol=2e-5, atol=2e-5)
    # sharding: rows [5, 9) drawn as their own call equal the same rows of the batch
    t2, l2 = ctx.sample_taken(lg, 4, seed, s0 + 5)
    assert torch.equal(t2, taken[5:9]) and torch.equal(l2, length[5:9])
    a2 = ctx.sample_order(lg, l2, seed, s0 + 5, dtype=id_dtype, ld=acts.shape[1])
    assert torch.equal(a2, acts[5:9])


@pytest.mark.parametrize("kind", ["peaked", "terminal_unlikely", "terminal_certain", "flat", "two_level"])
def test_race_orders_exactly_for_skewed_policies(ctx, kind):
    """Bucket occupancy depends on the policy; the order must not: heavy hitters, a terminal that
    almost never / almost always comes first, equal logits (ties broken by id), 60-nat spread."""
    a = 20011
    rng = np.random.default_rng(3)
    lg = np.zeros(a, dtype=np.float32)
    if kind == "peaked":
        lg[:] = rng.normal(size=a) * 0.1
        lg[rng.choice(a - 1, 50, replace=False)] += 12.0
    elif kind == "terminal_unlikely":
        lg[-1] = -60.0
    elif kind == "terminal_certain":
        lg[-1] = 30.0
    elif kind == "two_level":
        lg[: a // 2] = -40.0
        lg[-1] = -45.0
    t = torch.from_numpy(lg).cuda()
    taken, length, keys = ctx.sample_taken(t, 6, 99, 0, export_keys=True)
    acts = ctx.sample_order(t, length, 99, 0)
    _check_against_keys(keys, taken, length, acts, a)
    if kind == "terminal_unlikely":
        assert int(length.min()) == a            # every id is drawn before the terminal
    if kind == "terminal_certain":
        assert int(length.max()) <= 3


def test_race_distribution_matches_the_step_process(ctx):
    """Plackett-Luce equivalence, measured: first id ~ softmax(logits); second given the first
    ~ p_j / (1 - p_i); P(length = 1) = p_terminal."""
    a = 40
    rng = np.random.default_rng(5)
    logits = rng.normal(size=a).astype(np.float32)
    logits[-1] += 1.0
    p = np.exp(logits.astype(np.float64))
    p /= p.sum()
    bsz = 200000
    t = torch.from_numpy(logits).cuda()
    taken, length = ctx.sample_taken(t, bsz, 2024, 0)
    acts = ctx.sample_order(t, length, 2024, 0).cpu().numpy()
    first = acts[:, 0]
    counts = np.bincount(first, minlength=a).astype(np.float64)
    chi2 = float(((counts - bsz * p) ** 2 / (bsz * p)).sum())
    assert chi2 < 84.0                                            # chi2(39 dof) 99.997th percentile
    i = int(np.argmax(p[:-1]))
    sel = acts[first == i]
    c2 = np.bincount(sel[:, 1], minlength=a).astype(np.float64)
    e2 = sel.shape[0] * p / (1 - p[i])
    e2[i] = 0
    keep = e2 > 0
    assert c2[i] == 0
    assert float(((c2[keep] - e2[keep]) ** 2 / e2[keep]).sum()) < 82.0
    frac = float((length.cpu().numpy() == 1).mean())
    assert abs(frac - p[-1]) < 5 * np.sqrt(p[-1] * (1 - p[-1]) / bsz)


@pytest.mark.parametrize("a", [524281, 8372225])
def test_race_full_size_order_and_mask(ctx, a):
    """cfg2 / cfg4 action counts (K4 was only tested to A = 4099): the order the kernel
    writes equals a device-side stable sort of its own exported keys; mask and length agree."""
    g = torch.Generator(device="cuda").manual_seed(a)
    lg = torch.randn(a, generator=g, device="cuda") * 0.3
    bsz = 3
    taken, length, keys = ctx.sample_taken(lg, bsz, 77, 1000, export_keys=True)
    acts = ctx.sample_order(lg, length, 77, 1000)
    for b in range(bsz):
        k = keys[b, : a - 1]
        ids = torch.nonzero(k < 0).squeeze(1)
        order = ids[torch.sort(k[ids], stable=True).indices]       # ids ascending -> stable sort == (lq, id) order
        n = int(length[b])
        assert n == ids.numel() + 1
        assert torch.equal(acts[b, : n - 1].long(), order)
        assert int(acts[b, n - 1]) == a - 1 and bool((acts[b, n:] == -1).all())
    bits = torch.zeros((bsz, taken.shape[1] * 32), dtype=torch.bool, device="cuda")
    t64 = taken.long() & 0xFFFFFFFF
    for j in range(32):
        bits[:, j::32] = ((t64 >> j) & 1).bool()
    want = keys < 0
    want[:, a - 1] = True
    assert torch.equal(bits[:, :a], want)


def test_order_rejects_short_ld_and_foreign_lengths(ctx):
    lg = torch.zeros(1000, device="cuda")
    taken, length = ctx.sample_taken(lg, 8, 1, 0)
    with pytest.raises(ValueError):
        ctx.sample_order(lg, length, 1, 0, ld=max(1, int(length.max()) - 1))
    with pytest.raises(ValueError):
        ctx.sample_order(lg, length, 2, 0)          # lengths of another seed


# ---------------------------------------------------------------------------------------------
# K4p: many masked-categorical steps per launch (spai_sample_steps_dev)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("a,scale", [(5, 1.0), (461, 2.0), (1500, 6.0)])
def test_multi_step_kernel_follows_the_inverse_cdf_oracle_step_by_step(ctx, a, scale):
    """Whole trajectories with injected uniforms: every step's id must lie in the oracle's inverse-CDF
    interval for the ids taken so far (policy.py:64-73 masked softmax), its probability must match,
    the terminal ends the row, -1 / 1.0 afterwards (log.py:67-87) — the K4 test, for every step."""
    rng = np.random.default_rng(a)
    logits = (rng.normal(size=a) * scale).astype(np.float32)
    logits[-1] -= 2.0                                            # long trajectories
    bsz, nsteps = 6, a + 2
    u = rng.random((nsteps, bsz)).astype(np.float32)
    lg = torch.from_numpy(logits).cuda()
    words = (a + 31) // 32
    taken = torch.zeros((bsz, words), dtype=torch.int32, device="cuda")
    done = torch.zeros(bsz, dtype=torch.uint8, device="cuda")
    acts, probs, steps = ctx.sample_steps(lg, taken, done, nsteps, uniforms=torch.from_numpy(u).cuda(), dtype=torch.int64)
    acts, probs, steps = acts.cpu().numpy(), probs.cpu().numpy(), steps.cpu().numpy()
    assert bool(done.all())
    bits = _unpack(taken, a)
    for b in range(bsz):
        seen = []
        n = int(steps[b])
        assert acts[b, n - 1] == a - 1 and np.all(acts[b, n:] == -1) and np.all(probs[b, n:] == 1.0)
        for s in range(n):
            p = orc.masked_softmax_probs(logits, seen).astype(np.float64)
            cdf = np.cumsum(p)
            x = int(acts[b, s])
            assert 0 <= x < a and x not in seen
            lo = cdf[x - 1] if x else 0.0
            tgt = float(u[s, b]) * cdf[-1]
            assert lo - 2e-5 <= tgt <= cdf[x] + 2e-5
            assert probs[b, s] == pytest.approx(p[x] / cdf[-1], rel=5e-4, abs=1e-7)
            seen.append(x)
        want = np.zeros(a, dtype=bool)
        want[seen] = True
        assert np.array_equal(bits[b], want)


def test_multi_step_kernel_resumes_and_matches_single_launch(ctx):
    """nsteps split over several launches (step0 advancing) draws the same trajectory as one launch
    (Philox uniforms are a function of (seed, sample, step)); finished samples are left alone."""
    a = 3001
    lg = torch.randn(a, generator=torch.Generator(device="cuda").manual_seed(1), device="cuda")
    bsz, total = 9, a + 1
    words = (a + 31) // 32
    t1 = torch.zeros((bsz, words), dtype=torch.int32, device="cuda")
    d1 = torch.zeros(bsz, dtype=torch.uint8, device="cuda")
    a1, p1, s1 = ctx.sample_steps(lg, t1, d1, total, seed=42, sample0=100)
    t2 = torch.zeros_like(t1)
    d2 = torch.zeros_like(d1)
    a2 = torch.empty((bsz, total), dtype=torch.int32, device="cuda")
    p2 = torch.empty((bsz, total), dtype=torch.float32, device="cuda")
    step0 = 0
    for chunk in (1, 7, 500, total - 508):
        ctx.sample_steps(lg, t2, d2, chunk, seed=42, sample0=100, step0=step0, actions=a2, probs=p2)
        step0 += chunk
    assert bool(d1.all()) and bool(d2.all())
    assert torch.equal(a1, a2) and torch.equal(t1, t2)
    assert torch.allclose(p1, p2, rtol=1e-5, atol=1e-9)
    # every row: distinct ids, terminal last
    for b in range(bsz):
        n = int(s1[b])
        row = a1[b, :n].tolist()
        assert len(set(row)) == n and row[-1] == a - 1


def test_multi_step_kernel_at_cfg2_size_probabilities(ctx):
    """A = 524 281 (cfg2): 3000 steps of 8 samples; ids distinct, and the reported probability equals
    p[a_t] / (mass not yet taken) computed in float64 on the device."""
    a = 524281
    lg = torch.randn(a, generator=torch.Generator(device="cuda").manual_seed(2), device="cuda") * 0.5
    lg[-1] = -30.0                                               # the terminal practically never comes up
    bsz, nsteps = 8, 3000
    words = (a + 31) // 32
    taken = torch.zeros((bsz, words), dtype=torch.int32, device="cuda")
    done = torch.zeros(bsz, dtype=torch.uint8, device="cuda")
    acts, probs, steps = ctx.sample_steps(lg, taken, done, nsteps, seed=5)
    assert int(steps.min()) == nsteps and not bool(done.any())
    p = torch.softmax(lg.double(), 0)
    pa = p[acts.long()]
    rest = 1.0 - (torch.cumsum(pa, 1) - pa)
    assert torch.allclose(probs.double(), pa / rest, rtol=2e-4)
    srt = torch.sort(acts, dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())
    # first ids ~ softmax: the mean log-probability of the first draw is close to its expectation
    many_t = torch.zeros((4096, words), dtype=torch.int32, device="cuda")
    many_d = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    a1, _, _ = ctx.sample_steps(lg, many_t, many_d, 1, seed=6, want_probs=False)
    got = float(torch.log(p[a1[:, 0].long()]).mean())
    want = float((p * torch.log(p)).sum())
    sd = float(((p * torch.log(p) ** 2).sum() - want ** 2).sqrt()) / 64.0
    assert abs(got - want) < 5 * sd
