"""DRQN training-mode host logic on the CPU: sequence replay over a lock-step ring against a restatement of the
reference's SequenceReplayBuffer (scripts/train_rnn_iterative.py:100-171), and the last-step Double-DQN loss of
train_step_rnn (:400-531) against a step-by-step unrolled computation."""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import pingpong_selfplay_ai_b200 as pp
from pingpong_selfplay_ai_b200.train_rnn import SequenceSampler
from oracle.train_port import TorchDRQNTrainer

L = 8


def _lockstep_ring(n, T, steps, seed=0, p_done=0.08):
    """A ring written the way the rollout kernel does (slot = (t % T) * n + i): obs[.., 0] = absolute step, obs[.., 1] =
    env, so every sampled row can be traced back.  Returns (ring, done[steps, n])."""
    rs = np.random.RandomState(seed)
    done = rs.rand(steps, n) < p_done
    ring = pp.ReplayRing(n * T, device="cpu", lockstep_envs=n)
    for t in range(steps):
        for i in range(n):
            slot = (t % T) * n + i
            ring.obs[slot, 0], ring.obs[slot, 1] = float(t), float(i)
            ring.next_obs[slot, 0], ring.next_obs[slot, 1] = float(t) + 0.5, float(i)
            ring.act[slot] = (t + i) % 3
            ring.rew[slot] = float(done[t, i])
            ring.done[slot] = int(done[t, i])
    ring.steps_written = steps
    return ring, done


def _reference_buffer(done, n, T, steps, fresh=True):
    """SequenceReplayBuffer semantics per env over the steps still in the ring: episodes closed by done, kept iff
    len >= trace_length (:115-118); an episode whose first rows were overwritten is dropped.
    -> {(t_end, env): weight}, number of stored episodes."""
    lo = max(0, steps - T)
    weights, episodes = {}, 0
    for i in range(n):
        cur, known_start = [], (lo == 0 and fresh)
        for t in range(lo, steps):
            cur.append(t)
            if done[t, i]:
                if known_start and len(cur) >= L:
                    episodes += 1
                    for s in range(len(cur) - L + 1):                      # uniform start inside the episode (:139)
                        weights[(cur[s + L - 1], i)] = 1.0 / (len(cur) - L + 1)
                cur, known_start = [], True
    return weights, episodes


@pytest.mark.parametrize("steps", [25, 40, 41, 97])
def test_window_weights_equal_the_reference_episode_then_window_distribution(steps):
    n, T = 6, 40
    ring, done = _lockstep_ring(n, T, steps, seed=steps)
    s = SequenceSampler(ring, trace_length=L)
    assert s.refresh() == _reference_buffer(done, n, T, steps)[1] == len(s)
    w, d, shift = s.window_weights()
    want, _ = _reference_buffer(done, n, T, steps)
    lo = max(0, steps - T)
    got = {(lo + int(t), int(i)): float(w[t, i]) for t, i in zip(*np.nonzero(w.numpy()))}
    assert got.keys() == want.keys()
    assert all(abs(got[k] - want[k]) < 1e-12 for k in want)
    assert shift == (steps % T if steps > T else 0)


def test_sampled_windows_are_contiguous_and_inside_one_episode():
    n, T, steps = 7, 128, 300
    ring, done = _lockstep_ring(n, T, steps, seed=5, p_done=0.06)
    s = SequenceSampler(ring, trace_length=L)
    assert s.refresh() > 20
    g = torch.Generator().manual_seed(1)
    obs, act, rew, nxt, dn = s.sample(6000, g)
    assert obs.shape == (6000, L, 7) and act.dtype == torch.int64 and dn.dtype == torch.bool
    t, env = obs[:, :, 0].numpy().astype(int), obs[:, :, 1].numpy().astype(int)
    assert np.all(np.diff(t, axis=1) == 1) and np.all(env == env[:, :1])              # consecutive steps of one env
    assert t.min() >= steps - T and t.max() < steps                                    # only rows still in the ring
    assert not dn[:, :-1].any()                                                        # no episode boundary inside
    assert np.array_equal(dn.numpy(), done[t, env]) and np.array_equal(rew.numpy(), done[t, env].astype(np.float32))
    assert np.all(nxt[:, :, 0].numpy() == t + 0.5) and np.array_equal(act.numpy(), (t + env) % 3)
    # every stored episode is drawn about equally often, whatever its length (:131)
    want, episodes = _reference_buffer(done, n, T, steps)
    ends = {}
    for (te, i), wgt in want.items():                                                  # episode id = its last step
        e = te
        while not done[e, i]:
            e += 1
        ends[(te, i)] = (e, i)
    counts = {}
    for te, i in zip(t[:, -1], env[:, -1]):
        counts[ends[(te, i)]] = counts.get(ends[(te, i)], 0) + 1
    assert len(counts) == episodes
    assert max(counts.values()) < 2.0 * 6000 / episodes and min(counts.values()) > 0.4 * 6000 / episodes


def test_sampler_rejects_an_append_ring_and_an_empty_one():
    with pytest.raises(ValueError):
        SequenceSampler(pp.ReplayRing(64, device="cpu"))
    with pytest.raises(ValueError):
        pp.ReplayRing(65, device="cpu", lockstep_envs=8)
    s = SequenceSampler(pp.ReplayRing(64, device="cpu", lockstep_envs=8))
    assert s.refresh() == 0
    with pytest.raises(RuntimeError):
        s.sample(4)


def _unrolled_last_q(net, seq):
    """Q at the last step of each window with zero initial (h, c), one step at a time."""
    hc = net.init_hidden(seq.shape[0], seq.device)
    for t in range(seq.shape[1]):
        q, hc = net(seq[:, t:t + 1, :], hc)
    return q


def test_drqn_loss_and_update_follow_train_step_rnn():
    torch.manual_seed(4)
    net = pp.QNetRNN()
    tr = TorchDRQNTrainer(copy.deepcopy(net), gamma=0.99, lr=1e-3, batch_size=16, target_update_interval=2, device="cpu",
                     use_graph=False)
    assert sum(p.numel() for p in tr.params) == sum(p.numel() for p in net.parameters()) and all(p.requires_grad for p in tr.params)
    with torch.no_grad():                                            # make online and target differ
        for p in tr.target.parameters():
            p.add_(0.05 * torch.randn_like(p))
    b = 16
    obs, nxt = torch.rand(b, L, 7) * 2 - 1, torch.rand(b, L, 7) * 2 - 1
    act = torch.randint(0, 3, (b, L))
    rew = torch.randint(-1, 2, (b, L)).float()
    done = torch.rand(b, L) < 0.3
    loss = tr.loss_on(obs, act, rew, nxt, done)
    with torch.no_grad():                                            # train_step_rnn restated with unrolled steps
        q = _unrolled_last_q(tr.model, obs).gather(1, act[:, -1:]).squeeze(1)
        best = _unrolled_last_q(tr.model, nxt).argmax(1, keepdim=True)
        nq = _unrolled_last_q(tr.target, nxt).gather(1, best).squeeze(1)
        want = F.smooth_l1_loss(q, rew[:, -1] + 0.99 * nq * (~done[:, -1]))
    assert abs(loss.item() - want.item()) < 1e-6
    # update(): None until batch_size episodes are stored, then Adam on clipped gradients and target sync
    n, T = 8, 64
    ring, _ = _lockstep_ring(n, T, 20, seed=2, p_done=0.0)
    s = SequenceSampler(ring, trace_length=L)
    s.refresh()
    assert tr.update(s) is None and tr.train_steps == 0
    ring, _ = _lockstep_ring(n, T, 64, seed=3, p_done=0.09)
    ring.obs.uniform_(-1, 1); ring.next_obs.uniform_(-1, 1)
    s = SequenceSampler(ring, trace_length=L)
    assert s.refresh() >= 16
    before = [p.detach().clone() for p in tr.params]
    g = torch.Generator().manual_seed(0)
    l1 = tr.update(s, generator=g)
    assert l1 is not None and torch.isfinite(l1) and tr.train_steps == 1
    assert all(not torch.equal(a, p) for a, p in zip(before, tr.params))               # every parameter trains (:335)
    assert any(not torch.equal(p, q) for p, q in zip(tr.model.parameters(), tr.target.parameters()))
    tr.update(s, generator=g)                                                          # step 2 = target_update_interval
    assert all(torch.equal(p, q) for p, q in zip(tr.model.parameters(), tr.target.parameters()))
    # gradient clipping: a huge loss scale still moves each parameter by at most ~lr per step (Adam) and the stored
    # gradients have norm <= 1
    big = TorchDRQNTrainer(copy.deepcopy(net), lr=1e-3, batch_size=16, device="cpu", use_graph=False, grad_clip_norm=1.0)
    ring.rew.mul_(1e4)
    big.update(s, generator=g)
    total = torch.sqrt(sum((p.grad ** 2).sum() for p in big.params))
    assert total <= 1.0 + 1e-4
