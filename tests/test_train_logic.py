"""Training-mode host logic on the CPU: prioritised-replay bookkeeping and the PyTorch formulation of the Double-DQN head
update (oracle/train_port.TorchDQNTrainer = the product trainer's host logic with autograd in place of the CUDA kernels),
checked against a line-by-line restatement of the reference's train_step (scripts/train_iterative.py:132-168) on the same
batch.  The hand-written kernels are then checked against this formulation on the GPU (tests/test_gpu_train_kernels.py)."""
import copy

import pytest
import torch

import pingpong_selfplay_ai_b200 as pp
from oracle.train_port import TorchDQNTrainer, per_sample_torch


def _filled_ring(cap, rows, seed=0):
    g = torch.Generator().manual_seed(seed)
    ring = pp.ReplayRing(cap, device="cpu")
    k = min(rows, cap)
    ring.obs[:k] = torch.rand(k, 7, generator=g)
    ring.next_obs[:k] = torch.rand(k, 7, generator=g)
    ring.act[:k] = torch.randint(0, 3, (k,), generator=g).to(torch.uint8)
    ring.rew[:k] = torch.randint(-1, 2, (k,), generator=g).float()
    ring.done[:k] = (torch.rand(k, generator=g) < 0.1).to(torch.uint8)
    ring.head.fill_(rows)
    return ring


def test_prioritized_sampler_new_rows_wrap_and_probabilities():
    ring = _filled_ring(100, 0)
    s = pp.PrioritizedSampler(ring, alpha=0.6)
    ring.head.fill_(30)
    assert s.note_new_rows() == 30 and len(s) == 30
    assert torch.all(s.prios[:30] == 1.0) and torch.all(s.prios[30:] == 0)            # first rows: priority 1.0 (:57)
    s.update_priorities(torch.tensor([3, 4]), torch.tensor([5.0, -0.5]))
    assert s.prios[3] == pytest.approx(5.0 + 1e-6) and s.prios[4] == pytest.approx(0.5 + 1e-6)
    ring.head.fill_(120)                                                               # 90 new rows, wrapping past 100
    assert s.note_new_rows() == 90 and len(s) == 100
    assert torch.all(s.prios[30:] == s.prios[3]) and torch.all(s.prios[:20] == s.prios[3])   # max priority (:57,62)
    assert s.prios[25] == 1.0                                                          # untouched old row
    torch.manual_seed(0)
    s.prios.fill_(1.0); s.update_priorities(torch.tensor([7]), torch.tensor([1000.0]))
    idx, w = per_sample_torch(s, 4000, 0.5)
    p7 = 1000.0 ** 0.6 / (99 + 1000.0 ** 0.6)
    assert abs((idx == 7).float().mean().item() - p7) < 0.03
    assert w.max() == 1.0 and w[idx == 7][0] == pytest.approx((100 * p7) ** -0.5 / (100 * (1 - p7) / 99) ** -0.5, rel=1e-4)
    ring.head.fill_(500)                                                               # more new rows than capacity
    s.note_new_rows()
    assert torch.all(s.prios == s.max_prio) and float(s.max_prio) == pytest.approx(1000.0)     # the running maximum (:57)


def test_dqn_update_equals_reference_train_step_restated():
    torch.manual_seed(3)
    net = pp.QNet()
    ring = _filled_ring(4096, 3000, seed=1)
    s1 = pp.PrioritizedSampler(ring); s1.note_new_rows()
    s1.prios[:3000] = torch.rand(3000) + 0.01
    tr = TorchDQNTrainer(copy.deepcopy(net), gamma=0.99, lr=2.5e-4, batch_size=256, target_update_interval=2, device="cpu")
    ref_model, ref_target = copy.deepcopy(tr.model), copy.deepcopy(tr.target)
    ref_prios = s1.prios.clone()
    heads = list(ref_model.fc_V.parameters()) + list(ref_model.fc_A.parameters())
    assert sum(p.numel() for p in heads) == 520 and not any(p.requires_grad for p in ref_model.features.parameters())
    ref_opt = torch.optim.Adam(heads, lr=2.5e-4)
    for step in range(1, 4):
        torch.manual_seed(100 + step)
        loss = tr.update(s1)
        # ---- the reference's train_step on the same RNG stream (scripts/train_iterative.py:136-168)
        torch.manual_seed(100 + step)
        beta = min(1.0, 0.4 + step * (1.0 - 0.4) / 100000)
        probs = ref_prios[:3000] ** 0.6; probs = probs / probs.sum()
        cdf = probs.cumsum(0)                                  # np.random.choice(p=probs): inverse CDF
        idxs = torch.searchsorted(cdf, torch.rand(256) * cdf[-1], right=True).clamp(max=2999)
        iw = (3000 * probs[idxs]) ** (-beta); iw = iw / iw.max()
        ref_model.reset_noise(); ref_target.reset_noise()
        st, a, r = ring.obs[idxs], ring.act[idxs].long(), ring.rew[idxs]
        ns, d = ring.next_obs[idxs], ring.done[idxs].bool()
        q_vals = ref_model(st).gather(1, a.unsqueeze(1)).squeeze(1)
        with torch.no_grad():
            na = ref_model(ns).argmax(1, keepdim=True)
            nq = ref_target(ns).gather(1, na).squeeze(1)
        targets = r + 0.99 * nq * (~d)
        ref_loss = (iw * (q_vals - targets).pow(2)).mean()
        ref_opt.zero_grad(); ref_loss.backward(); ref_opt.step()
        ref_prios[idxs] = (q_vals - targets).detach().abs() + 1e-6
        if step % 2 == 0:
            ref_target.load_state_dict(ref_model.state_dict())
        # same batch, same noise; the sampler never materialises the normalised probabilities (pa[idx] / total instead of
        # (pa / total)[idx]), so importance weights agree to rounding, not bit for bit
        assert float(loss) == pytest.approx(float(ref_loss.detach()), rel=1e-5)
        for p, q in zip(tr.model.state_dict().values(), ref_model.state_dict().values()):
            assert torch.allclose(p, q, rtol=1e-5, atol=1e-8)
        for p, q in zip(tr.target.state_dict().values(), ref_target.state_dict().values()):
            assert torch.allclose(p, q, rtol=1e-5, atol=1e-8)
        assert torch.allclose(s1.prios, ref_prios, rtol=1e-5, atol=1e-9)
    assert tr.train_steps == 3 and torch.equal(tr.model.features[0].weight, net.features[0].weight)
    assert not torch.equal(tr.model.fc_A.weight_mu, net.fc_A.weight_mu)


def test_update_waits_for_a_full_batch():
    tr = TorchDQNTrainer(pp.QNet(), batch_size=64, device="cpu")
    ring = _filled_ring(128, 10)
    s = pp.PrioritizedSampler(ring); s.note_new_rows()
    assert tr.update(s) is None and tr.train_steps == 0
