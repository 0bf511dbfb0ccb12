"""CPU restatement of the reference's training-mode loop (TEST INFRASTRUCTURE and the bench's CPU baseline).

TEST INFRASTRUCTURE — only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import anything under `oracle/`; the product (`pingpong_selfplay_ai_b200/`) never does.

What it restates, line for line, from scripts/train_iterative.py of the reference:
    PrioritizedReplay            :49-76     (numpy priorities, np.random.choice(p=probs), max-priority insertion)
    select_action_B              :124-130   (reset_noise() per call, epsilon-greedy)
    train_step                   :132-168   (PER batch, Double-DQN target, importance-weighted MSE, Adam on the heads,
                                             priorities |td| + 1e-6, hard target sync)
    the rollout + train loop     :238-261   (one train_step per env step, epsilon decay per episode)
and from scripts/train_rnn_iterative.py:
    train_step_rnn               :400-531   (loss on [B, L, 7] windows: last-step Double-DQN, Huber, grad-clip 1.0)
Networks are the torch ports of oracle/policy_torch.py (the reference's modules restated; state_dict-compatible).
"""
from __future__ import annotations

import copy
import random

import numpy as np
import torch
import torch.nn.functional as F


class PrioritizedReplay:
    """scripts/train_iterative.py:49-76"""

    def __init__(self, capacity, alpha=0.6):
        self.cap, self.alpha, self.buffer, self.pos = capacity, alpha, [], 0
        self.prios = np.zeros((capacity,), dtype=np.float32)

    def push(self, trans):
        max_p = self.prios.max() if self.buffer else 1.0
        if len(self.buffer) < self.cap:
            self.buffer.append(trans)
        else:
            self.buffer[self.pos] = trans
        self.prios[self.pos] = max_p
        self.pos = (self.pos + 1) % self.cap

    def sample(self, bs, beta=0.4):
        pr = self.prios if len(self.buffer) == self.cap else self.prios[:self.pos]
        probs = pr ** self.alpha
        probs /= probs.sum()
        idxs = np.random.choice(len(self.buffer), bs, p=probs)
        batch = [self.buffer[i] for i in idxs]
        weights = (len(self.buffer) * probs[idxs]) ** (-beta)
        weights /= weights.max()
        return batch, idxs, torch.tensor(weights)

    def update_priorities(self, idxs, errors):
        for i, e in zip(idxs, errors):
            self.prios[i] = abs(e) + 1e-6


def dqn_loss(model, target, s, a, r, ns, d, iw, gamma):
    """The loss of train_step (:145-158) for given batch tensors -> (loss, td)."""
    q_vals = model(s).gather(1, a.unsqueeze(1)).squeeze(1)
    with torch.no_grad():
        na = model(ns).argmax(1, keepdim=True)
        nq = target(ns).gather(1, na).squeeze(1)
    targets = r + gamma * nq * (~d)
    td = q_vals - targets
    return (iw * td.pow(2)).mean(), td


class TrainLoopPort:
    """The reference's generation loop for ONE env on the CPU: B (epsilon-greedy, NoisyNet heads trained) vs frozen A."""

    def __init__(self, env, model_a, model_b, lr=2.5e-4, gamma=0.99, batch_size=256, memory_size=1_000_000,
                 target_update_interval=1000, epsilon=1.0, epsilon_decay=0.995, min_epsilon=0.02):
        self.env, self.A, self.B = env, model_a, model_b
        for p in self.B.features.parameters():
            p.requires_grad = False                                                   # :97
        self.target = copy.deepcopy(self.B); self.target.eval()                      # :100
        self.opt = torch.optim.Adam(list(self.B.fc_V.parameters()) + list(self.B.fc_A.parameters()), lr=lr)
        self.memory = PrioritizedReplay(memory_size, alpha=0.6)
        self.gamma, self.batch_size, self.target_update_interval = gamma, batch_size, target_update_interval
        self.epsilon, self.epsilon_decay, self.min_epsilon = epsilon, epsilon_decay, min_epsilon
        self.beta_start, self.beta_frames, self.frame_idx, self.train_steps = 0.4, 100000, 0, 0
        self.env_steps = self.episodes = 0

    def select_action_b(self, obs):                                                  # :124-130
        self.B.reset_noise()
        if random.random() < self.epsilon:
            return random.randint(0, 2)
        with torch.no_grad():
            return self.B(torch.tensor(obs, dtype=torch.float32).unsqueeze(0)).argmax(1).item()

    def train_step(self):                                                            # :132-168
        if len(self.memory.buffer) < self.batch_size:
            return None
        self.frame_idx += 1
        beta = min(1.0, self.beta_start + self.frame_idx * (1.0 - self.beta_start) / self.beta_frames)
        batch, idxs, iw = self.memory.sample(self.batch_size, beta)
        self.B.reset_noise(); self.target.reset_noise()
        s = torch.tensor(np.array([b[0] for b in batch]), dtype=torch.float32)
        a = torch.tensor([b[1] for b in batch], dtype=torch.int64)
        r = torch.tensor([b[2] for b in batch], dtype=torch.float32)
        ns = torch.tensor(np.array([b[3] for b in batch]), dtype=torch.float32)
        d = torch.tensor([b[4] for b in batch], dtype=torch.bool)
        loss, td = dqn_loss(self.B, self.target, s, a, r, ns, d, iw.to(torch.float32), self.gamma)
        self.opt.zero_grad(); loss.backward(); self.opt.step()
        self.memory.update_priorities(idxs, td.detach().numpy())
        self.train_steps += 1
        if self.train_steps % self.target_update_interval == 0:
            self.target.load_state_dict(self.B.state_dict())
        return float(loss.detach())

    def run(self, env_steps: int):
        """`env_steps` steps of the loop at :238-261 (episodes continue across calls)."""
        env = self.env
        oA, oB = getattr(self, "_obs", None) or env.reset()
        for _ in range(env_steps):
            with torch.no_grad():
                aA = self.A(torch.tensor(oA, dtype=torch.float32).unsqueeze(0)).argmax(1).item()
            aB = self.select_action_b(oB)
            (nA, nB), (rA, rB), done, _ = env.step(aA, aB)
            self.memory.push((oB, aB, rB, nB, done))
            self.train_step()
            oA, oB = nA, nB
            self.env_steps += 1
            if done:
                self.episodes += 1
                self.epsilon = max(self.min_epsilon, self.epsilon * self.epsilon_decay)     # :261
                oA, oB = env.reset()
        self._obs = (oA, oB)
        return self.env_steps


def drqn_loss(model, target, obs, act, rew, next_obs, done, gamma):
    """The loss of train_step_rnn for given windows (scripts/train_rnn_iterative.py:468-509)."""
    b = obs.shape[0]
    q_last, _ = model(obs, model.init_hidden(b, obs.device))
    q = q_last.gather(1, act[:, -1].unsqueeze(1)).squeeze(1)
    with torch.no_grad():
        q_next_online, _ = model(next_obs, model.init_hidden(b, obs.device))
        best = q_next_online.argmax(dim=1, keepdim=True)
        q_next_target, _ = target(next_obs, target.init_hidden(b, obs.device))
        nq = q_next_target.gather(1, best).squeeze(1)
        targets = rew[:, -1] + gamma * nq * (~done[:, -1])
    return F.smooth_l1_loss(q, targets)


# ------------------------------------------------------------------------------------------ PyTorch formulations of the updates
# The product trainers (pingpong_selfplay_ai_b200.train / .train_rnn) run hand-written CUDA kernels only.  These
# subclasses keep their host logic (batch gating, target sync, CUDA-graph capture, gradient all-reduce) and replace the
# kernels by plain PyTorch / autograd, on any device: the second opinion the kernels are tested against, and what the
# CPU tests compare with the line-by-line restatements above.
def per_sample_torch(sampler, batch_size: int, beta, generator=None):
    """PrioritizedReplay.sample (:64-73) with torch ops on the sampler's device -> (idx int64[bs], weights f32[bs])."""
    if sampler.seen == 0:
        raise RuntimeError("sampling from an empty replay ring")
    pa = sampler.prios.pow(sampler.alpha)
    cdf = pa.cumsum(0)                                   # np.random.choice(p=probs) is inverse-CDF sampling
    total = cdf[-1]
    u = torch.rand(batch_size, device=pa.device, generator=generator) * total
    idx = torch.searchsorted(cdf, u, right=True).clamp_(max=sampler.ring.capacity - 1)
    w = (sampler.size_t * (pa[idx] / total)).pow(-beta)
    return idx, w / w.max()


def _torch_trainers():
    from pingpong_selfplay_ai_b200.train import DQNTrainer
    from pingpong_selfplay_ai_b200.train_rnn import DRQNTrainer

    class TorchDQNTrainer(DQNTrainer):
        FUSED = False

        def _check_device(self, device):
            return torch.device(device)

        def _pre(self, sampler, beta, generator=None):
            ring = sampler.ring
            idx, iw = (sampler.sample(self.batch_size, beta, generator) if "sample" in vars(sampler)
                       else per_sample_torch(sampler, self.batch_size, beta, generator))
            self.model.reset_noise()                                                       # :142-143
            self.target.reset_noise()
            loss, td = dqn_loss(self.model, self.target, ring.obs[idx], ring.act[idx].to(torch.int64), ring.rew[idx],
                                ring.next_obs[idx], ring.done[idx] != 0, iw, self.gamma)
            self.opt.zero_grad(set_to_none=False)
            loss.backward()
            self._idx, self._td = idx, td.detach()
            return loss.detach()

        def _post(self, sampler):
            self.opt.step()
            sampler.update_priorities(self._idx, self._td)                                 # :163-164

    class TorchDRQNTrainer(DRQNTrainer):
        FUSED = False

        def _check_device(self, device):
            return torch.device(device)

        def loss_on(self, obs, act, rew, next_obs, done):
            return drqn_loss(self.model, self.target, obs, act, rew, next_obs, done, self.gamma)

        def _pre(self, sampler, beta=None, generator=None):
            loss = self.loss_on(*sampler.sample(self.batch_size, generator))
            self.opt.zero_grad(set_to_none=False)
            loss.backward()
            return loss.detach()

        def _post(self, sampler):
            torch.nn.utils.clip_grad_norm_(self.params, max_norm=self.grad_clip_norm)      # :516
            self.opt.step()

    return TorchDQNTrainer, TorchDRQNTrainer


TorchDQNTrainer, TorchDRQNTrainer = _torch_trainers()
