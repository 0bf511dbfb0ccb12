"""DRQN training mode: episode-sequence replay and the last-step Double-DQN update of
scripts/train_rnn_iterative.py around the fused recurrent rollout kernel.

    reference (scripts/train_rnn_iterative.py)             here
    SequenceReplayBuffer.push_step / sample    :100-171    ReplayRing(lockstep_envs=n) written by the rollout kernel +
                                                           SequenceSampler (windows drawn on the device)
    train_step_rnn                             :400-531    DRQNTrainer.update (zero initial (h, c), last-step Q, Double-DQN
                                                           target from the next-obs sequence, Huber loss, grad-clip 1.0)
    rollout + train loop                       :728-800    train_rnn_generation

The rollout (env + both players + replay rows, QNetRNN on the tensor cores or CUDA cores) runs in libpong_b200.so.  The
update itself is PyTorch on the device: the 175 k-parameter net through `nn.LSTM` (cuDNN) forward and backward over
[batch, trace_length, 7] windows — library code around the hot path, like cuBLAS — captured in a CUDA graph; gradients
are averaged over env slabs with one NCCL all-reduce.

Sequence replay on a lock-step ring.  The kernel writes the row of env i at lock-step step t to slot
(t % T) * n + i (T = capacity / n), so each env's transitions are in time order and an episode is a run of rows that
ends with done = 1.  The reference keeps whole episodes of at least `trace_length` steps, draws an episode uniformly
(with replacement) and then a window of `trace_length` consecutive steps uniformly inside it (:126-141).  The same
distribution here: a window ending at row (t, i) is eligible iff its episode is complete inside the ring, has
len >= trace_length and the window lies inside it; its weight is 1 / (len - trace_length + 1), so that every stored
episode carries total weight 1.  What differs: the ring holds the last T lock-step steps of n envs instead of the
last `memory_size` episodes, and an episode cut by the ring's oldest row is dropped.

How the sequential schedule maps to n lock-step envs: as in train.py — `updates_per_chunk` gradient steps follow every
chunk of `chunk` lock-step steps; B's NoisyNet noise is drawn once per chunk (the reference redraws it at every
exploiting action, :381, and trains with whatever noise is current); epsilon follows the per-episode decay on
average.  The reference's 1000-step episode cap (:752) is not applied: with config_rnn.yaml's speed scaling no
episode comes near it.
"""
from __future__ import annotations

import copy

import torch
import torch.nn.functional as F

from . import dist as ppd
from .policy import Policy, QNetRNN, pack_qnetrnn, pack_qnetrnn_tc
from .selfplay import ReplayRing, SelfPlayEngine
from .train import DQNTrainer


class SequenceSampler:
    """Windows of `trace_length` consecutive transitions of one env and one episode, from a lock-step ReplayRing.
    All tables are fixed-shape device tensors updated in place, so sampling can be captured in a CUDA graph."""

    def __init__(self, ring: ReplayRing, trace_length: int = 8, starts_at_episode_start: bool = True):
        if not ring.lockstep_envs:
            raise ValueError("sequence replay needs ReplayRing(lockstep_envs=n)")
        self.ring, self.trace_length = ring, int(trace_length)
        self.n, self.T = ring.lockstep_envs, ring.capacity // ring.lockstep_envs
        self.starts_fresh = bool(starts_at_episode_start)     # step 0 of the ring is the first step of an episode
        self.episodes = 0
        dev = ring.obs.device
        self._cdf = torch.zeros(self.T * self.n, dtype=torch.float64, device=dev)
        self._shift = torch.zeros((), dtype=torch.int64, device=dev)      # physical step of the oldest row
        self._offsets = torch.arange(-(self.trace_length - 1), 1, device=dev)
        self._t = torch.arange(self.T, device=dev).unsqueeze(1).expand(self.T, self.n)

    def __len__(self):
        """Stored episodes that can be sampled (len(memory), :170-171)."""
        return self.episodes

    def window_weights(self):
        """-> (weights float64 [T, n] in time order, oldest step first; done bool [T, n]; shift).  weights[t, i] > 0 iff
        a window may END at step t of env i.  Steps not written yet (before the ring has wrapped) hold done = 0 and so
        belong to no complete episode."""
        ring, T, n, L = self.ring, self.T, self.n, self.trace_length
        steps = ring.steps_written
        wrapped = steps > T
        shift = steps % T if wrapped else 0
        done = ring.done.view(T, n)
        d = (torch.roll(done, -shift, 0) if shift else done) != 0
        dev, t = d.device, self._t
        first = torch.full((1, n), self.starts_fresh and not wrapped, dtype=torch.bool, device=dev)
        prev_done = torch.cat([first, d[:-1]], 0)
        start = torch.cummax(torch.where(prev_done, t, torch.full_like(t, -1)), 0).values        # -1: cut by the ring
        end = torch.flip(torch.cummin(torch.flip(torch.where(d, t, torch.full_like(t, T)), [0]), 0).values, [0])
        ep_len = end - start + 1
        ok = (start >= 0) & (end < T) & (t - start + 1 >= L)
        w = torch.where(ok, 1.0 / (ep_len - L + 1).clamp(min=1).to(torch.float64), torch.zeros((), dtype=torch.float64, device=dev))
        return w, d, shift

    def refresh(self) -> int:
        """Rebuild the sampling table after new rows were written (one pass over the ring on the device; no per-row
        host work).  Returns the number of stored episodes."""
        if self.ring.steps_written == 0:
            self.episodes = 0
            return 0
        w, d, shift = self.window_weights()
        torch.cumsum(w.flatten(), 0, out=self._cdf)
        self._shift.fill_(shift)
        self.episodes = int(((w > 0) & d).sum().item())           # every stored episode ends with exactly one such row
        return self.episodes

    def sample_rows(self, batch_size: int, generator=None):
        """-> ring slots int64 [batch, trace_length], time ascending."""
        if self.episodes == 0:
            raise RuntimeError("no complete episode of at least trace_length steps in the ring")
        total = self._cdf[-1]
        u = torch.rand(batch_size, dtype=torch.float64, device=self._cdf.device, generator=generator) * total
        idx = torch.searchsorted(self._cdf, u, right=True)
        last = torch.searchsorted(self._cdf, total.reshape(1), right=False)       # the last row with a positive weight
        idx = torch.minimum(idx, last)
        t_end, env = idx // self.n, idx % self.n
        t = t_end.unsqueeze(1) + self._offsets.unsqueeze(0)                        # logical time, oldest = 0
        return ((t + self._shift) % self.T) * self.n + env.unsqueeze(1)

    def sample(self, batch_size: int, generator=None):
        """-> (obs [B, L, 7], act int64 [B, L], rew [B, L], next_obs [B, L, 7], done bool [B, L]) — :143-165."""
        rows = self.sample_rows(batch_size, generator)
        r = self.ring
        return r.obs[rows], r.act[rows].to(torch.int64), r.rew[rows], r.next_obs[rows], r.done[rows] != 0


class DRQNTrainer(DQNTrainer):
    """Last-step Double-DQN on QNetRNN, all parameters trained — scripts/train_rnn_iterative.py:335-338,400-531."""

    def __init__(self, model_b: QNetRNN, gamma: float = 0.99, lr: float = 1e-4, batch_size: int = 64,
                 target_update_interval: int = 2000, grad_clip_norm: float = 1.0, min_episodes_factor: int = 1,
                 device="cuda", use_graph: bool = True):
        self.device = torch.device(device)
        self.model = model_b.to(self.device)
        ppd.broadcast_module_(self.model)                # several ranks: every replica starts from rank 0's weights
        self.model.train()                                                               # :729
        self.target = copy.deepcopy(self.model)
        self.target.eval()                                                               # :337-338
        for m in (self.model, self.target):
            m.lstm.flatten_parameters()
        self.params = [p for p in self.model.parameters()]
        self.head_params = self.params                     # what DQNTrainer's helpers call the trainable set
        self.use_graph = use_graph
        on_cuda = self.device.type == "cuda"
        # fused: ONE multi-tensor kernel for the 20 parameter tensors (the capturable foreach form is ~100 tiny kernels)
        self.opt = torch.optim.Adam(self.params, lr=lr, capturable=use_graph and on_cuda, fused=True if on_cuda else None)  # :335
        self._graph, self._eager_runs = None, 0
        self.gamma, self.batch_size, self.target_update_interval = gamma, batch_size, target_update_interval
        self.grad_clip_norm, self.min_episodes = float(grad_clip_norm), int(batch_size * min_episodes_factor)
        self.frame_idx = self.train_steps = 0
        self.fused = False                               # the update is PyTorch (cuDNN LSTM backward)
        self._flat_grad = ppd.flatten_grads_(self.params)     # .grad tensors are views of one buffer: one collective

    def loss_on(self, obs, act, rew, next_obs, done):
        """The loss of train_step_rnn for given windows (:468-507)."""
        b = obs.shape[0]
        h0 = self.model.init_hidden(b, obs.device)
        q_last, _ = self.model(obs, h0)                                                  # :470
        q = q_last.gather(1, act[:, -1].unsqueeze(1)).squeeze(1)                         # :475-478
        with torch.no_grad():
            q_next_online, _ = self.model(next_obs, self.model.init_hidden(b, obs.device))    # :489
            best = q_next_online.argmax(dim=1, keepdim=True)                             # :490
            q_next_target, _ = self.target(next_obs, self.target.init_hidden(b, obs.device))  # :494
            nq = q_next_target.gather(1, best).squeeze(1)                                # :497
            targets = rew[:, -1] + self.gamma * nq * (~done[:, -1])                      # :505
        return F.smooth_l1_loss(q, targets)                                              # :509

    def _pre(self, sampler: SequenceSampler, beta=None, generator=None):
        loss = self.loss_on(*sampler.sample(self.batch_size, generator))
        self.opt.zero_grad(set_to_none=False)
        loss.backward()
        return loss.detach()

    def _post(self, sampler: SequenceSampler):
        torch.nn.utils.clip_grad_norm_(self.params, max_norm=self.grad_clip_norm)        # :516
        self.opt.step()

    def _body(self, sampler, beta=None, generator=None):
        loss = self._pre(sampler, beta, generator)
        self._allreduce_grads()                                                          # one NCCL all-reduce of 175 k floats
        self._post(sampler)
        return loss

    def ready(self, sampler) -> bool:
        """This rank's ring stores at least batch_size complete episodes (:404-407)."""
        return len(sampler) >= max(self.min_episodes, 1)

    def update(self, sampler: SequenceSampler, generator=None, ready: bool | None = None):
        """One train_step_rnn().  None while fewer than batch_size episodes are stored (:404-407).  `ready`: the
        decision when the caller made it collectively (every rank must issue the same gradient all-reduces)."""
        if not (self.ready(sampler) if ready is None else ready):
            return None
        if self.use_graph and generator is None and self.device.type == "cuda":
            loss = self._graphed(sampler, 0.0)
        else:
            loss = self._body(sampler, None, generator)
        self.train_steps += 1
        if self.train_steps % self.target_update_interval == 0:                          # :529-531
            self.target.load_state_dict(self.model.state_dict())
        return loss


def train_rnn_generation(engine: SelfPlayEngine, trainer: DRQNTrainer, ring: ReplayRing, sampler: SequenceSampler,
                         lockstep_steps: int, chunk: int = 16, updates_per_chunk: int = 4, epsilon: float = 1.0,
                         epsilon_decay: float = 0.999, min_epsilon: float = 0.05, precision: str = "f32") -> dict:
    """The rollout + train loop of one generation attempt (scripts/train_rnn_iterative.py:728-800) for the engine's slab:
    recurrent B (epsilon-greedy, train-mode NoisyNet weights of `trainer.model`) learns against the engine's player A
    (any kind).  Returns the counters summed over ranks, the final epsilon and the number of updates."""
    env = engine.env
    env.counters.zero_()
    dev = env.device
    pack = pack_qnetrnn_tc if precision == "f16" else pack_qnetrnn
    eps0, losses, done_steps = float(epsilon), [], 0
    if engine.pb.weights is None or engine.pb.h is None:
        engine.pb = Policy.qnetrnn(trainer.model, num_envs=env.n, noisy=True, eps=epsilon, precision=precision, device=dev)
    while done_steps < lockstep_steps:
        k = min(chunk, lockstep_steps - done_steps)
        trainer.model.reset_noise()                                                      # B's noise: one draw per chunk
        blob = pack(trainer.model, noisy=True).to(dev)
        engine.pb.weights.copy_(blob, non_blocking=True)
        engine.pb.eps = epsilon
        engine.run(k, ring=ring)
        done_steps += k
        sampler.refresh()
        # stored-episode counts differ between slabs: ONE decision per chunk for all ranks (all-reduce MIN of the flag)
        ready = ppd.all_ranks_ready(trainer.ready(sampler), dev)
        for _ in range(updates_per_chunk):
            loss = trainer.update(sampler, ready=ready)
            if loss is not None:
                losses.append(loss)
        episodes = int(env.counters[1].item())
        epsilon = max(min_epsilon, eps0 * epsilon_decay ** (episodes / env.n))           # :800, per env on average
    total = ppd.allreduce_counters(env.counters)
    out = dict(zip(("env_steps", "episodes", "wins_a", "wins_b", "points_a", "points_b", "paddle_hits", "ep_len_sum"),
                   total.tolist()))
    out.update(epsilon=epsilon, updates=len(losses), stored_episodes=len(sampler),
               mean_loss=(float(torch.stack(losses).mean().item()) if losses else None), train_steps=trainer.train_steps)
    return out
