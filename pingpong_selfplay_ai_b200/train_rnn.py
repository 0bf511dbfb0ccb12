"""DRQN training mode: episode-sequence replay and the last-step Double-DQN update of
scripts/train_rnn_iterative.py around the fused recurrent rollout kernel.

    reference (scripts/train_rnn_iterative.py)             here
    SequenceReplayBuffer.push_step / sample    :100-171    ReplayRing(lockstep_envs=n) written by the rollout kernel +
                                                           SequenceSampler (windows drawn on the device)
    train_step_rnn                             :400-531    DRQNTrainer.update (zero initial (h, c), last-step Q, Double-DQN
                                                           target from the next-obs sequence, Huber loss, grad-clip 1.0)
    rollout + train loop                       :728-800    train_rnn_generation

The rollout (env + both players + replay rows, QNetRNN on the tensor cores or CUDA cores) runs in libpong_b200.so, and so
does the update: pp_drqn_grads (csrc/drqn_kernels.cu) is the forward of the three streams, the last-step Double-DQN
Huber loss and the whole backward pass — heads, BPTT through the LSTM on thread-block clusters, feature layers — in
~15 hand-written launches; pp_clip_grad_norm and pp_adam_step_multi finish train_step_rnn.  All of it is captured in a
CUDA graph; gradients are averaged over env slabs with one NCCL all-reduce of the flat gradient buffer.  There is no CPU
path; the PyTorch formulation (autograd through nn.LSTM) the kernels are tested against is
oracle/train_port.TorchDRQNTrainer (test infrastructure).

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

import ctypes as C

from . import _lib
from . import dist as ppd
from .policy import Policy, QNetRNN, pack_qnetrnn
from .selfplay import ReplayRing, SelfPlayEngine, _ptr, _stream_ptr
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

    def _on_cuda(self) -> bool:
        return self.ring.obs.device.type == "cuda"

    def refresh(self) -> int:
        """Rebuild the sampling table after new rows were written.  On the device: ONE hand-written launch
        (pp_seq_window_weights: a thread per env walks its column of done flags) writes the window weights per ring
        slot and counts the stored episodes; `sample_rows` then draws with pp_per_sample.  The torch formulation below
        is the host-logic mirror the CPU tests compare with SequenceReplayBuffer (it also serves an explicit generator).
        Returns the number of stored episodes."""
        if self.ring.steps_written == 0:
            self.episodes = 0
            return 0
        self._cdf_valid = False
        if self._on_cuda():
            dev = self.ring.obs.device
            if getattr(self, "_k_w", None) is None:
                self._k_lib = _lib.load()
                self._k_w = torch.zeros(self.T * self.n, dtype=torch.float32, device=dev)
                self._k_eps = torch.zeros(1, dtype=torch.int64, device=dev)
                self._k_eps_host = torch.zeros(1, dtype=torch.int64).pin_memory()
            with torch.cuda.device(dev):
                _lib.check(self._k_lib.pp_seq_window_weights(_ptr(self.ring.done), self.n, self.T, int(self.ring.steps_written),
                                                             self.trace_length, int(self.starts_fresh), _ptr(self._k_w), _ptr(self._k_eps),
                                                             _stream_ptr(dev)), "pp_seq_window_weights")
            self._k_eps_host.copy_(self._k_eps, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            self.episodes = int(self._k_eps_host[0])
            return self.episodes
        return self._refresh_torch()

    def _refresh_torch(self) -> int:
        w, d, shift = self.window_weights()
        torch.cumsum(w.flatten(), 0, out=self._cdf)
        self._shift.fill_(shift)
        self.episodes = int(((w > 0) & d).sum().item())           # every stored episode ends with exactly one such row
        self._cdf_valid = True
        return self.episodes

    def sample_rows(self, batch_size: int, generator=None, seed: int = 0):
        """-> ring slots int64 [batch, trace_length], time ascending.  On the device (no explicit generator): window ends
        drawn by pp_per_sample over the weight array (alpha = 1; Philox keyed by seed and a device-side draw counter, so a
        replayed CUDA graph draws fresh windows), expanded by pp_seq_expand_rows: four launches, static buffers."""
        if self.episodes == 0:
            raise RuntimeError("no complete episode of at least trace_length steps in the ring")
        if self._on_cuda() and generator is None:
            dev = self.ring.obs.device
            if getattr(self, "_k_batch", None) != batch_size:
                cap = self.T * self.n
                self._k_batch = batch_size
                self._k_idx = torch.zeros(batch_size, dtype=torch.int64, device=dev)
                self._k_iw = torch.zeros(batch_size, dtype=torch.float32, device=dev)
                self._k_rows = torch.zeros(batch_size, self.trace_length, dtype=torch.int64, device=dev)
                self._k_sums = torch.zeros(int(self._k_lib.pp_per_sample_scratch_floats(cap)), dtype=torch.float32, device=dev)
                if getattr(self, "_k_counter", None) is None:                          # the draw counter outlives a change of batch size
                    self._k_counter = torch.zeros(1, dtype=torch.int64, device=dev)
                self._k_one = torch.ones(2, dtype=torch.float32, device=dev)             # beta, size: the importance weights are unused
            st = _stream_ptr(dev)
            with torch.cuda.device(dev):
                _lib.check(self._k_lib.pp_per_sample(_ptr(self._k_w), self.T * self.n, 1.0, _ptr(self._k_one[0:1]), _ptr(self._k_one[1:2]),
                                                     int(seed) & (2 ** 64 - 1), _ptr(self._k_counter), batch_size, _ptr(self._k_sums),
                                                     _ptr(self._k_idx), _ptr(self._k_iw), st), "pp_per_sample")
                _lib.check(self._k_lib.pp_seq_expand_rows(_ptr(self._k_idx), batch_size, self.trace_length, self.n, self.T,
                                                          _ptr(self._k_rows), st), "pp_seq_expand_rows")
            return self._k_rows
        if not getattr(self, "_cdf_valid", False):
            self._refresh_torch()
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

    FUSED = True                                         # oracle/train_port.TorchDRQNTrainer overrides (test infrastructure)

    def __init__(self, model_b: QNetRNN, gamma: float = 0.99, lr: float = 1e-4, batch_size: int = 64,
                 target_update_interval: int = 2000, grad_clip_norm: float = 1.0, min_episodes_factor: int = 1,
                 device="cuda", use_graph: bool = True, fused: bool | None = None):
        """Forward, loss and backward run in the hand-written kernels of csrc/drqn_kernels.cu, gradient clipping and Adam in
        two more.  `fused` exists for call-site compatibility: False is refused here (the PyTorch / autograd formulation is
        oracle/train_port.TorchDRQNTrainer, test infrastructure)."""
        if fused is False and self.FUSED:
            raise ValueError("the product trainer runs the hand-written update only; the PyTorch formulation it is tested "
                             "against is oracle/train_port.TorchDRQNTrainer")
        self.device = self._check_device(device)
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
        want_fused = self.FUSED
        # torch's own fused multi-tensor Adam serves the fused=False formulation (the capturable foreach form is ~100 kernels)
        self.opt = torch.optim.Adam(self.params, lr=lr, capturable=(use_graph or want_fused) and on_cuda,
                                    fused=True if on_cuda else None)                     # :335
        self._graph, self._eager_runs = None, 0
        self.gamma, self.batch_size, self.target_update_interval = gamma, batch_size, target_update_interval
        self.grad_clip_norm, self.min_episodes = float(grad_clip_norm), int(batch_size * min_episodes_factor)
        self.frame_idx = self.train_steps = 0
        self._flat_grad = ppd.flatten_grads_(self.params)     # .grad tensors are views of one buffer: one collective
        self.fused = want_fused
        if self.fused:
            self._init_fused_rnn()

    # ---- the hand-written update path (csrc/drqn_kernels.cu)
    @staticmethod
    def _noisy(mod, grads: bool):
        g = (lambda p: _ptr(p.grad)) if grads else (lambda p: None)
        return _lib.PPNoisyLayer(mod.in_features, mod.out_features, _ptr(mod.weight_mu), _ptr(mod.weight_sigma),
                                 _ptr(mod.weight_epsilon), _ptr(mod.bias_mu), _ptr(mod.bias_sigma), _ptr(mod.bias_epsilon),
                                 g(mod.weight_mu), g(mod.weight_sigma), g(mod.bias_mu), g(mod.bias_sigma))

    @classmethod
    def _net_struct(cls, m: QNetRNN, grads: bool = False):
        if (m.feature_dim, m.lstm_hidden_dim, m.lstm_layers, m.head_hidden_dim, m.input_dim) != (128, 128, 1, 128, 7):
            raise ValueError("the device DRQN update is built for the reference default 7-64-128 / 1-layer LSTM 128 / head 128")
        f, l = m.features_extractor, m.lstm
        plain = [f[0].weight, f[0].bias, f[2].weight, f[2].bias, l.weight_ih_l0, l.weight_hh_l0, l.bias_ih_l0, l.bias_hh_l0]
        ptrs = [_ptr(t.grad) if grads else _ptr(t) for t in plain]
        layers = [cls._noisy(mod, grads) for mod in (m.fc_shared_head[0], m.fc_V, m.fc_A)]
        return (_lib.PPQNetRNNGrads if grads else _lib.PPQNetRNNParams)(*ptrs, *layers)

    def _init_fused_rnn(self):
        self._lib = _lib.load()
        dev = self.device
        self._on, self._tg = self._net_struct(self.model), self._net_struct(self.target)
        self._gr = self._net_struct(self.model, grads=True)
        self._ws = {}                                            # (batch, trace) -> workspace
        self._loss_buf = torch.zeros(1, dtype=torch.float32, device=dev)
        self._td_buf = torch.zeros(self.batch_size, dtype=torch.float32, device=dev)
        self._norm_out = torch.zeros(2, dtype=torch.float32, device=dev)
        self._norm_scratch = torch.zeros(257, dtype=torch.float32, device=dev)
        for p in self.params:                                    # the optimiser's own state, as a capturable Adam creates it
            st = self.opt.state[p]
            if not st:
                st["step"] = torch.zeros((), dtype=torch.float32, device=dev)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        self._adam = (_lib.PPAdamParam * len(self.params))(*[
            _lib.PPAdamParam(_ptr(p), _ptr(p.grad), _ptr(self.opt.state[p]["exp_avg"]), _ptr(self.opt.state[p]["exp_avg_sq"]),
                             _ptr(self.opt.state[p]["step"]), p.numel()) for p in self.params])

    def reset_noise_and_pack_tc(self, image: torch.Tensor, seed: int = 0):
        """model.reset_noise() + pack_qnetrnn_tc(model, noisy=True) into `image` (a tensor-core player's weights) on the
        device: two launches (pp_noisy_reset, pp_pack_qnetrnn_tc) instead of ~100 framework kernels and a host round trip."""
        if not hasattr(self, "_noise_layers"):
            m = self.model
            self._noise_layers = (_lib.PPNoisyLayer * 3)(*[self._noisy(mod, False) for mod in (m.fc_shared_head[0], m.fc_V, m.fc_A)])
            self._noise_counter = torch.zeros(1, dtype=torch.int64, device=self.device)
        st = _stream_ptr(self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pp_noisy_reset(self._noise_layers, 3, int(seed) & (2 ** 64 - 1), _ptr(self._noise_counter), st),
                       "pp_noisy_reset")
            _lib.check(self._lib.pp_pack_qnetrnn_tc(C.byref(self._on), 1, _ptr(image), st), "pp_pack_qnetrnn_tc")

    def grads_on_rows(self, ring: ReplayRing, rows: torch.Tensor):
        """train_step_rnn up to loss.backward() for given windows (ring slots int64 [batch, trace], time ascending): the
        gradients land in the .grad tensors.  Returns the loss (0-d device tensor)."""
        b, l = int(rows.shape[0]), int(rows.shape[1])
        rows = rows.to(torch.int64).contiguous()
        if (b, l) not in self._ws:
            self._ws[(b, l)] = torch.zeros(int(self._lib.pp_drqn_workspace_floats(b, l)), dtype=torch.float32, device=self.device)
        if self._td_buf.numel() < b:
            self._td_buf = torch.zeros(b, dtype=torch.float32, device=self.device)
        rs = ring.struct()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pp_drqn_grads(C.byref(rs), _ptr(rows), b, l, C.byref(self._on), C.byref(self._tg),
                                               int(self.model.training), int(self.target.training), float(self.gamma),
                                               C.byref(self._gr), _ptr(self._loss_buf), _ptr(self._td_buf), _ptr(self._ws[(b, l)]),
                                               _stream_ptr(self.device)), "pp_drqn_grads")
        return self._loss_buf[0]

    def clip_and_step(self):
        """clip_grad_norm_ (:516) + optimizerB.step() (:517) on the flat gradient buffer: four launches."""
        g = self.opt.param_groups[0]
        st = _stream_ptr(self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pp_clip_grad_norm(_ptr(self._flat_grad), self._flat_grad.numel(), self.grad_clip_norm,
                                                   _ptr(self._norm_out), _ptr(self._norm_scratch), st), "pp_clip_grad_norm")
            _lib.check(self._lib.pp_adam_step_multi(self._adam, len(self.params), float(g["lr"]), float(g["betas"][0]),
                                                    float(g["betas"][1]), float(g["eps"]), st), "pp_adam_step_multi")

    def _pre(self, sampler: SequenceSampler, beta=None, generator=None):
        return self.grads_on_rows(sampler.ring, sampler.sample_rows(self.batch_size, generator))

    def _post(self, sampler: SequenceSampler):
        self.clip_and_step()

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
    eps0, losses, done_steps = float(epsilon), [], 0
    if engine.pb.weights is None or engine.pb.h is None:
        engine.pb = Policy.qnetrnn(trainer.model, num_envs=env.n, noisy=True, eps=epsilon, precision=precision, device=dev)
    while done_steps < lockstep_steps:
        k = min(chunk, lockstep_steps - done_steps)
        if precision == "f16":                                                           # B's noise: one draw per chunk
            trainer.reset_noise_and_pack_tc(engine.pb.weights)
        else:                                            # fp32 rollout kernel: the k-major blob, packed with device-side torch ops
            trainer.model.reset_noise()
            engine.pb.weights.copy_(pack_qnetrnn(trainer.model, noisy=True).to(dev), non_blocking=True)
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
