"""Training-mode host logic around the device rollout: prioritised replay sampling and the batched Double-DQN
update of scripts/train_iterative.py, with gradients averaged over env slabs (ranks).

    reference                                              here
    PrioritizedReplay.push / sample / update_priorities    PrioritizedSampler (device tensors; rows are written by the
      scripts/train_iterative.py:49-76                       fused rollout kernel / pp_replay_scatter into ReplayRing)
    train_step  :132-168                                   DQNTrainer.update
    rollout + train loop  :233-261                         train_generation

The rollout (env step + both players' actions + replay rows) runs in libpong_b200.so, and so does the update
(csrc/dqn_kernels.cu: prioritised sampling, NoisyNet noise, forward + TD error + head gradients + new priorities, Adam);
what is here sequences those launches, captures them in a CUDA graph and all-reduces the gradients over ranks (NCCL).
There is no CPU path: the trainers refuse non-CUDA devices.  The PyTorch / autograd formulation of the same update,
which the kernels are tested against, lives with the test infrastructure (oracle/train_port.py: TorchDQNTrainer).

How the reference's sequential schedule maps to n lock-step envs ("training semantics do not batch 1:1"):
  * the reference does one gradient step per env step; here `updates_per_chunk` gradient steps follow every chunk
    of `chunk` lock-step steps (n * chunk new transitions), each on `batch_size` rows PER RANK;
  * NoisyNet noise of B is resampled per action there (:125) and per chunk here (the packed weights of a launch are
    mu + sigma * eps for one draw); player A keeps the noise it was built with, as in the reference;
  * epsilon decays per episode there (:261); here epsilon = max(min, decay ** (episodes finished / n)), i.e. every env
    follows the reference's schedule on average (on CUDA the episode count is read one chunk late, through a pinned
    buffer, so that the host never waits for the device).
"""
from __future__ import annotations

import copy
import os

import torch

import ctypes as C

from . import _lib
from . import dist as ppd
from .env import _require_cuda
from .policy import NoisyLinear, Policy, QNet, pack_qnet
from .selfplay import ReplayRing, SelfPlayEngine, _ptr, _stream_ptr


class PrioritizedSampler:
    """Proportional prioritised replay over a ReplayRing (alpha 0.6; new rows get the maximum priority; priorities become
    |TD| + 1e-6 after a sample is trained on) — scripts/train_iterative.py:49-76 on the device.
    `max_prio` is a RUNNING maximum kept on the device (raised by pp_dqn_head_grads whenever it writes a larger
    priority, never lowered): the reference recomputes `prios.max()` over the whole buffer at every push (:57, O(capacity)
    per transition); the running maximum is an upper bound of it that costs nothing per push."""

    def __init__(self, ring: ReplayRing, alpha: float = 0.6):
        self.ring, self.alpha = ring, float(alpha)
        dev = ring.obs.device
        self.prios = torch.zeros(ring.capacity, dtype=torch.float32, device=dev)
        self.max_prio = torch.ones(1, dtype=torch.float32, device=dev)         # 1.0 while nothing has been trained on (:57)
        self.seen = 0                                   # ring.head at the last note_new_rows()
        self.size_t = torch.zeros((), dtype=torch.float32, device=dev)         # len(self) on the device

    def __len__(self):
        return min(self.seen, self.ring.capacity)

    def note_new_rows(self, new_rows: int | None = None) -> int:
        """Give the rows written since the last call the maximum priority (:57,62).  Returns their number.
        `new_rows` = how many rows the caller knows were appended (n * k for a rollout without frozen envs); passing it
        avoids reading the ring cursor back from the device."""
        head = self.seen + int(new_rows) if new_rows is not None else int(self.ring.head.item())
        new = head - self.seen
        if new <= 0:
            return 0
        cap = self.ring.capacity
        if new >= cap:
            self.prios.copy_(self.max_prio.expand_as(self.prios))
        else:
            lo, hi = self.seen % cap, head % cap
            if lo < hi:
                self.prios[lo:hi] = self.max_prio
            else:
                self.prios[lo:] = self.max_prio
                self.prios[:hi] = self.max_prio
        self.seen = head
        self.size_t.fill_(float(len(self)))
        return new

    def sample(self, batch_size: int, beta, seed: int = 0):
        """-> (idx int64[bs], importance weights f32[bs]) — :64-73: `np.random.choice(p = prios^alpha / sum)` as a two-level
        inverse-CDF draw in three hand-written launches (pp_per_sample: chunk sums, one warp per sample, normalisation),
        deterministic for a given seed and call count.  Unfilled slots have priority 0, hence probability 0.  `beta`: float
        or 0-d device tensor.  Returns static buffers (valid until the next call): what a CUDA graph wants."""
        if self.seen == 0:
            raise RuntimeError("sampling from an empty replay ring")
        dev = _require_cuda(self.prios.device)
        if getattr(self, "_f_batch", None) != batch_size:
            self._f_lib = _lib.load()
            self._f_batch = batch_size
            self._f_idx = torch.zeros(batch_size, dtype=torch.int64, device=dev)
            self._f_w = torch.zeros(batch_size, dtype=torch.float32, device=dev)
            self._f_sums = torch.zeros(int(self._f_lib.pp_per_sample_scratch_floats(self.ring.capacity)), dtype=torch.float32, device=dev)
            self._f_counter = torch.zeros(1, dtype=torch.int64, device=dev)
            self._f_beta = torch.zeros((), dtype=torch.float32, device=dev)
        if torch.is_tensor(beta):
            beta_t = beta
        else:
            self._f_beta.fill_(float(beta))
            beta_t = self._f_beta
        with torch.cuda.device(dev):
            _lib.check(self._f_lib.pp_per_sample(_ptr(self.prios), self.ring.capacity, self.alpha, _ptr(beta_t), _ptr(self.size_t),
                                                 int(seed) & (2 ** 64 - 1), _ptr(self._f_counter), batch_size, _ptr(self._f_sums),
                                                 _ptr(self._f_idx), _ptr(self._f_w), _stream_ptr(dev)), "pp_per_sample")
        return self._f_idx, self._f_w

    sample_fused = sample

    def update_priorities(self, idx, td_abs):
        """:74-76 for callers that computed TD errors themselves (the fused update writes priorities in its own kernel)."""
        p = td_abs.detach().abs().to(torch.float32) + 1e-6
        self.prios[idx] = p
        self.max_prio.copy_(torch.maximum(self.max_prio, p.max().reshape(1)))


class DQNTrainer:
    """Double-DQN on the NoisyNet heads of player B (features frozen) — scripts/train_iterative.py:93-104,132-168."""

    def __init__(self, model_b: QNet, gamma: float = 0.99, lr: float = 2.5e-4, batch_size: int = 256,
                 target_update_interval: int = 1000, beta_start: float = 0.4, beta_frames: int = 100000, device="cuda",
                 use_graph: bool = True, fused: bool | None = None, seed: int = 0):
        """Forward, TD error, loss and head gradients run in ONE hand-written kernel (pp_dqn_head_grads), NoisyNet noise
        in another (pp_noisy_reset), sampling and Adam likewise.  `fused` exists for call-site compatibility: False is
        refused here (the PyTorch formulation is oracle/train_port.TorchDQNTrainer, test infrastructure)."""
        if fused is False and self.FUSED:
            raise ValueError("the product trainer runs the hand-written update only; the PyTorch formulation it is tested "
                             "against is oracle/train_port.TorchDQNTrainer")
        self.device = self._check_device(device)
        self.model = model_b.to(self.device)
        ppd.broadcast_module_(self.model)                # several ranks: every replica starts from rank 0's weights
        for p in self.model.features.parameters():                                   # :97
            p.requires_grad = False
        self.target = copy.deepcopy(self.model)
        self.target.eval()                                                             # :100
        self.head_params = list(self.model.fc_V.parameters()) + list(self.model.fc_A.parameters())
        self.use_graph = use_graph
        self.fused = self.FUSED
        capturable = (use_graph or self.fused) and self.device.type == "cuda"          # device-side step counters
        self.opt = torch.optim.Adam(self.head_params, lr=lr, capturable=capturable)    # :101-104
        self._graph, self._eager_runs = None, 0
        self.gamma, self.batch_size, self.target_update_interval = gamma, batch_size, target_update_interval
        self.beta_start, self.beta_frames = beta_start, beta_frames
        self.frame_idx = self.train_steps = 0
        if self.fused:
            self._init_fused(seed)

    FUSED = True                                         # oracle/train_port.TorchDQNTrainer overrides (test infrastructure)

    def _check_device(self, device) -> torch.device:
        return _require_cuda(device)                     # no CPU path: raises without a CUDA device

    # ---- the hand-written update path (csrc/dqn_kernels.cu)
    @staticmethod
    def _layer(mod: NoisyLinear, grads: bool) -> _lib.PPNoisyLayer:
        g = (lambda p: _ptr(p.grad)) if grads else (lambda p: None)
        return _lib.PPNoisyLayer(mod.in_features, mod.out_features, _ptr(mod.weight_mu), _ptr(mod.weight_sigma),
                                 _ptr(mod.weight_epsilon), _ptr(mod.bias_mu), _ptr(mod.bias_sigma), _ptr(mod.bias_epsilon),
                                 g(mod.weight_mu), g(mod.weight_sigma), g(mod.bias_mu), g(mod.bias_sigma))

    def _init_fused(self, seed: int):
        self._lib = _lib.load()
        # the kernel writes the gradients Adam reads: fixed tensors, all views into one flat buffer (one collective)
        self._flat_grad = ppd.flatten_grads_(self.head_params)
        m, t = self.model, self.target
        self._on_v, self._on_a = self._layer(m.fc_V, True), self._layer(m.fc_A, True)
        self._tg_v, self._tg_a = self._layer(t.fc_V, False), self._layer(t.fc_A, False)
        self._noise_all = (_lib.PPNoisyLayer * 4)(self._on_v, self._on_a, self._tg_v, self._tg_a)
        self._noise_model = (_lib.PPNoisyLayer * 2)(self._on_v, self._on_a)
        self._noise_counter = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._noise_seed = int(seed) & (2 ** 64 - 1)
        self._td_buf = torch.zeros(self.batch_size, dtype=torch.float32, device=self.device)
        self._loss_buf = torch.zeros(1, dtype=torch.float32, device=self.device)
        # Adam on the optimiser's own state tensors (created here the way a capturable torch Adam creates them lazily)
        for p in self.head_params:
            st = self.opt.state[p]
            if not st:
                st["step"] = torch.zeros((), dtype=torch.float32, device=self.device)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        self._adam = (_lib.PPAdamParam * len(self.head_params))(*[
            _lib.PPAdamParam(_ptr(p), _ptr(p.grad), _ptr(self.opt.state[p]["exp_avg"]), _ptr(self.opt.state[p]["exp_avg_sq"]),
                             _ptr(self.opt.state[p]["step"]), p.numel()) for p in self.head_params])
        self._workspace = torch.zeros(int(self._lib.pp_dqn_workspace_floats(self.batch_size)), dtype=torch.float32, device=self.device)
        self._init_p2p()

    def _init_p2p(self):
        """Several ranks: peer-mapped gradient blocks for pp_adam_step_allreduce — the all-reduce of the 520 head gradients
        over NVLink fused into the Adam launch (no NCCL node in the update).  The blocks come from
        torch.distributed._symmetric_memory (CUDA VMM handles exchanged between the ranks' processes); every rank sees
        every block in its own address space.  Whether the path is used is ONE decision for all ranks (all-reduce MIN of
        `it worked here`); otherwise the update keeps the NCCL all-reduce.  PP_P2P_ALLREDUCE=0 turns it off."""
        self._p2p, self._p2p_note = None, "single rank"
        if not ppd.is_parallel():
            return
        import torch.distributed as tdist
        ok, err = 0, "disabled by PP_P2P_ALLREDUCE=0"
        if os.environ.get("PP_P2P_ALLREDUCE", "1") == "1" and tdist.get_world_size() <= 8:
            try:
                import torch.distributed._symmetric_memory as symm
                cap = (self._flat_grad.numel() + 63) // 64 * 64
                nfloats = int(self._lib.pp_peer_block_bytes(cap)) // 4
                block = symm.empty(nfloats, dtype=torch.float32, device=self.device)
                block.zero_()
                hdl = symm.rendezvous(block, tdist.group.WORLD)
                ptrs = [int(p) for p in hdl.buffer_ptrs]
                peers = _lib.PPPeerBlocks()
                for r, ptr in enumerate(ptrs):
                    peers.blocks[r] = ptr
                peers.rank, peers.world, peers.capacity_floats = tdist.get_rank(), tdist.get_world_size(), cap
                epoch = torch.zeros(1, dtype=torch.int64, device=self.device)
                torch.cuda.synchronize(self.device)
                ok, cand = 1, (peers, epoch, block, hdl)
            except Exception as e:                                             # noqa: BLE001 — any failure means: use NCCL
                err = f"symmetric memory unavailable: {type(e).__name__}: {e}"
        if ppd.all_ranks_ready(bool(ok), self.device):
            tdist.barrier()                                                    # every block is zeroed before anyone signals
            self._p2p, self._p2p_note = cand, "fused into the Adam kernel over peer-mapped memory (NVLink), no NCCL"
        else:
            self._p2p_note = f"NCCL all-reduce ({err if not ok else 'a peer could not map symmetric memory'})"

    def _adam_step(self):
        g = self.opt.param_groups[0]
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pp_adam_step(self._adam, len(self.head_params), float(g["lr"]), float(g["betas"][0]),
                                              float(g["betas"][1]), float(g["eps"]), _stream_ptr(self.device)), "pp_adam_step")

    def _feature_ptrs(self):
        f = self.model.features                          # frozen (:97) and identical in the target net
        return _ptr(f[0].weight), _ptr(f[0].bias), _ptr(f[2].weight), _ptr(f[2].bias)

    def reset_noise_and_pack(self, blob: torch.Tensor):
        """model.reset_noise() + pack_qnet(model, noisy=True) into `blob` (a player's weights): two launches."""
        st = _stream_ptr(self.device)
        m = self.model
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pp_noisy_reset(self._noise_model, 2, self._noise_seed, _ptr(self._noise_counter), st), "pp_noisy_reset")
            _lib.check(self._lib.pp_pack_qnet(*self._feature_ptrs(), C.byref(self._on_v), C.byref(self._on_a), 1,
                                              _ptr(blob), st), "pp_pack_qnet")

    def _pre(self, sampler: "PrioritizedSampler", beta, generator=None):
        """train_step() up to loss.backward(): local gradients are in p.grad afterwards.  `beta` is a float or a 0-d
        device tensor.  No host synchronisation.  (A `sample` attribute set on the sampler INSTANCE overrides the draw:
        the tests inject fixed batches that way.)"""
        if "sample" in vars(sampler):
            idx, iw = sampler.sample(self.batch_size, beta, generator)
            idx, iw = idx.contiguous(), iw.to(torch.float32).contiguous()
        else:
            idx, iw = sampler.sample(self.batch_size, beta, self._noise_seed)
        ring = sampler.ring.struct()
        st = _stream_ptr(self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pp_noisy_reset(self._noise_all, 4, self._noise_seed, _ptr(self._noise_counter), st), "pp_noisy_reset")  # :142-143
            _lib.check(self._lib.pp_dqn_head_grads(C.byref(ring), _ptr(idx), _ptr(iw), self.batch_size, *self._feature_ptrs(),
                                                   C.byref(self._on_v), C.byref(self._on_a), C.byref(self._tg_v), C.byref(self._tg_a),
                                                   int(self.model.training), int(self.target.training), float(self.gamma),
                                                   _ptr(self._td_buf), _ptr(self._loss_buf), _ptr(sampler.prios), _ptr(sampler.max_prio),
                                                   _ptr(self._workspace), st),
                       "pp_dqn_head_grads")
        self._idx, self._td = idx, self._td_buf
        return self._loss_buf[0]

    def _post(self, sampler: PrioritizedSampler):
        """The rest of train_step(): optimiser step on the (rank-averaged) gradients (one launch; the new priorities were
        written by pp_dqn_head_grads).  With peer-mapped gradient blocks the averaging happens in the same launch."""
        if getattr(self, "_p2p", None) is not None:
            peers, epoch = self._p2p[0], self._p2p[1]
            g = self.opt.param_groups[0]
            with torch.cuda.device(self.device):
                _lib.check(self._lib.pp_adam_step_allreduce(self._adam, len(self.head_params), _ptr(self._flat_grad),
                                                            self._flat_grad.numel(), C.byref(peers), _ptr(epoch), float(g["lr"]),
                                                            float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                                                            _stream_ptr(self.device)), "pp_adam_step_allreduce")
            return
        self._adam_step()

    def _body(self, sampler, beta, generator=None):
        loss = self._pre(sampler, beta, generator)
        self._allreduce_grads()                                                        # one NCCL all-reduce of 520 floats
        self._post(sampler)
        return loss

    def _allreduce_grads(self):
        if getattr(self, "_p2p", None) is not None:       # averaged inside the Adam launch (_post)
            return
        if getattr(self, "_flat_grad", None) is not None:
            ppd.allreduce_mean_flat_(self._flat_grad)
        else:
            ppd.allreduce_mean_grads(self.head_params)

    def ready(self, sampler) -> bool:
        """This rank's ring holds enough rows for a batch (:134-135)."""
        return len(sampler) >= self.batch_size

    def update(self, sampler: PrioritizedSampler, generator=None, ready: bool | None = None):
        """One train_step().  Returns the loss (a 0-d device tensor: no host sync), or None while the ring holds fewer
        than batch_size rows (:134-135).  On CUDA the update is captured into CUDA graphs after three eager calls
        (~150 tiny kernels per update are otherwise bound by launch overhead, 2.7 ms of host time each).
        `ready`: the decision to run, when the caller has made it collectively (dist.all_ranks_ready) — with several
        ranks every rank must run the same number of updates, because each one contains a gradient all-reduce."""
        if not (self.ready(sampler) if ready is None else ready):
            return None
        self.frame_idx += 1
        beta = min(1.0, self.beta_start + self.frame_idx * (1.0 - self.beta_start) / self.beta_frames)
        if self.use_graph and generator is None and self.device.type == "cuda":
            loss = self._graphed(sampler, beta)
        else:
            loss = self._body(sampler, beta, generator)
        self.train_steps += 1
        if self.train_steps % self.target_update_interval == 0:                        # :166-168
            self.target.load_state_dict(self.model.state_dict())
        return loss

    def _graphed(self, sampler, beta: float):
        """Replay of the captured update: ONE CUDA graph — sampling, forward / backward, the NCCL all-reduce of the flat
        gradient buffer (several ranks) and the optimiser step.  The collective is captured like any other node: the
        communicator exists by then (the first three updates run eagerly and ARE its warm-up) and the capture uses
        capture_error_mode="thread_local", so that the CUDA calls of NCCL's watchdog thread do not invalidate it.
        PP_SPLIT_UPDATE_GRAPH=1 selects the older form (two graphs around an eager all-reduce: a host round trip per
        update, which cost 8 GPUs half their update rate)."""
        if self._graph is None:
            if self._eager_runs < 3:                      # the first updates run eagerly: they ARE the warm-up
                self._eager_runs += 1
                return self._body(sampler, beta)
            self._beta_t = torch.zeros((), dtype=torch.float32, device=self.device)
            self._beta_t.fill_(beta)
            self._graph_sampler = sampler
            self._split = os.environ.get("PP_SPLIT_UPDATE_GRAPH") == "1"
            torch.cuda.synchronize(self.device)
            pool = torch.cuda.graph_pool_handle()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, pool=pool, capture_error_mode="thread_local"):
                self._loss_t = self._pre(sampler, self._beta_t) if self._split else self._body(sampler, self._beta_t)
            if self._split:
                self._graph_post = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph_post, pool=pool, capture_error_mode="thread_local"):
                    self._post(sampler)
            self._graph = graph                           # capture records without running: replay below IS this update
        if sampler is not self._graph_sampler:
            raise RuntimeError("the captured update is bound to the sampler (ring) it was built with")
        self._beta_t.fill_(beta)
        self._graph.replay()
        if self._split:
            self._allreduce_grads()
            self._graph_post.replay()
        return self._loss_t.clone()


def train_generation(engine: SelfPlayEngine, trainer: DQNTrainer, ring: ReplayRing, sampler: PrioritizedSampler,
                     lockstep_steps: int, chunk: int = 16, updates_per_chunk: int = 4, epsilon: float = 1.0,
                     epsilon_decay: float = 0.995, min_epsilon: float = 0.02, precision: str = "f32") -> dict:
    """The rollout + train loop of one generation (scripts/train_iterative.py:233-261) for the engine's slab:
    B (epsilon-greedy, train-mode NoisyNet weights of `trainer.model`) learns against the engine's player A.
    Returns the slab's counters summed over ranks, the final epsilon and the number of updates."""
    env = engine.env
    env.counters.zero_()
    dev = env.device
    eps0, losses, done_steps = float(epsilon), [], 0
    episodes, counters_event = 0, None
    counters_host = torch.zeros(8, dtype=torch.int64).pin_memory() if dev.type == "cuda" else None
    if engine.pb.weights is None:
        engine.pb = Policy.qnet(trainer.model, noisy=True, eps=epsilon, precision=precision, device=dev)
    if ring.capacity < env.n * min(chunk, lockstep_steps):
        # the kernel would keep only the last capacity / n steps of a launch while note_new_rows() counts n * k rows:
        # the host's cursor would drift away from the device head and max-priority marking would hit the wrong slots
        raise ValueError(f"replay ring of {ring.capacity} rows is smaller than one chunk ({env.n} envs x {chunk} steps)")
    all_ready = False                                    # once every rank's ring holds a batch it stays that way
    while done_steps < lockstep_steps:
        k = min(chunk, lockstep_steps - done_steps)
        trainer.reset_noise_and_pack(engine.pb.weights)                                # B's noise: one draw per chunk
        engine.pb.eps = epsilon
        engine.run(k, ring=ring)
        done_steps += k
        sampler.note_new_rows(env.n * k)                 # no quota in training mode: every env writes a row per step
        if not all_ready:
            all_ready = ppd.all_ranks_ready(trainer.ready(sampler), dev)
        for _ in range(updates_per_chunk):
            loss = trainer.update(sampler, ready=all_ready)
            if loss is not None:
                losses.append(loss)
        # epsilon follows the slab's own episode count (slabs are statistically identical; no collective needed here).
        # The count is read through a pinned buffer ONE CHUNK LATE, so the host never waits for the device and keeps
        # launching ahead (a synchronous .item() per chunk left the GPU idle for ~0.1 ms of launch latency each time).
        if dev.type == "cuda":
            if counters_event is not None:
                counters_event.synchronize()                                   # recorded a whole chunk ago
                episodes = int(counters_host[1])
            counters_host.copy_(env.counters, non_blocking=True)
            counters_event = torch.cuda.Event()
            counters_event.record(torch.cuda.current_stream(dev))
        else:
            episodes = int(env.counters[1].item())
        epsilon = max(min_epsilon, eps0 * epsilon_decay ** (episodes / env.n))               # :261, per env on average
    total = ppd.allreduce_counters(env.counters)
    out = dict(zip(("env_steps", "episodes", "wins_a", "wins_b", "points_a", "points_b", "paddle_hits", "ep_len_sum"),
                   total.tolist()))
    out.update(epsilon=epsilon, updates=len(losses),
               mean_loss=(float(torch.stack(losses).mean().item()) if losses else None), train_steps=trainer.train_steps)
    return out
