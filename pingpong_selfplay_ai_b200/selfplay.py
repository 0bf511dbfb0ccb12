"""Self-play drivers: the lock-step form of the reference's per-step loops.

    reference loop                                              here
    scripts/train_iterative.py:171-181  eval_vs_model           SelfPlayEngine.evaluate / host_selfplay_eval
    scripts/train_iterative.py:238-245  epsilon-greedy rollout  SelfPlayEngine.run(..., ring=ReplayRing)
    tests/arena.py:294-304              match loop              SelfPlayEngine.evaluate (QNet / follower / random)
    per-step model(obs).argmax(1)       :124-130,176-177        qnet_act / qnetrnn_act

Everything numeric happens in libpong_b200.so; this file sequences launches and owns buffers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .env import COUNTER_NAMES, VecPongEnv2P, _ptr, _stream_ptr
from .params import make_params, resolve_env_config
from .policy import Policy


class ReplayRing:
    """Device replay ring of player B's transitions (oB, aB, rB, nB, done): 62 bytes per row over five arrays
    (scripts/train_iterative.py:49-63,243).  `head` counts every row ever written; slot = head % capacity."""

    def __init__(self, capacity: int, device="cuda", lockstep_envs: int = 0):
        """lockstep_envs = n: the [capacity / n][n] time-major layout that sequence replay needs (row of env i at
        lock-step step t -> slot (t % (capacity / n)) * n + i); 0: compacted appends."""
        dev = torch.device(device)
        self.capacity = int(capacity)
        self.lockstep_envs = int(lockstep_envs)
        self.steps_written = 0                          # lock-step layout: the host keeps the time cursor
        if self.lockstep_envs and self.capacity % self.lockstep_envs:
            raise ValueError("lock-step layout: capacity must be a multiple of the number of envs")
        self.obs = torch.zeros(capacity, 7, dtype=torch.float32, device=dev)
        self.act = torch.zeros(capacity, dtype=torch.uint8, device=dev)
        self.rew = torch.zeros(capacity, dtype=torch.float32, device=dev)
        self.next_obs = torch.zeros(capacity, 7, dtype=torch.float32, device=dev)
        self.done = torch.zeros(capacity, dtype=torch.uint8, device=dev)
        self.head = torch.zeros(1, dtype=torch.int64, device=dev)

    def struct(self) -> _lib.PPReplayRing:
        return _lib.PPReplayRing(_ptr(self.obs), _ptr(self.act), _ptr(self.rew), _ptr(self.next_obs), _ptr(self.done),
                                 self.capacity, _ptr(self.head), self.lockstep_envs, self.steps_written)

    def note_lockstep_launch(self, n: int, k: int):
        """Lock-step layout: the kernel leaves the cursors to the host — advance them by the k steps just launched."""
        self.steps_written += int(k)
        self.head += int(n) * min(int(k), self.capacity // int(n))      # rows written (stream-ordered after the launch)

    def __len__(self):
        return min(int(self.head.item()), self.capacity)

    def scatter(self, obs, act, rew, next_obs, done, valid=None):
        """memory.push for a batch of rows (rows with valid == 0 are skipped)."""
        n = int(obs.shape[0])
        lib = _lib.load()
        f = lambda t, dt: None if t is None else torch.as_tensor(t, device=self.obs.device).to(dt).contiguous()
        obs, next_obs, rew = f(obs, torch.float32), f(next_obs, torch.float32), f(rew, torch.float32)
        act, done, valid = f(act, torch.uint8), f(done, torch.uint8), f(valid, torch.uint8)
        ring = self.struct()
        with torch.cuda.device(self.obs.device):
            _lib.check(lib.pp_replay_scatter(n, C.byref(ring), _ptr(obs), _ptr(act), _ptr(rew), _ptr(next_obs),
                                             _ptr(done), _ptr(valid), _stream_ptr(self.obs.device)), "pp_replay_scatter")


def qnet_act(obs, policy: Policy, seed=0, step_index=0, env_id_base=0, stream_id=_lib.STREAM_ACT_A, want_q=False):
    """obs[n, 7] -> (actions u8[n], q[n, 3] or None): the reference's `model(obs).argmax(1)` for one player."""
    lib = _lib.load()
    obs = obs.to(torch.float32).contiguous()
    n = int(obs.shape[0])
    actions = torch.empty(n, dtype=torch.uint8, device=obs.device)
    q = torch.empty(n, 3, dtype=torch.float32, device=obs.device) if want_q else None
    ps = policy.struct()
    with torch.cuda.device(obs.device):
        _lib.check(lib.pp_qnet_act(n, _ptr(obs), C.byref(ps), int(seed), int(step_index), int(env_id_base),
                                   int(stream_id), _ptr(actions), _ptr(q), _stream_ptr(obs.device)), "pp_qnet_act")
    return actions, q


def qnetrnn_act(obs, policy: Policy, reset_mask=None, seed=0, step_index=0, env_id_base=0,
                stream_id=_lib.STREAM_ACT_A, want_q=False):
    """One QNetRNN step (seq_len 1) with the policy's carried per-env (h, c), updated in place."""
    lib = _lib.load()
    obs = obs.to(torch.float32).contiguous()
    n = int(obs.shape[0])
    actions = torch.empty(n, dtype=torch.uint8, device=obs.device)
    q = torch.empty(n, 3, dtype=torch.float32, device=obs.device) if want_q else None
    m = None if reset_mask is None else torch.as_tensor(reset_mask, device=obs.device).to(torch.uint8).contiguous()
    ps = policy.struct()
    with torch.cuda.device(obs.device):
        _lib.check(lib.pp_qnetrnn_act(n, _ptr(obs), C.byref(ps), _ptr(m), int(seed), int(step_index), int(env_id_base),
                                      int(stream_id), _ptr(actions), _ptr(q), _stream_ptr(obs.device)), "pp_qnetrnn_act")
    return actions, q


class SelfPlayEngine:
    """n lock-step envs + player A + player B on one GPU; `run(k)` is one launch of the fused kernel."""

    def __init__(self, env: VecPongEnv2P, policy_a: Policy, policy_b: Policy, seed: int = 0):
        self.env, self.pa, self.pb = env, policy_a, policy_b
        self.seed = int(seed)
        self.step_base = 0
        self.lib = _lib.load()

    def run(self, k: int, quota: int = 0, ring: ReplayRing | None = None, log_cap: int = 0, want_actions: bool = False,
            serve=None, ep_log=None):
        env = self.env
        out, bufs = env.make_rollout_out(k, log_cap, want_actions, False, ep_log)
        pa, pb = self.pa.struct(), self.pb.struct()
        rs = ring.struct() if ring is not None else None
        serve = env.serve if serve is None else serve
        with torch.cuda.device(env.device):
            _lib.check(self.lib.pp_selfplay_rollout(
                env.mode_id, env.n, int(k), C.byref(env.params), C.byref(env.state), C.byref(pa), C.byref(pb),
                self.seed, self.step_base, C.byref(serve), int(quota), env.env_id_base, C.byref(out),
                C.byref(rs) if rs is not None else None, _stream_ptr(env.device)), "pp_selfplay_rollout")
        self.step_base += int(k)
        env._served_once = True
        if ring is not None and ring.lockstep_envs:
            ring.note_lockstep_launch(env.n, k)
        return bufs

    def _deterministic_players(self) -> bool:
        return all(p.kind != _lib.POLICY_RANDOM and p.eps == 0.0 for p in (self.pa, self.pb))

    def evaluate(self, episodes_per_env: int, chunk: int = 64, max_steps: int = 1 << 20, work_stealing: bool = True,
                 log_cap: int = 0) -> dict:
        """eval_vs_model (scripts/train_iterative.py:171-181) in lock step: num_envs x `episodes_per_env` episodes, a
        fixed number per env so that short games are not over-sampled.
        With a serve pool and deterministic players the serves form one queue (PP_SERVE_QUEUE): an env that finishes
        claims the next unplayed serve, the evaluation is ONE launch, and counters / episode log equal those of the
        per-env quota (an episode depends only on its serve and the players).  Otherwise every env plays its own
        quota and is frozen afterwards, in launches of `chunk` steps."""
        env = self.env
        env.counters.zero_()
        env._ep_log_count.zero_()
        want = env.n * int(episodes_per_env)
        pool = env._pool
        if work_stealing and pool is not None and pool.depth >= episodes_per_env and self._deterministic_players():
            head = torch.full((1,), env.n, dtype=torch.int64, device=env.device)
            src = _lib.PPServeSource(_lib.SERVE_QUEUE, pool.depth, _ptr(pool.vx), _ptr(pool.vy), _ptr(pool.spin), 0,
                                     _ptr(head), want)
            with torch.cuda.device(env.device):
                _lib.check(self.lib.pp_env_reset(env.mode_id, env.n, C.byref(env.params), C.byref(env.state), None,
                                                 C.byref(src), env.env_id_base, 0, _stream_ptr(env.device)), "pp_env_reset")
            bufs = self.run(min(int(max_steps), 0x7ffffffe), quota=want, serve=src, log_cap=log_cap)
            steps = None
        else:
            env.ep_idx.zero_()
            env._served_once = False
            env.reset()
            steps, bufs = 0, None
            log = torch.zeros(log_cap, 4, dtype=torch.int32, device=env.device) if log_cap else None
            while steps < max_steps:
                bufs = self.run(chunk, quota=episodes_per_env, ep_log=log)
                steps += chunk
                if int(env.counters[1].item()) >= want:
                    break
        c = env.read_counters()
        if log_cap:
            c["ep_log"] = bufs["ep_log"][:min(env.ep_log_count(), log_cap)]
        c["win_rate_b"] = c["wins_b"] / max(c["episodes"], 1)
        c["win_rate_a"] = c["wins_a"] / max(c["episodes"], 1)
        c["lockstep_steps"] = steps
        return c


def host_selfplay_eval(env_kwargs: dict, n: int, quota: int, pool, weights_a, weights_b, mode="f64", chunk=None,
                       max_steps=1 << 20, ep_log_cap=0, precision="f32", device: int = 0, seed: int = 0,
                       env_id_base: int = 0):
    """eval_vs_model for n envs x quota episodes from HOST buffers through the C ABI (pp_host_selfplay_eval) on GPU
    `device`: the two packed QNet blobs are numpy arrays; serves are `pool` = three numpy arrays [quota, n], or
    pool=None = drawn on the device from Philox(seed; env_id_base + i, episode).  Returns (counters dict, ep_log).
    Needs no torch CUDA context: everything below this call is the C library's."""
    del chunk                                            # one launch; kept for callers of the round-1 signature
    lib = _lib.load()
    rt = np.float64 if mode == "f64" else np.float32
    params = make_params(resolve_env_config(env_kwargs))
    if pool is not None:
        pvx, pvy, psp = (np.ascontiguousarray(a, dtype=rt) for a in pool)
        assert pvx.shape == (quota, n)
    else:
        pvx = pvy = psp = None
    wa = np.ascontiguousarray(weights_a, dtype=np.float32)
    wb = np.ascontiguousarray(weights_b, dtype=np.float32)
    counters = np.zeros(8, np.uint64)
    log = np.zeros((max(ep_log_cap, 1), 4), np.int32) if ep_log_cap else None
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    _lib.check(lib.pp_host_selfplay_eval(int(device), _lib.MODE_F64 if mode == "f64" else _lib.MODE_F32, n, quota,
                                         C.byref(params), vp(pvx), vp(pvy), vp(psp), int(seed) & (2 ** 64 - 1),
                                         int(env_id_base), vp(wa), vp(wb),
                                         {"f32": _lib.PREC_F32, "f16": _lib.PREC_F16}[precision], max_steps,
                                         vp(counters), vp(log), ep_log_cap), "pp_host_selfplay_eval")
    return dict(zip(COUNTER_NAMES, (int(v) for v in counters))), log


# ------------------------------------------------------------------------------------------ the reference's evaluators
def _play_quota(env_kwargs: dict, policy_a: Policy, policy_b: Policy, n: int, quota: int, seed: int, mode: str, device,
                max_steps: int = 1 << 20) -> dict:
    """n lock-step envs x `quota` games each in ONE launch (an env freezes after its last game) -> counters."""
    env = VecPongEnv2P(n, device=device, mode=mode, serve="philox", seed=seed, **env_kwargs)
    env.reset()
    SelfPlayEngine(env, policy_a, policy_b, seed=seed).run(max_steps, quota=quota)
    c = env.read_counters()
    if c["episodes"] != n * quota:
        raise RuntimeError(f"{c['episodes']} of {n * quota} games finished within the step limit")
    return c


def eval_vs_model(env_kwargs: dict, model_a, model_b, episodes: int, max_envs: int = 65536, seed: int = 0, mode: str = "f64",
                  precision: str = "f32", device="cuda", noisy_a: bool = False, noisy_b: bool = False) -> float:
    """eval_vs_model(env, A, B, episodes) of scripts/train_iterative.py:171-181 -> B's win rate (rB > rA on the last
    step = B reached max_score).  `episodes` games are spread over min(episodes, max_envs) lock-step envs; when that
    does not divide, every env plays one game more (the rate is over the games actually played).  noisy_*: play with
    mu + sigma * epsilon, as the reference's training script does for modelA / modelB (it never calls .eval() on them)."""
    n = max(1, min(int(episodes), int(max_envs)))
    quota = -(-int(episodes) // n)
    c = _play_quota(env_kwargs, Policy.qnet(model_a, noisy=noisy_a, precision=precision, device=device),
                    Policy.qnet(model_b, noisy=noisy_b, precision=precision, device=device), n, quota, seed, mode, device)
    return c["wins_b"] / c["episodes"]


def eval_vs_pool(env_kwargs: dict, model_b, pool: list, episodes: int, seed: int = 0, mode: str = "f64",
                 precision: str = "f32", device="cuda", noisy_b: bool = False, rng=None) -> float:
    """eval_vs_pool(env, B, pool, episodes) of scripts/train_iterative.py:183-196: every game's opponent is
    random.choice(pool) (drawn here for all games up front, from `rng` or the global `random` like the reference), the
    games against one opponent run as one lock-step batch; an empty pool scores 1.0 (:184-185)."""
    if not pool:
        return 1.0
    import random as _random
    pick = (rng or _random).choice
    games = [0] * len(pool)
    for _ in range(int(episodes)):
        games[pick(range(len(pool)))] += 1
    pol_b = Policy.qnet(model_b, noisy=noisy_b, precision=precision, device=device)
    wins = 0
    for k, (opp, g) in enumerate(zip(pool, games)):
        if g:                                            # pool models play in eval mode (:205)
            wins += _play_quota(env_kwargs, Policy.qnet(opp, precision=precision, device=device), pol_b, g, 1, seed + k,
                                mode, device)["wins_b"]
    return wins / int(episodes)
