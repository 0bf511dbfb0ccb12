"""PongEnv2P on the device: `VecPongEnv2P` (n lock-step envs, torch tensors) and the n = 1 drop-in `PongEnv2P`.

Both keep the reference's interface (envs/my_pong_env_2p.py:19-39,83,116): `reset(seed=None, options=None)
-> (obsA, obsB)` and `step(actionA, actionB) -> ((obsA, obsB), (rewardA, rewardB), done, {})`, the 7-D
observation layout (:235-257), the reward convention (:181-186,216-223) and the `env:` keywords of
config.yaml.  All arithmetic happens in libpong_b200.so (csrc/env_kernels.cu); this file owns the HBM
buffers and marshals pointers.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
import random

import numpy as np
import torch

from . import _lib
from .params import make_params, resolve_env_config

_REAL_FIELDS = ("ball_x", "ball_y", "ball_vx", "ball_vy", "spin", "top_paddle_x", "bottom_paddle_x")
_INT_FIELDS = ("score_a", "score_b", "bounce_count", "ep_idx", "ep_len")
COUNTER_NAMES = ("env_steps", "episodes", "wins_a", "wins_b", "points_a", "points_b", "paddle_hits", "ep_len_sum")


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(device) -> torch.device:
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.PongB200Error("pingpong_selfplay_ai_b200 runs on CUDA devices only (there is no CPU path)")
    if not torch.cuda.is_available():
        raise _lib.PongB200Error("no CUDA device is available and pingpong_selfplay_ai_b200 has no CPU fallback")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


class ServePool:
    """Host-generated serves (vx, vy, spin), each [depth, n]: env i's j-th episode uses row j % depth."""

    def __init__(self, vx, vy, spin, device, dtype):
        self.vx, self.vy, self.spin = (torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(device).contiguous()
                                       for a in (vx, vy, spin))
        assert self.vx.ndim == 2 and self.vx.shape == self.vy.shape == self.spin.shape
        self.depth = int(self.vx.shape[0])


class VecPongEnv2P:
    """n independent PongEnv2P environments stepped in lock step on one GPU.

    mode   "f64": the reference's IEEE-double arithmetic bit for bit;  "f32": same operation order in binary32.
    serve  "philox" (device RNG keyed by seed / global env id / episode index: independent of sharding) or a
           (vx, vy, spin) triple of [depth, n] arrays produced on the host with the reference formula.
    State is SoA in HBM: 7 real[n] + 5 int32[n] rows carved from one allocation (16-byte aligned rows).
    """

    def __init__(self, num_envs: int, device="cuda", mode: str = "f64", serve="philox", seed: int = 0,
                 env_id_base: int = 0, **env_kwargs):
        if mode not in ("f64", "f32"):
            raise ValueError("mode must be 'f64' or 'f32'")
        self.device = _require_cuda(device)
        self.lib = _lib.load()
        self.n = self.num_envs = int(num_envs)
        self.mode = mode
        self.mode_id = _lib.MODE_F64 if mode == "f64" else _lib.MODE_F32
        self.real_dtype = torch.float64 if mode == "f64" else torch.float32
        self.cfg = resolve_env_config(env_kwargs)
        self.params = make_params(self.cfg)
        self.seed, self.env_id_base = int(seed), int(env_id_base)
        n, dev = self.n, self.device
        npad = (n + 3) // 4 * 4                                   # rows stay 16-byte aligned
        rs = 8 if mode == "f64" else 4
        sizes = [("_real", 7 * npad * rs), ("_int", 5 * npad * 4), ("obs_a", (n * 28 + 15) // 16 * 16),
                 ("obs_b", (n * 28 + 15) // 16 * 16), ("reward_a", npad * 4), ("reward_b", npad * 4),
                 ("_done", (npad + 15) // 16 * 16)]
        self._arena = torch.zeros(sum(b for _, b in sizes), dtype=torch.uint8, device=dev)   # one allocation: the
        self._arena_off, off = {}, 0                                                         # n = 1 adaptor reads it
        for name, b in sizes:                                                                # back with one copy
            self._arena_off[name] = (off, b)
            off += b
        carve = lambda name, dt: self._arena[self._arena_off[name][0]:sum(self._arena_off[name])].view(dt)
        self._real = carve("_real", self.real_dtype).view(7, npad)
        self._int = carve("_int", torch.int32).view(5, npad)
        for k, name in enumerate(_REAL_FIELDS):
            setattr(self, name, self._real[k, :n])
        for k, name in enumerate(_INT_FIELDS):
            setattr(self, name, self._int[k, :n])
        self._real[5:7].fill_(0.5)
        self._real[0:2].fill_(0.5)
        self.obs_a = carve("obs_a", torch.float32)[:n * 7].view(n, 7)
        self.obs_b = carve("obs_b", torch.float32)[:n * 7].view(n, 7)
        self.reward_a = carve("reward_a", torch.float32)[:n]
        self.reward_b = carve("reward_b", torch.float32)[:n]
        self._done = carve("_done", torch.uint8)[:n]
        self.counters = torch.zeros(8, dtype=torch.int64, device=dev)        # see COUNTER_NAMES
        self._ep_log_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.state = _lib.PPEnvState(*[_ptr(getattr(self, f)) for f in _REAL_FIELDS + _INT_FIELDS])
        self._pool = None
        self.set_serve_source(serve)
        self._served_once = False

    # ------------------------------------------------------------------ reference attribute names
    scoreA = property(lambda self: self.score_a)
    scoreB = property(lambda self: self.score_b)

    @property
    def done(self):
        return self._done.view(torch.bool)

    def set_serve_source(self, serve):
        if isinstance(serve, str):
            if serve != "philox":
                raise ValueError("serve must be 'philox' or a (vx, vy, spin) pool")
            self._pool = None
            self.serve = _lib.PPServeSource(_lib.SERVE_PHILOX, 0, None, None, None, self.seed, None, 0)
        else:
            pool = serve if isinstance(serve, ServePool) else ServePool(*serve, self.device, self.real_dtype)
            if pool.vx.shape[1] != self.n or pool.vx.dtype != self.real_dtype:
                raise ValueError("serve pool must be [depth, num_envs] in the env's real dtype")
            self._pool = pool
            self.serve = _lib.PPServeSource(_lib.SERVE_POOL, pool.depth, _ptr(pool.vx), _ptr(pool.vy), _ptr(pool.spin), 0, None, 0)

    # ------------------------------------------------------------------ reset / step
    def _mask_ptr(self, mask):
        if mask is None:
            return None, None
        m = torch.as_tensor(mask, device=self.device)
        m = m.to(torch.uint8).contiguous() if m.dtype != torch.uint8 else m.contiguous()
        if m.numel() != self.n:
            raise ValueError("mask must have num_envs elements")
        return m, _ptr(m)

    def observe(self):
        """(obsA, obsB) of the current state                        envs/my_pong_env_2p.py:235-263"""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pp_env_observe(self.mode_id, self.n, C.byref(self.state), _ptr(self.obs_a),
                                               _ptr(self.obs_b), _stream_ptr(self.device)), "pp_env_observe")
        return self.obs_a, self.obs_b

    def reset(self, seed=None, options=None, mask=None, serves=None):
        """reset() for all envs (or those in `mask`).  `serves=(vx, vy, spin)` injects the serve of every env;
        otherwise the serve source is used and each reset consumes one serve per env (the first reset of an env
        uses serve 0, like the reference whose constructor already called reset() once consumes draws).
        `seed` re-keys the Philox source; like the reference (:84) it has no effect on injected serves."""
        keep, mptr = self._mask_ptr(mask)
        st = _stream_ptr(self.device)
        with torch.cuda.device(self.device):
            if serves is not None:
                vx, vy, sp = (torch.as_tensor(a, dtype=self.real_dtype).to(self.device).contiguous().reshape(-1)
                              for a in serves)
                if not (vx.numel() == vy.numel() == sp.numel() == self.n):
                    raise ValueError("serves must hold num_envs values each")
                _lib.check(self.lib.pp_env_serve(self.mode_id, self.n, C.byref(self.state), mptr, _ptr(vx), _ptr(vy),
                                                 _ptr(sp), st), "pp_env_serve")
                # pp_env_serve leaves ep_idx alone: the host decides what an injected serve means
                if mask is None:
                    self.ep_len.zero_()
                else:
                    self.ep_len.masked_fill_(keep.view(torch.bool), 0)
            else:
                if seed is not None and self._pool is None:
                    self.seed = int(seed)
                    self.serve.seed = self.seed
                advance = 1 if self._served_once else 0
                _lib.check(self.lib.pp_env_reset(self.mode_id, self.n, C.byref(self.params), C.byref(self.state), mptr,
                                                 C.byref(self.serve), self.env_id_base, advance, st), "pp_env_reset")
                self._served_once = True
        return self.observe()

    @staticmethod
    def _as_action(a, n, device):
        t = torch.as_tensor(a, device=device)
        if t.dtype != torch.uint8:
            # any value other than 0 / 2 means "stay" (my_pong_env_2p.py:118-121); keep that through the narrowing
            t = torch.where((t == 0) | (t == 2), t, torch.ones_like(t)).to(torch.uint8)
        t = t.reshape(-1).contiguous()
        if t.numel() != n:
            raise ValueError("one action per env expected")
        return t

    def step(self, action_a, action_b):
        """One reference step() for every env.  No auto-reset: call reset(mask=done) like the reference's callers do."""
        aa = self._as_action(action_a, self.n, self.device)
        ab = self._as_action(action_b, self.n, self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pp_env_step(self.mode_id, self.n, C.byref(self.params), C.byref(self.state), _ptr(aa),
                                            _ptr(ab), _ptr(self.obs_a), _ptr(self.obs_b), _ptr(self.reward_a),
                                            _ptr(self.reward_b), _ptr(self._done), _stream_ptr(self.device)),
                       "pp_env_step")
        return (self.obs_a, self.obs_b), (self.reward_a, self.reward_b), self.done, {}

    # ------------------------------------------------------------------ multi-step
    def make_rollout_out(self, k=0, log_cap=0, want_actions=False, trace=False, ep_log=None):
        """Per-launch outputs.  `ep_log` (int32 [cap, 4]) may be passed in to keep appending to one log over several
        launches (the log cursor `_ep_log_count` is never reset by a launch)."""
        n, dev = self.n, self.device
        if ep_log is not None:
            log_cap = int(ep_log.shape[0])
        bufs = dict(
            ep_log=ep_log if ep_log is not None else (torch.zeros(max(log_cap, 1), 4, dtype=torch.int32, device=dev) if log_cap else None),
            actions=torch.zeros(k, n, 2, dtype=torch.uint8, device=dev) if want_actions else None,
            trace_real=torch.zeros(k, 7, n, dtype=self.real_dtype, device=dev) if trace else None,
            trace_int=torch.zeros(k, 4, n, dtype=torch.int32, device=dev) if trace else None)
        out = _lib.PPRolloutOut(_ptr(self.counters), _ptr(bufs["ep_log"]), int(log_cap), _ptr(self._ep_log_count),
                                _ptr(bufs["actions"]), _ptr(bufs["trace_real"]), _ptr(bufs["trace_int"]))
        return out, bufs

    def rollout(self, actions, quota: int = 0, log_cap: int = 0, trace: bool = False):
        """k lock-step steps from an injected action stream actions[k, n, 2] (uint8) with auto-reset from the serve
        source — the `step(); if done: reset()` loop of scripts/train_iterative.py:174-179 in one launch."""
        actions = torch.as_tensor(actions, device=self.device)
        if actions.dtype != torch.uint8 or actions.ndim != 3 or actions.shape[1:] != (self.n, 2):
            raise ValueError("actions must be uint8 [k, num_envs, 2]")
        actions = actions.contiguous()
        k = int(actions.shape[0])
        out, bufs = self.make_rollout_out(k, log_cap, False, trace)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pp_env_rollout(self.mode_id, self.n, k, C.byref(self.params), C.byref(self.state),
                                               _ptr(actions), C.byref(self.serve), int(quota), self.env_id_base,
                                               C.byref(out), _stream_ptr(self.device)), "pp_env_rollout")
        self._served_once = True
        return bufs

    def read_counters(self) -> dict:
        return dict(zip(COUNTER_NAMES, self.counters.tolist()))

    def ep_log_count(self) -> int:
        return int(self._ep_log_count.item())

    def close(self):
        pass


def collide_batch(vn, vt, u, omega, e, mu, m, R, mode: str = "f64", device="cuda"):
    """collide_sphere_with_moving_plane (envs/physics.py:3-23) for arrays of impacts on the device (pp_collide) ->
    (vn_post, vt_post, omega_post) tensors.  e, mu, m, R are scalars; their derived constants are formed with the
    reference's own Python expressions (physics.py:7-11)."""
    dev = _require_cuda(device)
    lib = _lib.load()
    dt = torch.float64 if mode == "f64" else torch.float32
    ins = [torch.as_tensor(np.asarray(a, dtype=np.float64)).to(dev, dt).reshape(-1).contiguous() for a in (vn, vt, u, omega)]
    n = ins[0].numel()
    if any(t.numel() != n for t in ins):
        raise ValueError("vn, vt, u, omega must have the same number of elements")
    p = _lib.PPParams()
    p.neg_e, p.m_1pe, p.inertia = float(-e), float(m * (1 + e)), float((2 / 5) * m * R ** 2)
    p.two_m_over_7, p.mu, p.mass, p.radius = float(2 * m / 7.0), float(mu), float(m), float(R)
    outs = [torch.empty(n, dtype=dt, device=dev) for _ in range(3)]
    with torch.cuda.device(dev):
        _lib.check(lib.pp_collide(_lib.MODE_F64 if mode == "f64" else _lib.MODE_F32, n, C.byref(p), *[_ptr(t) for t in ins],
                                  *[_ptr(t) for t in outs], _stream_ptr(dev)), "pp_collide")
    return tuple(outs)


def collide_sphere_with_moving_plane(vn, vt, u, omega, e, mu, m, R):
    """Drop-in for envs/physics.py:3 — one impact, Python floats in, a tuple of Python floats out."""
    out = collide_batch([vn], [vt], [u], [omega], e, mu, m, R)
    return tuple(float(t.item()) for t in out)


class _MultiDiscrete:
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)


class _Box:
    def __init__(self, low, high, dtype):
        self.low, self.high, self.dtype, self.shape = low, high, dtype, low.shape


class PongEnv2P:
    """Drop-in for the reference class of the same name (envs/my_pong_env_2p.py:10), n = 1, bit-exact.

    The serve is drawn on the host from Python's global `random` module in the reference's order (speed, coin,
    angle, spin — :98-111), so `random.seed(s)` reproduces the reference's episodes exactly, the constructor
    consumes one serve like the reference's (:81), and `reset(seed=...)` does not touch that RNG (:84).
    One step = one kernel launch + one device-to-host copy; returns numpy float32 observations, Python float
    rewards and a Python bool, exactly the reference's types.  Public attributes can be read and assigned.
    """

    def __init__(self, device="cuda", **env_kwargs):
        cfg = resolve_env_config(env_kwargs)
        if cfg["enable_render"]:
            raise NotImplementedError("rendering is outside the B200 hot path; use the reference viewer")
        for k, v in cfg.items():
            setattr(self, k, v)
        self._vec = VecPongEnv2P(1, device=device, mode="f64", **env_kwargs)
        self.action_space = _MultiDiscrete([3, 3])
        self.observation_space = _Box(np.array([0, 0, -1, -1, 0, 0, -10], dtype=np.float32),
                                      np.array([1, 1, 1, 1, 1, 1, 10], dtype=np.float32), np.float32)
        self._host = torch.zeros(self._vec._arena.numel(), dtype=torch.uint8).pin_memory()
        hv = lambda name, dt: self._host[self._vec._arena_off[name][0]:sum(self._vec._arena_off[name])].view(dt).numpy()
        self._host_real = hv("_real", torch.float64).reshape(7, 4)[:, 0]       # views into the pinned mirror
        self._host_int = hv("_int", torch.int32).reshape(5, 4)[:3, 0]
        self._host_obs = (hv("obs_a", torch.float32)[:7], hv("obs_b", torch.float32)[:7])
        self._host_rew = hv("reward_a", torch.float32)
        self._host_done = hv("_done", torch.uint8)
        self._dirty = False
        self._act = torch.zeros(2, 16, dtype=torch.uint8).pin_memory()     # one 16-byte-aligned row per player
        self._act_dev = torch.zeros(2, 16, dtype=torch.uint8, device=self._vec.device)
        self.spin_angle = 0.0
        self.reset()

    def _pull(self):
        """One device-to-host copy of the whole n = 1 arena (state, both observations, rewards, done)."""
        self._host.copy_(self._vec._arena, non_blocking=True)
        torch.cuda.current_stream(self._vec.device).synchronize()
        self._dirty = False

    def _push(self):
        v = self._vec
        v._real[:, 0].copy_(torch.from_numpy(np.array(self._host_real)))
        v._int[:3, 0].copy_(torch.from_numpy(np.array(self._host_int)))
        self._dirty = False

    def reset(self, seed=None, options=None):
        speed = random.uniform(*self.ball_speed_range)
        which = 0 if random.random() < 0.5 else 1
        rad = math.radians(random.uniform(*self.ball_angle_intervals[which]))
        vx, vy = speed * math.cos(rad), speed * math.sin(rad)
        spin = random.uniform(*self.spin_range)
        self.spin_angle = 0.0
        self._vec.reset(serves=([vx], [vy], [spin]))
        self._pull()
        return self._host_obs[0].copy(), self._host_obs[1].copy()

    def step(self, actionA, actionB):
        if self._dirty:
            self._push()
        a, b = int(actionA), int(actionB)
        self._act[0, 0] = a if a in (0, 2) else 1
        self._act[1, 0] = b if b in (0, 2) else 1
        self._act_dev.copy_(self._act, non_blocking=True)
        self._vec.step(self._act_dev[0, :1], self._act_dev[1, :1])
        self._pull()
        rew = float(self._host_rew[0])
        return (self._host_obs[0].copy(), self._host_obs[1].copy()), (rew, -rew + 0.0), bool(self._host_done[0]), {}

    def _get_obs(self):
        if self._dirty:
            self._push()
        self._vec.observe()
        self._pull()
        return self._host_obs[0].copy(), self._host_obs[1].copy()

    def render(self):
        return None

    def close(self):
        self._vec.close()


def _real_prop(idx):
    def get(self):
        return float(self._host_real[idx])

    def set_(self, v):
        self._host_real[idx] = float(v)
        self._dirty = True
    return property(get, set_)


def _int_prop(idx):
    def get(self):
        return int(self._host_int[idx])

    def set_(self, v):
        self._host_int[idx] = int(v)
        self._dirty = True
    return property(get, set_)


for _i, _name in enumerate(_REAL_FIELDS):
    setattr(PongEnv2P, _name, _real_prop(_i))
for _i, _name in enumerate(("scoreA", "scoreB", "bounce_count")):
    setattr(PongEnv2P, _name, _int_prop(_i))
