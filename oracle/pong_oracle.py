"""ctypes front-end of the C oracle (`oracle/libpong_oracle.so`).

TEST INFRASTRUCTURE — only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
cpu_baseline / `--impl reference` leg may import this.  See the header of
`oracle/pong_oracle.c` for what is restated (reference file:line) and how it is pinned.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpong_oracle.so")


class OracleParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "paddle_speed", "half_width", "magnus_factor", "neg_e", "m_1pe", "inertia",
        "two_m_over_7", "mu", "mass", "radius", "speed_scale")] + [
        ("enable_spin", C.c_int32), ("max_score", C.c_int32),
        ("speed_scale_every", C.c_int32), ("pad_", C.c_int32)]


class OraclePolicy(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pad_", C.c_int32), ("eps_threshold", C.c_uint64),
                ("follower_tol", C.c_double)] + [
        (n, C.c_void_p) for n in ("W1", "b1", "W2", "b2", "Wv", "bv", "Wa", "ba")]


POLICY_QNET, POLICY_QNETRNN, POLICY_FOLLOWER, POLICY_RANDOM = 0, 1, 2, 3


def build(force: bool = False) -> str:
    """Compile the oracle library in place (gcc, a second or two)."""
    srcs = [os.path.join(_HERE, f) for f in ("pong_oracle.c", "pong_oracle_step.inc", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)
             or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libpong_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_version.restype = C.c_int
    return _lib


def make_params(cfg: dict) -> OracleParams:
    """Constants computed with the reference's own Python expressions (physics.py:7-11,
    my_pong_env_2p.py:152,230) so that ints from YAML (`restitution: 1`) behave as there."""
    e, m, R = cfg["restitution"], cfg["ball_mass"], cfg["world_ball_radius"]
    p = OracleParams()
    p.paddle_speed = float(cfg["paddle_speed"])
    p.half_width = cfg["paddle_width"] / 2
    p.magnus_factor = float(cfg["magnus_factor"])
    p.neg_e = float(-e)
    p.m_1pe = float(m * (1 + e))
    p.inertia = float((2 / 5) * m * R ** 2)
    p.two_m_over_7 = float(2 * m / 7.0)
    p.mu = float(cfg["friction"])
    p.mass = float(m)
    p.radius = float(R)
    p.speed_scale = 1.0 + cfg["speed_increment"]
    p.enable_spin = int(bool(cfg["enable_spin"]))
    p.max_score = int(cfg["max_score"])
    p.speed_scale_every = int(cfg["speed_scale_every"])
    return p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_REAL = {"f64": np.float64, "f32": np.float32}
STATE_REAL = ("x", "y", "vx", "vy", "spin", "top", "bot")
STATE_INT = ("sa", "sb", "bounce")


class EnvBatch:
    """SoA state of n envs as numpy arrays, in the order the C functions take them."""

    def __init__(self, n: int, mode: str = "f64"):
        self.n, self.mode = n, mode
        rt = _REAL[mode]
        for k in STATE_REAL:
            setattr(self, k, np.full(n, 0.5, dtype=rt))
        for k in ("vx", "vy", "spin"):
            getattr(self, k)[:] = 0
        for k in STATE_INT + ("ep_idx", "ep_len"):
            setattr(self, k, np.zeros(n, dtype=np.int32))

    def serve(self, vx, vy, spin, mask=None):
        rt = _REAL[self.mode]
        m = slice(None) if mask is None else np.asarray(mask, dtype=bool)
        self.x[m] = 0.5; self.y[m] = 0.5; self.top[m] = 0.5; self.bot[m] = 0.5
        self.vx[m] = np.asarray(vx, dtype=rt)[m] if np.ndim(vx) else rt(vx)
        self.vy[m] = np.asarray(vy, dtype=rt)[m] if np.ndim(vy) else rt(vy)
        self.spin[m] = np.asarray(spin, dtype=rt)[m] if np.ndim(spin) else rt(spin)
        self.sa[m] = 0; self.sb[m] = 0; self.bounce[m] = 0; self.ep_len[m] = 0

    def real_ptrs(self):
        return [_ptr(getattr(self, k)) for k in STATE_REAL]

    def int_ptrs(self):
        return [_ptr(getattr(self, k)) for k in STATE_INT]

    def copy(self):
        o = EnvBatch(self.n, self.mode)
        for k in STATE_REAL + STATE_INT + ("ep_idx", "ep_len"):
            getattr(o, k)[:] = getattr(self, k)
        return o

    def state_matrix(self):
        return np.stack([getattr(self, k) for k in STATE_REAL]), np.stack([getattr(self, k) for k in STATE_INT])


def collide(p: OracleParams, vn, vt, u, om, mode="f64"):
    rt = _REAL[mode]
    ct = C.c_double if mode == "f64" else C.c_float
    out = np.zeros(3, dtype=rt)
    getattr(lib(), f"oracle_collide_{mode}")(C.byref(p), ct(vn), ct(vt), ct(u), ct(om), _ptr(out))
    return tuple(out.tolist()) if mode == "f64" else tuple(out)


def step(p: OracleParams, b: EnvBatch, actA, actB):
    """One reference step() for every env.  Returns (obsA, obsB, rewA, rewB, done)."""
    n = b.n
    actA = np.ascontiguousarray(actA, dtype=np.uint8); actB = np.ascontiguousarray(actB, dtype=np.uint8)
    obsA = np.empty((n, 7), np.float32); obsB = np.empty((n, 7), np.float32)
    rewA = np.empty(n, np.float32); rewB = np.empty(n, np.float32); done = np.empty(n, np.uint8)
    getattr(lib(), f"oracle_step_{b.mode}")(
        C.byref(p), C.c_int64(n), *b.real_ptrs(), *b.int_ptrs(), _ptr(actA), _ptr(actB),
        _ptr(obsA), _ptr(obsB), _ptr(rewA), _ptr(rewB), _ptr(done))
    return obsA, obsB, rewA, rewB, done.astype(bool)


def observe(b: EnvBatch):
    obsA = np.empty((b.n, 7), np.float32); obsB = np.empty((b.n, 7), np.float32)
    getattr(lib(), f"oracle_observe_{b.mode}")(C.c_int64(b.n), *b.real_ptrs(), _ptr(obsA), _ptr(obsB))
    return obsA, obsB


def rollout(p: OracleParams, b: EnvBatch, actions, pool, quota=0, env_id_base=0, trace=False, log_cap=0):
    """K lock-step steps with actions[K,n,2] u8 and serve pool (vx,vy,spin) each [depth,n]."""
    K, n = actions.shape[0], b.n
    rt = _REAL[b.mode]
    actions = np.ascontiguousarray(actions, dtype=np.uint8)
    pvx, pvy, psp = (np.ascontiguousarray(a, dtype=rt) for a in pool)
    depth = pvx.shape[0]
    tr = np.zeros((K, 7, n), rt) if trace else None
    ti = np.zeros((K, 4, n), np.int32) if trace else None
    counters = np.zeros(8, np.int64)
    ep_log = np.zeros((max(log_cap, 1), 4), np.int32)
    n_log = np.zeros(1, np.int64)
    getattr(lib(), f"oracle_rollout_{b.mode}")(
        C.byref(p), C.c_int64(n), C.c_int64(K), *b.real_ptrs(), *b.int_ptrs(), _ptr(b.ep_idx), _ptr(b.ep_len),
        _ptr(actions), _ptr(pvx), _ptr(pvy), _ptr(psp), C.c_int32(depth), C.c_int32(quota),
        C.c_int64(env_id_base), _ptr(tr), _ptr(ti), _ptr(counters),
        _ptr(ep_log) if log_cap else None, C.c_int64(log_cap), _ptr(n_log))
    return dict(trace_real=tr, trace_int=ti, counters=counters, ep_log=ep_log[:min(int(n_log[0]), log_cap)],
                n_log=int(n_log[0]))


def qnet_weights_from_state_dict(sd, noisy: bool = False):
    """Effective fp32 matrices of a reference QNet state_dict (models/qnet.py:43-50,56-64):
    eval mode -> mu ; train mode (noisy=True) -> mu + sigma * epsilon."""
    g = lambda k: np.ascontiguousarray(np.asarray(sd[k].detach().cpu().numpy() if hasattr(sd[k], "detach") else sd[k],
                                                  dtype=np.float32))

    def head(prefix):
        w, b = g(prefix + ".weight_mu"), g(prefix + ".bias_mu")
        if noisy:
            w = w + g(prefix + ".weight_sigma") * g(prefix + ".weight_epsilon")
            b = b + g(prefix + ".bias_sigma") * g(prefix + ".bias_epsilon")
        return np.ascontiguousarray(w, dtype=np.float32), np.ascontiguousarray(b, dtype=np.float32)

    Wv, bv = head("fc_V")
    Wa, ba = head("fc_A")
    return dict(W1=g("features.0.weight"), b1=g("features.0.bias"), W2=g("features.2.weight"),
                b2=g("features.2.bias"), Wv=Wv, bv=bv, Wa=Wa, ba=ba)


def qnet_forward(w: dict, obs):
    obs = np.ascontiguousarray(obs, dtype=np.float32).reshape(-1, 7)
    n = obs.shape[0]
    q = np.empty((n, 3), np.float32); a = np.empty(n, np.uint8)
    lib().oracle_qnet_forward(C.c_int64(n), _ptr(obs), *[_ptr(w[k]) for k in ("W1", "b1", "W2", "b2", "Wv", "bv", "Wa", "ba")],
                              _ptr(q), _ptr(a))
    return q, a


def qnetrnn_weights_from_state_dict(sd, noisy: bool = False):
    """Effective matrices of a reference QNetRNN state_dict (models/qnet_rnn.py:71-99)."""
    g = lambda k: np.ascontiguousarray(np.asarray(sd[k].detach().cpu().numpy() if hasattr(sd[k], "detach") else sd[k],
                                                  dtype=np.float32))

    def noisy_layer(prefix):
        w, b = g(prefix + ".weight_mu"), g(prefix + ".bias_mu")
        if noisy:
            w = w + g(prefix + ".weight_sigma") * g(prefix + ".weight_epsilon")
            b = b + g(prefix + ".bias_sigma") * g(prefix + ".bias_epsilon")
        return np.ascontiguousarray(w, dtype=np.float32), np.ascontiguousarray(b, dtype=np.float32)

    Ws, bs = noisy_layer("fc_shared_head.0")
    Wv, bv = noisy_layer("fc_V")
    Wa, ba = noisy_layer("fc_A")
    return dict(Wf1=g("features_extractor.0.weight"), bf1=g("features_extractor.0.bias"),
                Wf2=g("features_extractor.2.weight"), bf2=g("features_extractor.2.bias"),
                Wih=g("lstm.weight_ih_l0"), bih=g("lstm.bias_ih_l0"),
                Whh=g("lstm.weight_hh_l0"), bhh=g("lstm.bias_hh_l0"),
                Ws=Ws, bs=bs, Wv=Wv, bv=bv, Wa=Wa, ba=ba)


_RNN_KEYS = ("Wf1", "bf1", "Wf2", "bf2", "Wih", "bih", "Whh", "bhh", "Ws", "bs", "Wv", "bv", "Wa", "ba")


def qnetrnn_forward(w: dict, obs, h, c):
    """One LSTM-cell step for every env; h, c are [n,H] fp32 and are updated IN PLACE."""
    obs = np.ascontiguousarray(obs, dtype=np.float32).reshape(-1, 7)
    n = obs.shape[0]
    F, H, S = w["Wf2"].shape[0], w["Whh"].shape[1], w["Ws"].shape[0]
    assert h.dtype == np.float32 and c.dtype == np.float32 and h.flags.c_contiguous and c.flags.c_contiguous
    q = np.empty((n, 3), np.float32); a = np.empty(n, np.uint8)
    lib().oracle_qnetrnn_forward(C.c_int64(n), C.c_int(F), C.c_int(H), C.c_int(S), _ptr(obs),
                                 *[_ptr(w[k]) for k in _RNN_KEYS], _ptr(h), _ptr(c), _ptr(q), _ptr(a))
    return q, a


def philox4x32(c0, c1, c2, c3, k0, k1):
    out = np.zeros(4, np.uint32)
    lib().oracle_philox4x32(C.c_uint32(c0), C.c_uint32(c1), C.c_uint32(c2), C.c_uint32(c3),
                            C.c_uint32(k0), C.c_uint32(k1), _ptr(out))
    return out


def philox_serve(seed, env_id, ep_idx, cfg):
    ang = np.asarray(cfg.get("ball_angle_intervals") or [[-60, -30], [30, 60]], dtype=np.float64).reshape(4)
    out = np.zeros(3, np.float64)
    lib().oracle_philox_serve(C.c_uint64(seed), C.c_uint32(env_id), C.c_uint32(ep_idx),
                              C.c_double(cfg["ball_speed_range"][0]), C.c_double(cfg["ball_speed_range"][1]),
                              _ptr(ang), C.c_double(cfg["spin_range"][0]), C.c_double(cfg["spin_range"][1]), _ptr(out))
    return out


def eps_threshold(eps: float) -> int:
    return int(min(max(float(eps), 0.0), 1.0) * 4294967296.0)


def make_policy(kind, weights=None, eps=0.0, tol=0.02):
    """Returns (OraclePolicy, keepalive) — keep `keepalive` referenced while the struct is in use."""
    pol = OraclePolicy()
    pol.kind = kind
    pol.eps_threshold = eps_threshold(eps)
    pol.follower_tol = tol
    keep = weights
    if kind == POLICY_QNET:
        for k in ("W1", "b1", "W2", "b2", "Wv", "bv", "Wa", "ba"):
            setattr(pol, k, weights[k].ctypes.data)
    return pol, keep


def selfplay(p: OracleParams, b: EnvBatch, polA, polB, K, pool, seed=0, step_base=0, quota=0,
             env_id_base=0, log_cap=0, want_actions=False, replay_cap=0):
    """Closed-loop K lock-step steps {act A, act B, step, auto-reset}; see oracle_selfplay_* in C."""
    n = b.n
    rt = _REAL[b.mode]
    pvx, pvy, psp = (np.ascontiguousarray(a, dtype=rt) for a in pool)
    depth = pvx.shape[0]
    counters = np.zeros(8, np.int64)
    ep_log = np.zeros((max(log_cap, 1), 4), np.int32)
    n_log = np.zeros(1, np.int64)
    acts = np.zeros((K, n, 2), np.uint8) if want_actions else None
    rp = None
    n_rp = np.zeros(1, np.int64)
    if replay_cap:
        rp = dict(obs=np.zeros((replay_cap, 7), np.float32), act=np.zeros(replay_cap, np.uint8),
                  rew=np.zeros(replay_cap, np.float32), next=np.zeros((replay_cap, 7), np.float32),
                  done=np.zeros(replay_cap, np.uint8))
    getattr(lib(), f"oracle_selfplay_{b.mode}")(
        C.byref(p), C.c_int64(n), C.c_int64(K), *b.real_ptrs(), *b.int_ptrs(), _ptr(b.ep_idx), _ptr(b.ep_len),
        C.byref(polA[0]), C.byref(polB[0]), C.c_uint64(seed), C.c_int64(step_base),
        _ptr(pvx), _ptr(pvy), _ptr(psp), C.c_int32(depth), C.c_int32(quota), C.c_int64(env_id_base),
        _ptr(acts), _ptr(counters), _ptr(ep_log) if log_cap else None, C.c_int64(log_cap), _ptr(n_log),
        _ptr(rp["obs"]) if rp else None, _ptr(rp["act"]) if rp else None, _ptr(rp["rew"]) if rp else None,
        _ptr(rp["next"]) if rp else None, _ptr(rp["done"]) if rp else None, C.c_int64(replay_cap), _ptr(n_rp))
    out = dict(counters=counters, ep_log=ep_log[:min(int(n_log[0]), log_cap)], n_log=int(n_log[0]), actions=acts)
    if rp:
        k = min(int(n_rp[0]), replay_cap)
        out["replay"] = {key: v[:k] for key, v in rp.items()}
        out["n_replay"] = int(n_rp[0])
    return out


def serve_pool_from_reference_rng(seed: int, n: int, depth: int, cfg: dict):
    """Serve pool [depth, n] x (vx, vy, spin) drawn with the reference's own formula and CPython's
    MT19937 (envs/my_pong_env_2p.py:98-111): env-major draw order, i.e. env i's j-th serve is the
    (i*depth + j)-th reset() of one `random.Random(seed)` stream."""
    import random
    from .pong_port import draw_serve
    rng = random.Random(seed)
    ang = cfg.get("ball_angle_intervals") or [[-60, -30], [30, 60]]
    out = np.empty((3, depth, n), np.float64)
    for i in range(n):
        for j in range(depth):
            out[:, j, i] = draw_serve(rng, cfg["ball_speed_range"], ang, cfg["spin_range"])
    return out[0], out[1], out[2]


# ------------------------------------------------------------------------------------------ env slabs on threads
# Envs are independent, so a batch can be cut into contiguous slabs that run on their own host threads (ctypes
# releases the GIL inside the C calls).  Results are those of the single call except for the ORDER of ep_log /
# replay rows (slab-major instead of step-major) — callers that compare them sort first.
def _slab_view(b: EnvBatch, lo: int, hi: int) -> EnvBatch:
    v = EnvBatch.__new__(EnvBatch)
    v.n, v.mode = hi - lo, b.mode
    for k in STATE_REAL + STATE_INT + ("ep_idx", "ep_len"):
        setattr(v, k, getattr(b, k)[lo:hi])             # contiguous views: the C code updates the parent in place
    return v


def _slabs(n: int, threads):
    import os
    t = max(1, min(int(threads or len(os.sched_getaffinity(0))), max(1, n // 256)))
    cuts = np.linspace(0, n, t + 1).astype(np.int64)
    return [(int(lo), int(hi)) for lo, hi in zip(cuts[:-1], cuts[1:]) if hi > lo]


def _run_slabs(fn, slabs):
    from concurrent.futures import ThreadPoolExecutor
    if len(slabs) == 1:
        return [fn(*slabs[0])]
    with ThreadPoolExecutor(max_workers=len(slabs)) as ex:
        return list(ex.map(lambda s: fn(*s), slabs))


def selfplay_parallel(p, b, polA, polB, K, pool, threads=None, env_id_base=0, log_cap=0, want_actions=False,
                      replay_cap=0, **kw):
    """selfplay() over env slabs on host threads; log_cap / replay_cap are per-batch totals."""
    slabs = _slabs(b.n, threads)
    pool = tuple(np.asarray(a) for a in pool)

    def one(lo, hi):
        sub = tuple(np.ascontiguousarray(a[:, lo:hi]) for a in pool)
        frac = (hi - lo) / b.n
        return selfplay(p, _slab_view(b, lo, hi), polA, polB, K, sub, env_id_base=env_id_base + lo, log_cap=log_cap,
                        want_actions=want_actions, replay_cap=int(np.ceil(replay_cap * frac)) if replay_cap else 0, **kw)
    parts = _run_slabs(one, slabs)
    out = dict(counters=sum(r["counters"] for r in parts), ep_log=np.concatenate([r["ep_log"] for r in parts]),
               n_log=sum(r["n_log"] for r in parts),
               actions=np.concatenate([r["actions"] for r in parts], axis=1) if want_actions else None)
    if replay_cap:
        out["replay"] = {k: np.concatenate([r["replay"][k] for r in parts]) for k in parts[0]["replay"]}
        out["n_replay"] = sum(r["n_replay"] for r in parts)
    return out


def rollout_parallel(p, b, actions, pool, threads=None, env_id_base=0, log_cap=0, quota=0):
    """rollout() (injected action stream) over env slabs on host threads."""
    slabs = _slabs(b.n, threads)
    pool = tuple(np.asarray(a) for a in pool)

    def one(lo, hi):
        sub = tuple(np.ascontiguousarray(a[:, lo:hi]) for a in pool)
        return rollout(p, _slab_view(b, lo, hi), np.ascontiguousarray(actions[:, lo:hi]), sub, quota=quota,
                       env_id_base=env_id_base + lo, log_cap=log_cap)
    parts = _run_slabs(one, slabs)
    return dict(counters=sum(r["counters"] for r in parts), ep_log=np.concatenate([r["ep_log"] for r in parts]),
                n_log=sum(r["n_log"] for r in parts))


def qnetrnn_forward_parallel(w: dict, obs, h, c, threads=None):
    """qnetrnn_forward() over row slabs on host threads; h, c updated in place."""
    n = obs.shape[0]
    obs = np.ascontiguousarray(obs, dtype=np.float32)
    parts = _run_slabs(lambda lo, hi: qnetrnn_forward(w, obs[lo:hi], h[lo:hi], c[lo:hi]), _slabs(n, threads))
    return np.concatenate([q for q, _ in parts]), np.concatenate([a for _, a in parts])
