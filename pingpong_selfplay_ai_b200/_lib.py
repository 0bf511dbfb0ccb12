"""ctypes binding of libpong_b200.so — the C ABI declared in include/pong_b200.h.

Nothing here computes anything: it loads the CUDA library, mirrors its POD structs and turns non-zero
status codes into exceptions.  There is no CPU fallback; a missing library or a missing CUDA device
is an error at the first call.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

c_i32, c_i64, c_u64, c_f32, c_f64, c_vp = C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_double, C.c_void_p

PP_ABI_VERSION = 6
MODE_F64, MODE_F32 = 0, 1
SERVE_POOL, SERVE_PHILOX, SERVE_QUEUE = 0, 1, 2
POLICY_QNET, POLICY_QNETRNN, POLICY_FOLLOWER, POLICY_RANDOM = 0, 1, 2, 3
PREC_F32, PREC_F16 = 0, 1
STREAM_ACT_A, STREAM_ACT_B = 1, 2

QNET_BLOB_FLOATS = 4932
QNET_OFF = dict(W1T=0, B1=448, W2T=512, B2=4608, WHT=4672, BH=4928)
RNNTC_BLOB_BYTES = 659968
RNN_BLOB_FLOATS = 157444
RNN_OFF = dict(WF1T=0, BF1=448, WF2T=512, BF2=8704, WGT=8832, BG=139904, WST=140416, BS=156800, WHT=156928,
               BH=157440)


class PPParams(C.Structure):
    _fields_ = [(n, c_f64) for n in (
        "paddle_speed", "half_width", "magnus_factor", "neg_e", "m_1pe", "inertia", "two_m_over_7", "mu", "mass",
        "radius", "speed_scale", "speed_lo", "speed_hi")] + [
        ("angle_lo", c_f64 * 2), ("angle_hi", c_f64 * 2), ("spin_lo", c_f64), ("spin_hi", c_f64),
        ("enable_spin", c_i32), ("max_score", c_i32), ("speed_scale_every", c_i32), ("reserved", c_i32)]


class PPEnvState(C.Structure):
    _fields_ = [(n, c_vp) for n in (
        "ball_x", "ball_y", "ball_vx", "ball_vy", "spin", "top_paddle_x", "bottom_paddle_x",
        "score_a", "score_b", "bounce_count", "ep_idx", "ep_len")]


class PPServeSource(C.Structure):
    _fields_ = [("kind", c_i32), ("depth", c_i32), ("pool_vx", c_vp), ("pool_vy", c_vp), ("pool_spin", c_vp),
                ("seed", c_u64), ("queue_head", c_vp), ("queue_total", c_i64)]


class PPPolicy(C.Structure):
    _fields_ = [("kind", c_i32), ("precision", c_i32), ("eps_threshold", c_u64), ("follower_tol", c_f64),
                ("weights", c_vp), ("h", c_vp), ("c", c_vp)]


class PPRolloutOut(C.Structure):
    _fields_ = [("counters", c_vp), ("ep_log", c_vp), ("ep_log_cap", c_i64), ("ep_log_count", c_vp),
                ("actions_out", c_vp), ("trace_real", c_vp), ("trace_int", c_vp)]


class PPReplayRing(C.Structure):
    _fields_ = [("obs", c_vp), ("act", c_vp), ("rew", c_vp), ("next_obs", c_vp), ("done", c_vp),
                ("capacity", c_i64), ("head", c_vp), ("lockstep_envs", c_i64), ("lockstep_step0", c_i64)]


class PPNoisyLayer(C.Structure):
    _fields_ = [("in_features", c_i32), ("out_features", c_i32),
                ("weight_mu", c_vp), ("weight_sigma", c_vp), ("weight_epsilon", c_vp),
                ("bias_mu", c_vp), ("bias_sigma", c_vp), ("bias_epsilon", c_vp),
                ("grad_weight_mu", c_vp), ("grad_weight_sigma", c_vp), ("grad_bias_mu", c_vp), ("grad_bias_sigma", c_vp)]


class PPAdamParam(C.Structure):
    _fields_ = [("param", c_vp), ("grad", c_vp), ("exp_avg", c_vp), ("exp_avg_sq", c_vp), ("step", c_vp), ("numel", c_i64)]


class PPQNetRNNParams(C.Structure):
    _fields_ = [(n, c_vp) for n in ("f0_w", "f0_b", "f2_w", "f2_b", "w_ih", "w_hh", "b_ih", "b_hh")] + [
        ("shared", PPNoisyLayer), ("v", PPNoisyLayer), ("a", PPNoisyLayer)]


class PPQNetRNNGrads(C.Structure):
    _fields_ = [(n, c_vp) for n in ("f0_w", "f0_b", "f2_w", "f2_b", "w_ih", "w_hh", "b_ih", "b_hh")] + [
        ("shared", PPNoisyLayer), ("v", PPNoisyLayer), ("a", PPNoisyLayer)]


class PPPeerBlocks(C.Structure):
    _fields_ = [("blocks", c_vp * 8), ("rank", c_i32), ("world", c_i32), ("capacity_floats", c_i64)]


P = C.POINTER
_PROTOTYPES = {
    "pp_version": (C.c_int, []),
    "pp_last_error": (C.c_char_p, []),
    "pp_env_step": (C.c_int, [C.c_int, c_i64, P(PPParams), P(PPEnvState), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "pp_env_observe": (C.c_int, [C.c_int, c_i64, P(PPEnvState), c_vp, c_vp, c_vp]),
    "pp_env_serve": (C.c_int, [C.c_int, c_i64, P(PPEnvState), c_vp, c_vp, c_vp, c_vp, c_vp]),
    "pp_env_reset": (C.c_int, [C.c_int, c_i64, P(PPParams), P(PPEnvState), c_vp, P(PPServeSource), c_i64, C.c_int, c_vp]),
    "pp_collide": (C.c_int, [C.c_int, c_i64, P(PPParams), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "pp_env_rollout": (C.c_int, [C.c_int, c_i64, c_i64, P(PPParams), P(PPEnvState), c_vp, P(PPServeSource), c_i32,
                                 c_i64, P(PPRolloutOut), c_vp]),
    "pp_qnet_act": (C.c_int, [c_i64, c_vp, P(PPPolicy), c_u64, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp]),
    "pp_qnetrnn_act": (C.c_int, [c_i64, c_vp, P(PPPolicy), c_vp, c_u64, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp]),
    "pp_selfplay_rollout": (C.c_int, [C.c_int, c_i64, c_i64, P(PPParams), P(PPEnvState), P(PPPolicy), P(PPPolicy),
                                      c_u64, c_i64, P(PPServeSource), c_i32, c_i64, P(PPRolloutOut), P(PPReplayRing),
                                      c_vp]),
    "pp_replay_scatter": (C.c_int, [c_i64, P(PPReplayRing), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "pp_noisy_reset": (C.c_int, [P(PPNoisyLayer), c_i32, c_u64, c_vp, c_vp]),
    "pp_pack_qnet": (C.c_int, [c_vp, c_vp, c_vp, c_vp, P(PPNoisyLayer), P(PPNoisyLayer), c_i32, c_vp, c_vp]),
    "pp_dqn_head_grads": (C.c_int, [P(PPReplayRing), c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, P(PPNoisyLayer), P(PPNoisyLayer),
                                    P(PPNoisyLayer), P(PPNoisyLayer), c_i32, c_i32, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "pp_dqn_workspace_floats": (c_i64, [c_i32]),
    "pp_per_sample": (C.c_int, [c_vp, c_i64, c_f32, c_vp, c_vp, c_u64, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "pp_per_sample_scratch_floats": (c_i64, [c_i64]),
    "pp_adam_step": (C.c_int, [P(PPAdamParam), c_i32, c_f64, c_f64, c_f64, c_f64, c_vp]),
    "pp_adam_step_allreduce": (C.c_int, [P(PPAdamParam), c_i32, c_vp, c_i64, P(PPPeerBlocks), c_vp, c_f64, c_f64, c_f64, c_f64, c_vp]),
    "pp_peer_block_bytes": (c_i64, [c_i64]),
    "pp_host_selfplay_eval": (C.c_int, [C.c_int, C.c_int, c_i64, c_i32, P(PPParams), c_vp, c_vp, c_vp, c_u64, c_i64,
                                        c_vp, c_vp, c_i32, c_i64, c_vp, c_vp, c_i64]),
    "pp_host_release": (C.c_int, [C.c_int]),
    "pp_drqn_grads": (C.c_int, [P(PPReplayRing), c_vp, c_i32, c_i32, P(PPQNetRNNParams), P(PPQNetRNNParams), c_i32, c_i32,
                                c_f32, P(PPQNetRNNGrads), c_vp, c_vp, c_vp, c_vp]),
    "pp_drqn_workspace_floats": (c_i64, [c_i32, c_i32]),
    "pp_seq_window_weights": (C.c_int, [c_vp, c_i64, c_i64, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "pp_seq_expand_rows": (C.c_int, [c_vp, c_i32, c_i32, c_i64, c_i64, c_vp, c_vp]),
    "pp_pack_qnetrnn_tc": (C.c_int, [P(PPQNetRNNParams), c_i32, c_vp, c_vp]),
    "pp_clip_grad_norm": (C.c_int, [c_vp, c_i64, c_f32, c_vp, c_vp, c_vp]),
    "pp_adam_step_multi": (C.c_int, [P(PPAdamParam), c_i32, c_f64, c_f64, c_f64, c_f64, c_vp]),
}

EXPORTS = tuple(_PROTOTYPES)
_lib = None


class PongB200Error(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first when stale and nvcc is present) libpong_b200.so.  Raises when unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path) or (_build.is_stale() and os.environ.get("PP_NO_REBUILD") != "1"):
        try:
            _build.build()
        except Exception as e:  # a stale-but-present library on a box without nvcc is still usable
            if not os.path.exists(path):
                raise PongB200Error(f"libpong_b200.so is missing and cannot be built: {e}") from e
    lib = C.CDLL(path)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError = header/library mismatch: fail loudly
        fn.restype, fn.argtypes = res, args
    got = lib.pp_version()
    if got != PP_ABI_VERSION:
        raise PongB200Error(f"libpong_b200.so ABI {got} != binding ABI {PP_ABI_VERSION}: rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().pp_last_error().decode("utf-8", "replace")
        raise PongB200Error(f"{what or 'libpong_b200'} failed with status {rc}: {msg}")
