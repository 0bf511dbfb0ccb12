"""Shared helpers of the `-m gpu` parity tests: same seeded inputs into the CUDA path (through the C ABI of
libpong_b200.so) and into the oracle, bitwise comparison."""
import json
import os

import numpy as np
import torch

import pingpong_selfplay_ai_b200 as pp
from oracle import pong_oracle as po

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def hashes():
    with open(os.path.join(GOLDEN, "env_hashes.json")) as f:
        return json.load(f)


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({8: np.uint64, 4: np.uint32}[a.dtype.itemsize])


def np_of(t):
    return t.detach().cpu().numpy()


def load_state(env: "pp.VecPongEnv2P", real, ints):
    """real [n,7], ints [n,3] -> device SoA."""
    n = env.n
    env._real[:, :n].copy_(torch.as_tensor(np.ascontiguousarray(real.T), dtype=env.real_dtype))
    env._int[:3, :n].copy_(torch.as_tensor(np.ascontiguousarray(ints.T), dtype=torch.int32))


def read_state(env):
    n = env.n
    return np_of(env._real[:, :n]), np_of(env._int[:, :n])


def oracle_batch_like(env, mode):
    """EnvBatch holding the env's current device state."""
    b = po.EnvBatch(env.n, mode)
    real, ints = read_state(env)
    for j, k in enumerate(po.STATE_REAL):
        getattr(b, k)[:] = real[j]
    for j, k in enumerate(po.STATE_INT + ("ep_idx", "ep_len")):
        getattr(b, k)[:] = ints[j]
    return b


def assert_state_equal(env, b, what=""):
    real, ints = read_state(env)
    sr, si = b.state_matrix()
    assert np.array_equal(bits(real), bits(sr)), f"real state differs {what}"
    assert np.array_equal(ints[:3], si), f"int state differs {what}"
    assert np.array_equal(ints[3], b.ep_idx) and np.array_equal(ints[4], b.ep_len), f"episode bookkeeping differs {what}"


def make_pool(seed, n, depth, cfg, mode):
    vx, vy, sp = po.serve_pool_from_reference_rng(seed, n, depth, cfg)
    rt = np.float64 if mode == "f64" else np.float32
    return tuple(a.astype(rt) for a in (vx, vy, sp))


def random_actions(seed, k, n, with_invalid=False):
    rs = np.random.RandomState(seed)
    return rs.randint(0, 4 if with_invalid else 3, size=(k, n, 2)).astype(np.uint8)


def golden_sd(g, name):
    pre = name + "/"
    return {k[len(pre):]: v for k, v in g.items() if k.startswith(pre) and not k[len(pre):].startswith(("q_", "h_", "c_"))}
