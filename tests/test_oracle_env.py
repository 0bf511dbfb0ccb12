"""Oracle (python port + C restatement) against the golden vectors generated from the reference.

CPU-only.  These pin the oracle; the `-m gpu` tests then pin the CUDA kernels to the oracle.
"""
import hashlib
import json
import os
import random
import struct

import numpy as np
import pytest

from oracle import pong_oracle as po
from oracle import pong_port


@pytest.fixture(scope="module")
def hashes(golden_dir):
    with open(os.path.join(golden_dir, "env_hashes.json")) as f:
        return json.load(f)


def _replay_port(env_cfg, steps):
    """SURVEY.md section 8c protocol, driven through the pure-Python port."""
    random.seed(12345)
    env = pong_port.PongPort(**env_cfg)
    env.reset()
    arng = random.Random(777)
    hs, ho = hashlib.sha256(), hashlib.sha256()
    eps = wa = wb = 0
    for _ in range(steps):
        a, b = arng.randint(0, 2), arng.randint(0, 2)
        (oa, ob), (ra, rb), done, info = env.step(a, b)
        assert info == {}
        hs.update(struct.pack("<7d3i?", *env.state_tuple(), bool(done)))
        ho.update(oa.tobytes() + ob.tobytes() + struct.pack("<2d", ra, rb))
        if done:
            eps += 1; wa += env.scoreA > env.scoreB; wb += env.scoreB > env.scoreA
            env.reset()
    return hs.hexdigest(), ho.hexdigest(), eps, wa, wb


@pytest.mark.parametrize("key,cfgkey", [("config_yaml", "env_config_yaml"), ("config_rnn_yaml", "env_config_rnn_yaml")])
def test_python_port_reproduces_reference_hashes(hashes, key, cfgkey):
    g = hashes[key]
    hs, ho, eps, wa, wb = _replay_port(hashes[cfgkey], g["steps"])
    assert hs == g["state_sha256"]
    assert ho == g["obs_sha256"]
    assert (eps, wa, wb) == (g["episodes"], g["wins_a"], g["wins_b"])


def test_port_config_constants_match_reference_yaml(hashes):
    assert pong_port.CONFIG_YAML_ENV == hashes["env_config_yaml"]
    assert pong_port.CONFIG_RNN_YAML_ENV == hashes["env_config_rnn_yaml"]


def test_collision_known_answers(hashes):
    cfg = hashes["env_config_yaml"]
    p = po.make_params(cfg)
    assert float(p.inertia).hex() == "0x1.797cc39ffd60fp-12"
    assert float(p.two_m_over_7).hex() == "0x1.2492492492492p-2"
    for kat in hashes["collision_kat"]:
        vn, vt, u, om = (float.fromhex(v) for v in kat["inp"])
        want = [float.fromhex(v) for v in kat["out"]]
        got_c = po.collide(p, vn, vt, u, om)
        got_py = pong_port.collide(vn, vt, u, om, cfg["restitution"], cfg["friction"], cfg["ball_mass"],
                                   cfg["world_ball_radius"])
        for w, c, y in zip(want, got_c, got_py):
            assert struct.pack("<d", w) == struct.pack("<d", c) == struct.pack("<d", y)


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint64 if a.dtype == np.float64 else np.uint32)


@pytest.mark.parametrize("fname,cfgkey", [("env_traj_config.npz", "env_config_yaml"), ("env_traj_rnncfg.npz", "env_config_rnn_yaml")])
def test_c_oracle_step_replays_reference_trajectory(golden_dir, hashes, fname, cfgkey):
    """Teacher-free replay: one env, injected serves, 10k (4k) steps -> every bit of state/obs/reward/done."""
    g = dict(np.load(os.path.join(golden_dir, fname)))
    p = po.make_params(hashes[cfgkey])
    b = po.EnvBatch(1, "f64")
    serves = g["serves"]
    k = 0
    b.serve(serves[k, 0], serves[k, 1], serves[k, 2])
    for t in range(g["actions"].shape[0]):
        oa, ob, ra, rb, done = po.step(p, b, g["actions"][t, :1], g["actions"][t, 1:])
        sr, si = b.state_matrix()
        assert np.array_equal(_bits(sr[:, 0]), _bits(g["state"][t])), t
        assert np.array_equal(si[:, 0], g["ints"][t]), t
        assert np.array_equal(_bits(oa[0]), _bits(g["obs"][t, 0])) and np.array_equal(_bits(ob[0]), _bits(g["obs"][t, 1])), t
        assert (ra[0], rb[0]) == tuple(g["rew"][t]) and bool(done[0]) == bool(g["done"][t]), t
        if done[0]:
            k += 1
            b.serve(serves[k, 0], serves[k, 1], serves[k, 2])


def test_c_oracle_rollout_with_serve_pool_matches_trajectory(golden_dir, hashes):
    """The K-step auto-reset rollout (what the fused CUDA kernel mirrors) == step();reset() loop."""
    g = dict(np.load(os.path.join(golden_dir, "env_traj_config.npz")))
    p = po.make_params(hashes["env_config_yaml"])
    serves = g["serves"]
    b = po.EnvBatch(1, "f64")
    b.serve(*serves[0])
    pool = tuple(serves[:, i].reshape(-1, 1) for i in range(3))
    K = g["actions"].shape[0]
    out = po.rollout(p, b, g["actions"].reshape(K, 1, 2), pool, trace=True, log_cap=1024)
    assert np.array_equal(_bits(out["trace_real"][:, :, 0]), _bits(g["state"]))
    assert np.array_equal(out["trace_int"][:, :3, 0], g["ints"])
    assert np.array_equal(out["trace_int"][:, 3, 0] & 1, g["done"])
    c = out["counters"]
    gh = hashes["config_yaml"]
    assert (c[0], c[1], c[2], c[3], c[6]) == (K, gh["episodes"], gh["wins_a"], gh["wins_b"], gh["paddle_hits"])
    assert out["n_log"] == gh["episodes"]
    done_steps = np.nonzero(g["done"])[0]
    lens = np.diff(np.concatenate([[-1], done_steps]))
    assert np.array_equal(out["ep_log"][:, 3], lens)
    assert np.array_equal(out["ep_log"][:, 1], np.arange(gh["episodes"]))


@pytest.mark.parametrize("grp,cfgkey", [("cfg", "env_config_yaml"), ("rnn", "env_config_rnn_yaml")])
def test_c_oracle_single_steps_incl_quirks(golden_dir, hashes, grp, cfgkey):
    g = dict(np.load(os.path.join(golden_dir, "env_random_steps.npz")))
    pre, pre_i, acts = g[f"{grp}/pre"], g[f"{grp}/pre_i"], g[f"{grp}/actions"]
    n = pre.shape[0]
    p = po.make_params(hashes[cfgkey])
    b = po.EnvBatch(n, "f64")
    for j, k in enumerate(po.STATE_REAL):
        getattr(b, k)[:] = pre[:, j]
    for j, k in enumerate(po.STATE_INT):
        getattr(b, k)[:] = pre_i[:, j]
    oa, ob, ra, rb, done = po.step(p, b, acts[:, 0], acts[:, 1])
    sr, si = b.state_matrix()
    assert np.array_equal(_bits(sr.T), _bits(g[f"{grp}/post"]))
    assert np.array_equal(si.T, g[f"{grp}/post_i"])
    assert np.array_equal(_bits(oa), _bits(g[f"{grp}/obs"][:, 0])) and np.array_equal(_bits(ob), _bits(g[f"{grp}/obs"][:, 1]))
    assert np.array_equal(ra, g[f"{grp}/rew"][:, 0]) and np.array_equal(rb, g[f"{grp}/rew"][:, 1])
    assert np.array_equal(done, g[f"{grp}/done"].astype(bool))
    # the fixture really contains the quirk events
    hit = (si.T[:, 2] - pre_i[:, 2]) == 1
    assert hit.sum() > 100 and (ra != 0).sum() > 100 and done.sum() > 10
    assert ((sr[0] < 0) | (sr[0] > 1)).sum() > 10          # x left outside [0,1] by a single reflection


def test_f32_mode_tracks_f64_teacher_forced(golden_dir, hashes):
    """fast mode = same op order in binary32: per-step relative error vs the fp64 reference <= 1e-5
    when re-seeded from the reference state each step (north_star tolerance)."""
    g = dict(np.load(os.path.join(golden_dir, "env_random_steps.npz")))
    pre, pre_i, acts, post = g["cfg/pre"], g["cfg/pre_i"], g["cfg/actions"], g["cfg/post"]
    keep = (np.abs(pre[:, 2]) < 0.2) & (np.arange(pre.shape[0]) % 8 != 6)  # drop 'very fast ball' and exact-edge stressors
    pre32 = pre.astype(np.float32)
    exact = np.all(pre32.astype(np.float64) == pre, axis=1)                # irrelevant; inputs get rounded anyway
    p = po.make_params(hashes["env_config_yaml"])
    n = pre.shape[0]
    b32, b64 = po.EnvBatch(n, "f32"), po.EnvBatch(n, "f64")
    for j, k in enumerate(po.STATE_REAL):
        getattr(b32, k)[:] = pre32[:, j]
        getattr(b64, k)[:] = pre32[:, j].astype(np.float64)                # same (rounded) inputs for both
    for j, k in enumerate(po.STATE_INT):
        getattr(b32, k)[:] = pre_i[:, j]; getattr(b64, k)[:] = pre_i[:, j]
    _, _, ra32, _, d32 = po.step(p, b32, acts[:, 0], acts[:, 1])
    _, _, ra64, _, d64 = po.step(p, b64, acts[:, 0], acts[:, 1])
    same_event = (ra32 == ra64) & (d32 == d64) & (b32.bounce == b64.bounce) & keep
    assert same_event.sum() > 0.97 * keep.sum()                            # knife-edge hit/miss flips are rare
    s32, _ = b32.state_matrix(); s64, _ = b64.state_matrix()
    err = np.abs(s32.astype(np.float64) - s64)[:, same_event]
    # relative to the magnitude of the quantity before/after the step (an impact can cancel spin 40 -> 0.04)
    scale = np.maximum(np.maximum(np.abs(s64[:, same_event]), np.abs(pre32.T.astype(np.float64)[:, same_event])), 1e-2)
    assert (err / scale).max() < 1e-5
    del exact, post


@pytest.mark.parametrize("ci", range(5))
def test_c_oracle_and_port_replay_reference_trajectories_of_other_configs(golden_dir, ci):
    """Constructor keyword sets beyond the YAML blocks (defaults, spin off, fast balls, friction 0 / > 1, max_score 1..5):
    3 000-step trajectories generated from the unmodified reference, replayed bit for bit by the C oracle."""
    from oracle.pong_port import ENV_DEFAULTS, EXTRA_ENV_CONFIGS
    g = {k.split("/", 1)[1]: v for k, v in np.load(os.path.join(golden_dir, "env_extra_cfgs.npz")).items() if k.startswith(f"c{ci}/")}
    cfg = dict(ENV_DEFAULTS, **EXTRA_ENV_CONFIGS[ci])
    if not cfg["ball_angle_intervals"]:
        cfg["ball_angle_intervals"] = [[-60, -30], [30, 60]]
    serves = g["serves"]
    b = po.EnvBatch(1, "f64")
    b.serve(*serves[0])
    K = g["actions"].shape[0]
    out = po.rollout(po.make_params(cfg), b, g["actions"].reshape(K, 1, 2), tuple(serves[:, i].reshape(-1, 1) for i in range(3)), trace=True)
    assert np.array_equal(_bits(out["trace_real"][:, :, 0]), _bits(g["state"]))
    assert np.array_equal(out["trace_int"][:, :3, 0], g["ints"])
    assert np.array_equal(out["trace_int"][:, 3, 0] & 1, g["done"])
    assert g["done"].sum() >= 5
