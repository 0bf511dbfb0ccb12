"""K1 parity (`-m gpu`): the CUDA env kernels, called through the C ABI, against the golden vectors generated
from the reference and against the C oracle on the same seeded inputs.  Bit-exact in both modes."""
import hashlib
import os
import random
import struct

import numpy as np
import pytest
import torch

import pingpong_selfplay_ai_b200 as pp
from oracle import pong_oracle as po
import pp_testutil as gu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H():
    return gu.hashes()


@pytest.mark.parametrize("grp,cfgkey", [("cfg", "env_config_yaml"), ("rnn", "env_config_rnn_yaml")])
def test_step_kernel_single_steps_incl_quirks_vs_reference_golden(H, grp, cfgkey):
    """8192 (2048) single steps from injected states covering the quirk list of SURVEY.md 8a: out-of-bounds re-hit,
    paddle pinned at a wall with u != 0, exact paddle edge, vrel = +-0.0, match point, action value 3 = stay."""
    g = dict(np.load(os.path.join(gu.GOLDEN, "env_random_steps.npz")))
    pre, pre_i, acts = g[f"{grp}/pre"], g[f"{grp}/pre_i"], g[f"{grp}/actions"]
    env = pp.VecPongEnv2P(pre.shape[0], mode="f64", **H[cfgkey])
    gu.load_state(env, pre, pre_i)
    (oa, ob), (ra, rb), done, info = env.step(torch.from_numpy(acts[:, 0].copy()), torch.from_numpy(acts[:, 1].copy()))
    assert info == {}
    real, ints = gu.read_state(env)
    assert np.array_equal(gu.bits(real.T), gu.bits(g[f"{grp}/post"]))
    assert np.array_equal(ints[:3].T, g[f"{grp}/post_i"])
    assert np.array_equal(gu.bits(gu.np_of(oa)), gu.bits(g[f"{grp}/obs"][:, 0]))
    assert np.array_equal(gu.bits(gu.np_of(ob)), gu.bits(g[f"{grp}/obs"][:, 1]))
    assert np.array_equal(gu.bits(gu.np_of(ra)), gu.bits(g[f"{grp}/rew"][:, 0]))      # incl. the sign of zero
    assert np.array_equal(gu.bits(gu.np_of(rb)), gu.bits(g[f"{grp}/rew"][:, 1]))
    assert np.array_equal(gu.np_of(done), g[f"{grp}/done"].astype(bool))


@pytest.mark.parametrize("key,cfgkey", [("config_yaml", "env_config_yaml"), ("config_rnn_yaml", "env_config_rnn_yaml")])
def test_dropin_pongenv2p_reproduces_reference_sha256(H, key, cfgkey):
    """BASELINE config 1 through the drop-in class: the protocol of SURVEY.md 8c (global `random` seeded, the
    constructor consumes one serve) must give the reference's sha256 over 10 000 steps of state, obs and rewards."""
    gh = H[key]
    random.seed(12345)
    env = pp.PongEnv2P(**H[cfgkey])
    env.reset()
    assert [float(v).hex() for v in (env.ball_vx, env.ball_vy, env.spin)] == gh["first_serve"]
    arng = random.Random(777)
    hs, ho = hashlib.sha256(), hashlib.sha256()
    eps = wa = wb = 0
    for _ in range(gh["steps"]):
        a, b = arng.randint(0, 2), arng.randint(0, 2)
        (oa, ob), (ra, rb), done, info = env.step(a, b)
        assert type(ra) is float and type(done) is bool and oa.dtype == np.float32 and oa.shape == (7,)
        hs.update(struct.pack("<7d3i?", env.ball_x, env.ball_y, env.ball_vx, env.ball_vy, env.spin, env.top_paddle_x,
                              env.bottom_paddle_x, env.scoreA, env.scoreB, env.bounce_count, done))
        ho.update(oa.tobytes() + ob.tobytes() + struct.pack("<2d", ra, rb))
        if done:
            eps += 1; wa += env.scoreA > env.scoreB; wb += env.scoreB > env.scoreA
            env.reset()
    assert hs.hexdigest() == gh["state_sha256"]
    assert ho.hexdigest() == gh["obs_sha256"]
    assert (eps, wa, wb) == (gh["episodes"], gh["wins_a"], gh["wins_b"])


def test_dropin_attribute_assignment_and_unknown_kwargs(H):
    env = pp.PongEnv2P(**H["env_config_yaml"])
    env.ball_x, env.ball_y, env.ball_vx, env.ball_vy, env.spin = 0.5, 0.01, 0.0, -0.05, 0.0
    env.top_paddle_x, env.scoreB = 0.9, 2                     # ball misses the top paddle at match point
    (_, _), (ra, rb), done, _ = env.step(1, 1)
    assert (ra, rb, done, env.scoreB) == (-1.0, 1.0, True, 3)
    assert env.action_space.nvec.tolist() == [3, 3] and env.observation_space.shape == (7,)
    with pytest.raises(TypeError):
        pp.PongEnv2P(no_such_parameter=1)


def test_rollout_kernel_replays_reference_trajectory(H):
    """One env, the reference's own serves as the pool, 10k steps in ONE launch: per-step trace == reference."""
    g = dict(np.load(os.path.join(gu.GOLDEN, "env_traj_config.npz")))
    serves = g["serves"]
    pool = tuple(serves[:, i].reshape(-1, 1).copy() for i in range(3))
    env = pp.VecPongEnv2P(1, mode="f64", serve=pool, **H["env_config_yaml"])
    env.reset()
    K = g["actions"].shape[0]
    out = env.rollout(torch.from_numpy(g["actions"].reshape(K, 1, 2).copy()).cuda(), trace=True, log_cap=1024)
    assert np.array_equal(gu.bits(gu.np_of(out["trace_real"])[:, :, 0]), gu.bits(g["state"]))
    ti = gu.np_of(out["trace_int"])
    assert np.array_equal(ti[:, :3, 0], g["ints"]) and np.array_equal(ti[:, 3, 0] & 1, g["done"])
    c, gh = env.read_counters(), H["config_yaml"]
    assert (c["env_steps"], c["episodes"], c["wins_a"], c["wins_b"], c["paddle_hits"]) == \
           (K, gh["episodes"], gh["wins_a"], gh["wins_b"], gh["paddle_hits"])
    log = gu.np_of(out["ep_log"])[:env.ep_log_count()]
    lens = np.diff(np.concatenate([[-1], np.nonzero(g["done"])[0]]))
    assert np.array_equal(log[:, 3], lens) and np.array_equal(log[:, 1], np.arange(gh["episodes"]))


@pytest.mark.parametrize("mode", ["f64", "f32"])
def test_config2_4096_envs_rollout_bit_exact_vs_oracle(H, mode):
    """BASELINE config 2: 4096 lock-step envs, random actions, serve pool from the reference RNG formula.
    Every step's state (trace), the counters, the episode log and the final state equal the oracle bit for bit."""
    cfg = H["env_config_yaml"]
    n, K, depth = 4096, 600, 8
    pool = gu.make_pool(2024, n, depth, cfg, mode)
    acts = gu.random_actions(7, K, n, with_invalid=True)
    env = pp.VecPongEnv2P(n, mode=mode, serve=pool, **cfg)
    env.reset()
    b = gu.oracle_batch_like(env, mode)
    got = env.rollout(torch.from_numpy(acts).cuda(), trace=True, log_cap=1 << 17)
    want = po.rollout(po.make_params(cfg), b, acts, pool, trace=True, log_cap=1 << 17)
    assert np.array_equal(gu.bits(gu.np_of(got["trace_real"])), gu.bits(want["trace_real"]))
    assert np.array_equal(gu.np_of(got["trace_int"]), want["trace_int"])
    gu.assert_state_equal(env, b)
    assert np.array_equal(gu.np_of(env.counters), want["counters"])
    assert want["counters"][1] > 20000 and want["counters"][6] > 10000          # the run really has episodes and hits
    glog = gu.np_of(got["ep_log"])[:env.ep_log_count()]
    key = lambda a: a[np.lexsort((a[:, 1], a[:, 0]))]
    assert env.ep_log_count() == want["n_log"] and np.array_equal(key(glog), key(want["ep_log"]))


@pytest.mark.parametrize("mode", ["f64", "f32"])
def test_rollout_bit_exact_on_random_constructor_configs(H, mode):
    """The device env against the oracle on RANDOM constructor keyword sets (paddle geometry, restitution, friction incl.
    0 and > 1, mass, radius, spin on / off, speed-scaling schedule, max_score 1..5): every step's state, the counters
    and the final state are bit-exact — the kernels bake no constant of config.yaml in."""
    rs = np.random.RandomState(20260 + (mode == "f32"))
    pick = lambda xs: xs[rs.randint(len(xs))]
    for trial in range(10):
        cfg = dict(H["env_config_yaml"])
        cfg.update(paddle_width=pick([0.05, 0.2, 0.35, 0.9]), paddle_speed=pick([0.01, 0.03, 0.08]), max_score=int(rs.randint(1, 6)),
                   enable_spin=bool(rs.randint(2)), magnus_factor=pick([0.0, 0.025, 0.1]), restitution=pick([0.5, 0.9, 1.0, 1.1]),
                   friction=pick([0.0, 0.3, 0.6, 1.5]), ball_mass=pick([0.5, 1.0, 2.0]), world_ball_radius=pick([0.01, 0.03, 0.05]),
                   speed_scale_every=int(rs.randint(1, 7)), speed_increment=pick([0.0, 0.1, 0.2, 0.5]))
        n, K = 257, 300
        pool = gu.make_pool(100 + trial, n, 8, cfg, mode)
        acts = gu.random_actions(trial, K, n, with_invalid=True)
        env = pp.VecPongEnv2P(n, mode=mode, serve=pool, **cfg)
        env.reset()
        b = gu.oracle_batch_like(env, mode)
        got = env.rollout(torch.from_numpy(acts).cuda(), trace=True)
        want = po.rollout(po.make_params(cfg), b, acts, pool, trace=True)
        assert np.array_equal(gu.bits(gu.np_of(got["trace_real"])), gu.bits(want["trace_real"])), (trial, cfg)
        assert np.array_equal(gu.np_of(got["trace_int"]), want["trace_int"]), (trial, cfg)
        gu.assert_state_equal(env, b, what=str(cfg))
        assert np.array_equal(gu.np_of(env.counters), want["counters"]) and want["counters"][1] > 0


@pytest.mark.parametrize("mode", ["f64", "f32"])
@pytest.mark.parametrize("n", [1, 2, 3, 31, 513, 1024, 4099])
def test_step_kernel_ragged_sizes_and_vector_paths(H, mode, n):
    """Scalar and 16-byte vector variants of the single-step kernel, full and partial tiles."""
    cfg = H["env_config_rnn_yaml"]
    pool = gu.make_pool(n, n, 2, cfg, mode)
    env = pp.VecPongEnv2P(n, mode=mode, serve=pool, **cfg)
    oa0, ob0 = env.reset()
    b = gu.oracle_batch_like(env, mode)
    p = po.make_params(cfg)
    woa, wob = po.observe(b)
    assert np.array_equal(gu.bits(gu.np_of(oa0)), gu.bits(woa)) and np.array_equal(gu.bits(gu.np_of(ob0)), gu.bits(wob))
    acts = gu.random_actions(n + 1, 60, n)
    for t in range(60):
        (oa, ob), (ra, rb), done, _ = env.step(torch.from_numpy(acts[t, :, 0].copy()), torch.from_numpy(acts[t, :, 1].copy()))
        woa, wob, wra, wrb, wdone = po.step(p, b, acts[t, :, 0], acts[t, :, 1])
        assert np.array_equal(gu.bits(gu.np_of(oa)), gu.bits(woa)) and np.array_equal(gu.bits(gu.np_of(ob)), gu.bits(wob)), t
        assert np.array_equal(gu.np_of(ra), wra) and np.array_equal(gu.np_of(rb), wrb), t
        assert np.array_equal(gu.np_of(done), wdone), t
        if wdone.any():                                            # the callers' `if done: reset()`
            env.reset(mask=done)
            b.ep_idx[wdone] += 1
            j = b.ep_idx % 2
            idx = np.arange(n)
            b.serve(pool[0][j, idx], pool[1][j, idx], pool[2][j, idx], mask=wdone)
    real, ints = gu.read_state(env)
    sr, si = b.state_matrix()
    assert np.array_equal(gu.bits(real), gu.bits(sr)) and np.array_equal(ints[:3], si)


def test_int64_actions_and_out_of_range_values_mean_stay(H):
    cfg = H["env_config_yaml"]
    env = pp.VecPongEnv2P(4, mode="f64", **cfg)
    env.reset(serves=([0.01] * 4, [0.01] * 4, [0.0] * 4))
    env.step(torch.tensor([0, 1, 2, 300]), torch.tensor([-1, 2, 0, 7]))
    assert gu.np_of(env.top_paddle_x).tolist() == [0.5 - 0.03, 0.5, 0.5 + 0.03, 0.5]
    assert gu.np_of(env.bottom_paddle_x).tolist() == [0.5, 0.5 + 0.03, 0.5 - 0.03, 0.5]


def test_empty_batch_and_bad_arguments(H):
    lib = pp._lib.load()
    env = pp.VecPongEnv2P(8, **H["env_config_yaml"])
    import ctypes as C
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    a = torch.zeros(8, dtype=torch.uint8, device="cuda")
    ptr = lambda t: C.c_void_p(t.data_ptr())
    args = (C.byref(env.params), C.byref(env.state), ptr(a), ptr(a), ptr(env.obs_a), ptr(env.obs_b),
            ptr(env.reward_a), ptr(env.reward_b), ptr(env._done), st)
    assert lib.pp_env_step(0, 0, *args) == 0                                   # n = 0 is a no-op
    assert lib.pp_env_step(7, 8, *args) == -3                                  # PP_E_MODE
    assert lib.pp_env_step(0, -1, *args) == -2                                 # PP_E_SIZE
    bad = (C.byref(env.params), C.byref(env.state), None) + args[3:]
    assert lib.pp_env_step(0, 8, *bad) == -1                                   # PP_E_NULL
    assert b"pp_env_step" in lib.pp_last_error()
    with pytest.raises(pp.PongB200Error):
        pp._lib.check(-1, "pp_env_step")


def test_philox_serves_follow_reference_formula_and_are_shard_invariant(H):
    """Device serves: same formula as reset() (:98-111) driven by Philox; equal to the oracle's restatement up to
    the last ulps of sincos, inside the reference's ranges, and independent of how envs are split over slabs."""
    cfg = H["env_config_yaml"]
    n = 4096
    env = pp.VecPongEnv2P(n, mode="f64", serve="philox", seed=99, **cfg)
    env.reset()
    vx, vy, sp = (gu.np_of(t) for t in (env.ball_vx, env.ball_vy, env.spin))
    want = np.array([po.philox_serve(99, i, 0, cfg) for i in range(256)])
    assert np.allclose(vx[:256], want[:, 0], rtol=1e-14, atol=0) and np.allclose(vy[:256], want[:, 1], rtol=1e-14, atol=0)
    assert np.array_equal(sp[:256], want[:, 2])
    speed, ang = np.hypot(vx, vy), np.degrees(np.arctan2(vy, vx))
    assert speed.min() >= 0.03 - 1e-12 and speed.max() <= 0.05 + 1e-12 and np.all(vx > 0)
    assert np.all((np.abs(ang) >= 30 - 1e-9) & (np.abs(ang) <= 60 + 1e-9)) and 0.45 < (ang > 0).mean() < 0.55
    halves = []
    for r in range(2):
        e = pp.VecPongEnv2P(n // 2, mode="f64", serve="philox", seed=99, env_id_base=r * n // 2, **cfg)
        e.reset()
        halves.append(gu.np_of(e._real[:, :n // 2]))
    assert np.array_equal(gu.bits(np.concatenate(halves, 1)), gu.bits(gu.np_of(env._real[:, :n])))
    env.reset(mask=torch.arange(n, device="cuda") % 2 == 0)          # second serve of the even envs only
    vx2 = gu.np_of(env.ball_vx)
    assert np.all(vx2[1::2] == vx[1::2]) and np.all(vx2[0::2] != vx[0::2])
    assert gu.np_of(env.ep_idx).tolist() == [1, 0] * (n // 2)


def test_full_size_properties_1m_envs(H):
    """At the full size of BASELINE config 5 (1 M envs) the oracle is too slow for a trace; check size-independent
    properties instead: counters are consistent, and the run equals the same envs stepped as 4 independent slabs."""
    cfg = H["env_config_yaml"]
    n, K = 1 << 20, 200
    g = torch.Generator(device="cuda").manual_seed(5)
    acts = torch.randint(0, 3, (K, n, 2), dtype=torch.uint8, device="cuda", generator=g)
    env = pp.VecPongEnv2P(n, mode="f64", serve="philox", seed=3, **cfg)
    env.reset()
    env.rollout(acts, log_cap=0)
    c = env.read_counters()
    assert c["env_steps"] == n * K and c["wins_a"] + c["wins_b"] == c["episodes"]
    assert c["points_a"] + c["points_b"] >= 3 * c["episodes"]
    assert int(env.ep_idx.sum().item()) == c["episodes"]
    assert c["ep_len_sum"] + int(env.ep_len.sum().item()) == n * K
    q = n // 4
    tot = torch.zeros(8, dtype=torch.int64, device="cuda")
    for r in range(4):
        e = pp.VecPongEnv2P(q, mode="f64", serve="philox", seed=3, env_id_base=r * q, **cfg)
        e.reset()
        e.rollout(acts[:, r * q:(r + 1) * q].contiguous())
        assert torch.equal(e._real[:, :q].view(torch.int64), env._real[:, r * q:(r + 1) * q].view(torch.int64))
        tot += e.counters
    assert torch.equal(tot, env.counters)


from oracle.pong_port import EXTRA_ENV_CONFIGS as _FUZZ_CFGS


@pytest.mark.parametrize("ci", range(len(_FUZZ_CFGS)))
@pytest.mark.parametrize("mode", ["f64", "f32"])
def test_other_env_parameters_bit_exact_vs_oracle(ci, mode):
    """The env kernels take every constructor keyword of the reference, not just the two YAML blocks: defaults, spin off,
    ints from YAML, wide paddles, fast balls (multiple events per few steps), friction 0 / > 1, max_score 1..5."""
    cfg = pp.resolve_env_config(_FUZZ_CFGS[ci])
    n, K, depth = 1536, 160, 5
    pool = gu.make_pool(100 + ci, n, depth, cfg, mode)
    acts = gu.random_actions(ci, K, n, with_invalid=True)
    env = pp.VecPongEnv2P(n, mode=mode, serve=pool, **_FUZZ_CFGS[ci])
    env.reset()
    b = gu.oracle_batch_like(env, mode)
    got = env.rollout(torch.from_numpy(acts).cuda(), trace=True)
    want = po.rollout(po.make_params(cfg), b, acts, pool, trace=True)
    assert np.array_equal(gu.bits(gu.np_of(got["trace_real"])), gu.bits(want["trace_real"]))
    assert np.array_equal(gu.np_of(got["trace_int"]), want["trace_int"])
    assert np.array_equal(gu.np_of(env.counters), want["counters"])
    gu.assert_state_equal(env, b)
    # and the single-step kernel with all outputs on the same state
    a1 = gu.random_actions(77 + ci, 1, n)[0]
    (oa, ob), (ra, rb), done, _ = env.step(torch.from_numpy(a1[:, 0].copy()), torch.from_numpy(a1[:, 1].copy()))
    woa, wob, wra, wrb, wdone = po.step(po.make_params(cfg), b, a1[:, 0], a1[:, 1])
    assert np.array_equal(gu.bits(gu.np_of(oa)), gu.bits(woa)) and np.array_equal(gu.bits(gu.np_of(ob)), gu.bits(wob))
    assert np.array_equal(gu.np_of(ra), wra) and np.array_equal(gu.np_of(done), wdone)


@pytest.mark.parametrize("ci", range(5))
def test_rollout_kernel_replays_reference_trajectories_of_other_configs(ci):
    """The reference's own 3 000-step trajectories for the extra constructor keyword sets, replayed on the device."""
    g = {k.split("/", 1)[1]: v for k, v in np.load(os.path.join(gu.GOLDEN, "env_extra_cfgs.npz")).items() if k.startswith(f"c{ci}/")}
    serves = g["serves"]
    pool = tuple(serves[:, i].reshape(-1, 1).copy() for i in range(3))
    env = pp.VecPongEnv2P(1, mode="f64", serve=pool, **_FUZZ_CFGS[ci])
    env.reset()
    K = g["actions"].shape[0]
    out = env.rollout(torch.from_numpy(g["actions"].reshape(K, 1, 2).copy()).cuda(), trace=True)
    assert np.array_equal(gu.bits(gu.np_of(out["trace_real"])[:, :, 0]), gu.bits(g["state"]))
    ti = gu.np_of(out["trace_int"])
    assert np.array_equal(ti[:, :3, 0], g["ints"]) and np.array_equal(ti[:, 3, 0] & 1, g["done"])
