"""Parity AT THE SHAPES THE NUMBERS ARE QUOTED ON (`-m gpu`).

The fused kernels pick their work decomposition from n: above 148 x 4 x 32 = 18 944 envs the tensor-core kernel gives a
group more than one warp and the launcher cuts the batch into whole rounds of chunks over the SMs; the recurrent
tensor-core kernel walks several 128-env tiles per CTA; the host-buffer entry drains a serve queue across all of it.
None of that is exercised by the small-n tests, so here the same oracle comparisons run at BASELINE.json's sizes
(65 536 envs for configs[2], 32 768 envs per GPU for configs[3]) and at a ragged size just above the switch (19 001).
The oracle's env slabs run on host threads (oracle.pong_oracle.*_parallel).
"""
import numpy as np
import pytest
import torch

import pingpong_selfplay_ai_b200 as pp
from oracle import pong_oracle as po
import pp_testutil as gu

pytestmark = pytest.mark.gpu

BENCH_ENV_CFG = dict(  # bench.py's ENV_CFG = the env block of the reference's config.yaml
    render_size=400, paddle_width=0.2, paddle_speed=0.03, max_score=3, enable_render=False, enable_spin=True,
    magnus_factor=0.025, restitution=1, friction=0.6, ball_mass=1.0, world_ball_radius=0.03,
    ball_speed_range=[0.03, 0.05], spin_range=[-5, 5], ball_angle_intervals=[[-60, -30], [30, 60]],
    speed_scale_every=1, speed_increment=0.1)


def _key(a):
    a = np.asarray(a)
    return a[np.lexsort((a[:, 1], a[:, 0]))]


def _bench_nets():
    torch.manual_seed(0); a = pp.QNet()
    torch.manual_seed(1); b = pp.QNet()
    return a, b


def _device_philox_pool(n, depth, seed, env_id_base, cfg, mode="f64"):
    """The Philox serves (seed; env_id_base + i, episode j) for j < depth AS THE DEVICE DRAWS THEM, read back through
    pp_env_reset: the serve pool an oracle replay of a Philox-served launch needs."""
    env = pp.VecPongEnv2P(n, mode=mode, serve="philox", seed=seed, env_id_base=env_id_base, **cfg)
    rows = []
    for j in range(depth):
        env.ep_idx.fill_(j)
        env._served_once = False                                    # reset(): serve ep_idx without advancing it
        env.reset()
        rows.append([gu.np_of(env.ball_vx).copy(), gu.np_of(env.ball_vy).copy(), gu.np_of(env.spin).copy()])
    return tuple(np.stack([r[k] for r in rows]) for k in range(3))


@pytest.mark.parametrize("n", [65536, 19001])
def test_fp32_fused_rollout_at_headline_shape_bit_exact_vs_oracle_closed_loop(n):
    """configs[2] on the bit-identical fp32 path: 65 536 (and 19 001) envs, QNet A vs QNet B, closed loop — actions,
    state, counters and the episode log equal the oracle's (envs/my_pong_env_2p.py:116-225, models/qnet.py:71-75)."""
    cfg, K, depth, seed = BENCH_ENV_CFG, 32, 4, 77
    na, nb = _bench_nets()
    pool = gu.make_pool(17, n, depth, cfg, "f64")
    env = pp.VecPongEnv2P(n, mode="f64", serve=pool, env_id_base=1 << 20, **cfg)
    env.reset()
    b = gu.oracle_batch_like(env, "f64")
    eng = pp.SelfPlayEngine(env, pp.Policy.qnet(na), pp.Policy.qnet(nb, eps=0.05), seed=seed)
    got = eng.run(K, log_cap=n, want_actions=True)
    oa = po.make_policy(po.POLICY_QNET, po.qnet_weights_from_state_dict(na.state_dict()))
    ob = po.make_policy(po.POLICY_QNET, po.qnet_weights_from_state_dict(nb.state_dict()), eps=0.05)
    w = po.selfplay_parallel(po.make_params(cfg), b, oa, ob, K, pool, seed=seed, env_id_base=1 << 20, log_cap=n,
                             want_actions=True)
    assert np.array_equal(gu.np_of(got["actions"]), w["actions"])
    gu.assert_state_equal(env, b)
    assert np.array_equal(gu.np_of(env.counters), w["counters"]) and w["counters"][1] > n // 4
    assert np.array_equal(_key(gu.np_of(got["ep_log"])[:env.ep_log_count()]), _key(w["ep_log"]))


@pytest.mark.parametrize("n,serve", [(65536, "philox"), (65536, "pool"), (19001, "philox"), (19001, "pool"), (75776, "philox")])
def test_tensor_core_fused_rollout_at_headline_shape_env_bit_exact(n, serve):
    """THE code path of the headline number (bench.py: 65 536 envs, tcgen05 QNet, Philox serves prefetched into shared
    memory, several warps per group, whole rounds of chunks): the kernel's own action stream replayed through the
    oracle env gives the same state, counters and episode log bit for bit — every env, every warp slot of every group.
    The actions themselves agree with the fp32 oracle's closed loop on >= 99.9 % of env-steps (near-ties aside)."""
    cfg, K, depth, seed, base = BENCH_ENV_CFG, 96, 12, 2026, 3 * n
    na, nb = _bench_nets()
    if serve == "philox":
        pool = _device_philox_pool(n, depth, seed, base, cfg)
        env = pp.VecPongEnv2P(n, mode="f64", serve="philox", seed=seed, env_id_base=base, **cfg)
    else:
        pool = gu.make_pool(23, n, depth, cfg, "f64")
        env = pp.VecPongEnv2P(n, mode="f64", serve=pool, env_id_base=base, **cfg)
    env.reset()
    b = gu.oracle_batch_like(env, "f64")
    b2 = b.copy()
    assert np.array_equal(gu.bits(b.vx), gu.bits(pool[0][0]))           # episode 0 is pool row 0 in both modes
    eng = pp.SelfPlayEngine(env, pp.Policy.qnet(na, precision="f16"), pp.Policy.qnet(nb, precision="f16"), seed=7)
    acts = []
    for k in (64, K - 64):                                              # two launches: registers -> HBM -> registers
        acts.append(gu.np_of(eng.run(k, log_cap=0, want_actions=True)["actions"]))
    acts = np.concatenate(acts)
    p = po.make_params(cfg)
    want = po.rollout_parallel(p, b, acts, pool, env_id_base=base, log_cap=4 * n)
    assert int(b.ep_idx.max()) < depth                                  # the replay never wrapped the pool
    gu.assert_state_equal(env, b)
    assert np.array_equal(gu.np_of(env.counters), want["counters"]) and want["counters"][0] == n * K
    assert want["counters"][1] > n
    oa = po.make_policy(po.POLICY_QNET, po.qnet_weights_from_state_dict(na.state_dict()))
    ob = po.make_policy(po.POLICY_QNET, po.qnet_weights_from_state_dict(nb.state_dict()))
    w = po.selfplay_parallel(p, b2, oa, ob, 32, pool, seed=7, env_id_base=base, want_actions=True)
    same = (acts[:32] == w["actions"]).all(axis=2)
    first_diff = np.where(same.all(axis=0), 32, np.argmin(same, axis=0))
    agree = first_diff.sum() / (32 * n)
    print(f"n={n} {serve}: {100 * agree:.3f}% of env-steps before the first disagreement with the fp32 oracle")
    assert agree >= 0.999


def test_tensor_core_fused_rollout_at_headline_shape_episode_log_and_replay_rows():
    """Same shape with the episode log and the replay ring on (training-mode stores from every warp of every group):
    log rows and B's transitions equal an oracle replay of the kernel's actions."""
    cfg, n, K, depth = BENCH_ENV_CFG, 65536, 24, 4
    na, nb = _bench_nets()
    pool = gu.make_pool(29, n, depth, cfg, "f64")
    env = pp.VecPongEnv2P(n, mode="f64", serve=pool, **cfg)
    env.reset()
    b = gu.oracle_batch_like(env, "f64")
    eng = pp.SelfPlayEngine(env, pp.Policy.qnet(na, precision="f16"), pp.Policy.qnet(nb, precision="f16", eps=0.3), seed=11)
    ring = pp.ReplayRing(n * K, lockstep_envs=n)
    got = eng.run(K, ring=ring, log_cap=2 * n, want_actions=True)
    acts = gu.np_of(got["actions"])
    p = po.make_params(cfg)
    obs_rows, nxt_rows = gu.np_of(ring.obs).reshape(K, n, 7), gu.np_of(ring.next_obs).reshape(K, n, 7)
    act_rows, rew_rows, done_rows = (gu.np_of(t).reshape(K, n) for t in (ring.act, ring.rew, ring.done))
    logs = []
    for t in range(K):
        _, ob = po.observe(b)
        sa0, sb0 = b.sa.copy(), b.sb.copy()
        r = po.rollout_parallel(p, b, acts[t:t + 1], pool, log_cap=n)
        logs.append(r["ep_log"])
        assert np.array_equal(gu.bits(obs_rows[t]), gu.bits(ob)), t
        assert np.array_equal(act_rows[t], acts[t, :, 1])
        fin = b.ep_len == 0                                             # auto-reset happened: the episode ended this step
        pts_b = np.where(fin, 0, b.sb - sb0); pts_a = np.where(fin, 0, b.sa - sa0)
        live = ~fin
        assert np.array_equal(rew_rows[t][live], (pts_b - pts_a)[live].astype(np.float32))
        assert np.array_equal(done_rows[t] != 0, fin)
        assert np.all(np.abs(rew_rows[t][fin]) == 1.0)
        _, ob_next = po.observe(b)
        assert np.array_equal(gu.bits(nxt_rows[t][live]), gu.bits(ob_next[live]))   # terminal rows hold the pre-reset obs
    gu.assert_state_equal(env, b)
    assert np.array_equal(_key(gu.np_of(got["ep_log"])[:env.ep_log_count()]), _key(np.concatenate(logs)))
    assert int(ring.head.item()) == n * K


def test_host_selfplay_eval_at_bench_shape_vs_oracle():
    """pp_host_selfplay_eval at the bench's e2e shape (65 536 envs x 8 episodes from host buffers).  fp32 precision:
    counters and all 524 288 episode records equal the oracle's closed loop.  f16 (the timed configuration): every
    serve is played exactly once and the outcome statistics match fp32's within 0.5 %."""
    cfg, n, quota = BENCH_ENV_CFG, 65536, 8
    na, nb = _bench_nets()
    pool = gu.make_pool(31, n, quota, cfg, "f64")
    wa, wb = pp.pack_qnet(na).numpy(), pp.pack_qnet(nb).numpy()
    c32, log32 = pp.host_selfplay_eval(cfg, n, quota, pool, wa, wb, ep_log_cap=n * quota, precision="f32")
    b = po.EnvBatch(n, "f64")
    b.serve(pool[0][0], pool[1][0], pool[2][0])
    oa = po.make_policy(po.POLICY_QNET, po.qnet_weights_from_state_dict(na.state_dict()))
    ob = po.make_policy(po.POLICY_QNET, po.qnet_weights_from_state_dict(nb.state_dict()))
    w = po.selfplay_parallel(po.make_params(cfg), b, oa, ob, 4096, pool, quota=quota, log_cap=n * quota)
    assert w["counters"][1] == n * quota
    for i, k in enumerate(pp.COUNTER_NAMES):
        assert c32[k] == w["counters"][i], k
    assert np.array_equal(_key(log32), _key(w["ep_log"]))
    c16, log16 = pp.host_selfplay_eval(cfg, n, quota, pool, wa, wb, ep_log_cap=n * quota, precision="f16")
    assert c16["episodes"] == n * quota
    served = log16[:, 0].astype(np.int64) * quota + log16[:, 1]
    assert np.array_equal(np.sort(served), np.arange(n * quota))        # each (env, episode) serve exactly once
    same = (_key(log16) == _key(log32)).all(axis=1).mean()
    print(f"host eval f16 vs f32: {100 * same:.3f}% of the 524288 episode records identical")
    assert same > 0.995
    assert abs(c16["wins_b"] - c32["wins_b"]) <= 0.005 * n * quota
    assert abs(c16["env_steps"] - c32["env_steps"]) <= 0.005 * c32["env_steps"]


def test_recurrent_tensor_core_rollout_at_config4_shape():
    """configs[3] per-GPU shape (262 144 envs / 8 GPUs = 32 768): selfplay_rnn_tc_kernel walks 256 tiles over 148 SMs.
    The kernel's action stream replayed through the oracle env is bit-exact for all envs; the oracle QNetRNN
    (models/qnet_rnn.py:107-144), teacher-forced on a subset of envs spread over every tile, picks the same greedy
    actions except at near-ties and ends with the same (h, c) within 1e-3."""
    cfg = dict(BENCH_ENV_CFG, speed_scale_every=5, speed_increment=0.2)      # config_rnn.yaml:27-28
    n, K, depth = 32768, 20, 4
    torch.manual_seed(0); net_a = pp.QNetRNN()
    torch.manual_seed(1); net_b = pp.QNetRNN()
    pool = gu.make_pool(37, n, depth, cfg, "f64")
    env = pp.VecPongEnv2P(n, mode="f64", serve=pool, **cfg)
    env.reset()
    b = gu.oracle_batch_like(env, "f64")
    pa = pp.Policy.qnetrnn(net_a, num_envs=n, precision="f16")
    pb = pp.Policy.qnetrnn(net_b, num_envs=n, precision="f16")
    for pol in (pa, pb):
        pol.h.fill_(3.0); pol.c.fill_(-2.0)                             # garbage: the first step of an episode zeroes it
    eng = pp.SelfPlayEngine(env, pa, pb, seed=5)
    acts = np.concatenate([gu.np_of(eng.run(k, want_actions=True)["actions"]) for k in (12, K - 12)])
    sub = np.unique(np.concatenate([np.arange(0, n, 13), np.arange(n - 300, n), np.arange(128 * 147, 128 * 149)]))
    wts = [po.qnetrnn_weights_from_state_dict(m.state_dict()) for m in (net_a, net_b)]
    hc = [(np.zeros((len(sub), 128), np.float32), np.zeros((len(sub), 128), np.float32)) for _ in range(2)]
    p = po.make_params(cfg)
    fresh = np.ones(n, bool)
    counters = np.zeros(8, np.int64)
    checked = agree = 0
    for t in range(K):
        obs = po.observe(b)
        for side in range(2):
            h, c = hc[side]
            h[fresh[sub]] = 0; c[fresh[sub]] = 0
            q, a = po.qnetrnn_forward_parallel(wts[side], obs[side][sub], h, c)
            srt = np.sort(q, axis=1)
            clear = (srt[:, 2] - srt[:, 1]) > 1e-3
            checked += int(clear.sum()); agree += int((acts[t, sub, side][clear] == a[clear]).sum())
        ep_before = b.ep_idx.copy()
        counters += po.rollout_parallel(p, b, acts[t:t + 1], pool)["counters"]
        fresh = b.ep_idx != ep_before
    gu.assert_state_equal(env, b)
    assert np.array_equal(gu.np_of(env.counters), counters) and counters[0] == n * K and counters[1] > 0
    assert checked > 0.8 * 2 * K * len(sub) and agree >= checked - max(2, checked // 2000), (agree, checked)
    keep = ~fresh[sub]                                                  # an env that just finished is zeroed at its next step
    for side, pol in enumerate((pa, pb)):
        assert np.abs(gu.np_of(pol.hidden()[0])[sub][keep] - hc[side][0][keep]).max() < 1e-3
        assert np.abs(gu.np_of(pol.hidden()[1])[sub][keep] - hc[side][1][keep]).max() < 1e-3


_PAIR_SCRIPT = r"""
import hashlib, sys
import numpy as np, torch
sys.path.insert(0, sys.argv[1])
import pingpong_selfplay_ai_b200 as pp
cfg = eval(sys.argv[2])
n = int(sys.argv[3])
torch.manual_seed(0); net_a = pp.QNetRNN()
torch.manual_seed(1); net_b = pp.QNetRNN()
env = pp.VecPongEnv2P(n, mode="f64", serve="philox", seed=11, **cfg)
env.reset()
pa = pp.Policy.qnetrnn(net_a, num_envs=n, precision="f16")
pb = pp.Policy.qnetrnn(net_b, num_envs=n, precision="f16")
eng = pp.SelfPlayEngine(env, pa, pb, seed=5)
h = hashlib.sha256()
for k in (7, 18):
    h.update(eng.run(k, want_actions=True)["actions"].cpu().numpy().tobytes())
for t in (env._real[:, :n].contiguous(), env._int[:, :n].contiguous(), env.counters, pa.hidden()[0], pa.hidden()[1], pb.hidden()[0], pb.hidden()[1]):
    h.update(t.cpu().numpy().tobytes())
print("DIGEST", h.hexdigest())
"""


def test_recurrent_rollout_with_a_shared_weight_stream_equals_the_unpaired_kernel():
    """Above 148 x 4 warps of envs the recurrent kernel runs as clusters of two CTAs that multicast half of every weight
    stage to each other (lstm_tc_kernels.cu, PAIR); PP_RNN_PAIR=0 selects the one-CTA form.  Same arithmetic, so the action
    stream, the env state and both players' (h, c) are bit-identical — and a ragged n (whole rounds of unequal chunks)
    must not deadlock the pair."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    digests = {}
    for pair in ("1", "0"):
        env = dict(os.environ, PP_RNN_PAIR=pair)
        r = subprocess.run([sys.executable, "-c", _PAIR_SCRIPT, root, repr(BENCH_ENV_CFG), "40001"], env=env, capture_output=True,
                           text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        digests[pair] = [ln for ln in r.stdout.splitlines() if ln.startswith("DIGEST")][0]
    assert digests["1"] == digests["0"]
