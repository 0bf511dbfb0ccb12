"""Fused self-play kernel, replay scatter and the host-buffer entry (`-m gpu`) against the oracle's closed loop.

The fp32 QNet path computes the oracle's exact fmaf chain, so greedy actions — and with them every trajectory,
score, done step and winner — are bit-identical to the oracle even in closed loop."""
import os

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


@pytest.fixture(scope="module")
def nets():
    qg = dict(np.load(os.path.join(gu.GOLDEN, "qnet_golden.npz")))
    return {k: gu.golden_sd(qg, k) for k in ("seed0", "seed1", "ckpt_model5_1_fault_B")}


def _oracle_policy(kind, sd=None, eps=0.0, tol=0.02, noisy=False):
    w = po.qnet_weights_from_state_dict(sd, noisy=noisy) if sd is not None else None
    return po.make_policy(kind, w, eps=eps, tol=tol)


def _sorted_rows(*cols):
    m = np.concatenate([np.ascontiguousarray(c).reshape(len(c), -1).view(np.uint8).reshape(len(c), -1) for c in cols], 1)
    return m[np.lexsort(m.T[::-1])]


CASES = {
    "qnet_vs_qnet_greedy": dict(a=("qnet", "seed0", 0.0), b=("qnet", "seed1", 0.0)),
    "ckpt_vs_follower": dict(a=("follower", None, 0.0), b=("qnet", "ckpt_model5_1_fault_B", 0.0)),
    "eps_greedy_vs_random": dict(a=("random", None, 0.0), b=("qnet", "seed1", 0.25)),
}


def _mk(spec, nets, which):
    kind, name, eps = spec
    if kind == "qnet":
        return (pp.Policy.qnet(nets[name], eps=eps), _oracle_policy(po.POLICY_QNET, nets[name], eps=eps))
    if kind == "follower":
        return pp.Policy.follower(tol=0.02), _oracle_policy(po.POLICY_FOLLOWER, tol=0.02)
    return pp.Policy.random(), _oracle_policy(po.POLICY_RANDOM)


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("mode", ["f64", "f32"])
def test_selfplay_rollout_bit_exact_vs_oracle(H, nets, case, mode):
    cfg = H["env_config_yaml"]
    n, K, depth, seed = 1000, 300, 6, 424242
    pool = gu.make_pool(11, n, depth, cfg, mode)
    env = pp.VecPongEnv2P(n, mode=mode, serve=pool, env_id_base=5000, **cfg)
    env.reset()
    b = gu.oracle_batch_like(env, mode)
    (ga, oa), (gb, ob) = _mk(CASES[case]["a"], nets, 0), _mk(CASES[case]["b"], nets, 1)
    eng = pp.SelfPlayEngine(env, ga, gb, seed=seed)
    ring = pp.ReplayRing(n * K)
    acts = []
    wacts = []
    want_c = np.zeros(8, np.int64)
    wlog, wrep = [], []
    for chunk in (100, 200):                                   # two launches: state and step index carry over
        got = eng.run(chunk, ring=ring, log_cap=1 << 15, want_actions=True)
        acts.append(gu.np_of(got["actions"]))
        w = po.selfplay(po.make_params(cfg), b, oa, ob, chunk, pool, seed=seed, step_base=eng.step_base - chunk,
                        env_id_base=5000, log_cap=1 << 15, want_actions=True, replay_cap=n * chunk)
        wacts.append(w["actions"]); want_c += w["counters"]; wlog.append(w["ep_log"]); wrep.append(w["replay"])
    assert np.array_equal(np.concatenate(acts), np.concatenate(wacts))
    gu.assert_state_equal(env, b)
    assert np.array_equal(gu.np_of(env.counters), want_c) and want_c[1] > 100
    rows = int(ring.head.item())
    assert rows == sum(len(r["act"]) for r in wrep) == n * K
    g_rows = _sorted_rows(gu.np_of(ring.obs[:rows]), gu.np_of(ring.act[:rows]), gu.np_of(ring.rew[:rows]),
                          gu.np_of(ring.next_obs[:rows]), gu.np_of(ring.done[:rows]))
    w_rows = _sorted_rows(*[np.concatenate([r[k] for r in wrep]) for k in ("obs", "act", "rew", "next", "done")])
    assert np.array_equal(g_rows, w_rows)                       # same transitions; order inside a step is free


def test_selfplay_quota_freezes_envs_and_matches_oracle(H, nets):
    """Config 3 shape (fixed episode quota per env): outcomes (scores, episode lengths, winners) bit-exact."""
    cfg = H["env_config_yaml"]
    n, quota = 2048, 3
    pool = gu.make_pool(5, n, quota, cfg, "f64")
    env = pp.VecPongEnv2P(n, mode="f64", serve=pool, **cfg)
    ga, oa = _mk(("qnet", "seed0", 0.0), nets, 0)
    gb, ob = _mk(("qnet", "seed1", 0.0), nets, 1)
    eng = pp.SelfPlayEngine(env, ga, gb, seed=1)
    res = eng.evaluate(quota, chunk=128, work_stealing=False)
    assert res["episodes"] == n * quota and int(env.ep_idx.min().item()) == quota
    b = po.EnvBatch(n, "f64")
    b.serve(pool[0][0], pool[1][0], pool[2][0])
    w = po.selfplay(po.make_params(cfg), b, oa, ob, res["lockstep_steps"], pool, seed=1, quota=quota, log_cap=n * quota)
    assert np.array_equal(gu.np_of(env.counters), w["counters"])
    gu.assert_state_equal(env, b)
    assert res["win_rate_b"] == w["counters"][3] / (n * quota)


def test_host_selfplay_eval_from_host_buffers(H, nets):
    """The HOST-buffer C-ABI entry (what a reference caller holds: numpy serves + weights) == oracle outcomes."""
    cfg = H["env_config_yaml"]
    n, quota = 1500, 2
    pool = gu.make_pool(8, n, quota, cfg, "f64")
    wa, wb = pp.pack_qnet(nets["seed1"]).numpy(), pp.pack_qnet(nets["ckpt_model5_1_fault_B"]).numpy()
    c, log = pp.host_selfplay_eval(cfg, n, quota, pool, wa, wb, chunk=64, ep_log_cap=n * quota)
    b = po.EnvBatch(n, "f64")
    b.serve(pool[0][0], pool[1][0], pool[2][0])
    oa = _oracle_policy(po.POLICY_QNET, nets["seed1"]); ob = _oracle_policy(po.POLICY_QNET, nets["ckpt_model5_1_fault_B"])
    w = po.selfplay(po.make_params(cfg), b, oa, ob, 4096, pool, quota=quota, log_cap=n * quota)
    assert c["episodes"] == n * quota
    for i, k in enumerate(pp.COUNTER_NAMES):
        if k != "env_steps":
            assert c[k] == w["counters"][i], k
    assert c["env_steps"] == w["counters"][0]
    key = lambda a: a[np.lexsort((a[:, 1], a[:, 0]))]
    assert np.array_equal(key(log[:n * quota]), key(w["ep_log"]))


def test_host_selfplay_eval_device_serves_sharded_and_vs_oracle(H, nets):
    """pool=None: serves are drawn on the device from Philox(seed; global env id, episode).  One call for n envs ==
    two calls for the two half slabs (device ordinal, env_id_base per call: the sharded form of the entry), from host
    threads at the same time; and the games are the oracle's for the same serves."""
    from concurrent.futures import ThreadPoolExecutor
    cfg = H["env_config_yaml"]
    n, quota, seed = 900, 2, 4242
    wa, wb = pp.pack_qnet(nets["seed0"]).numpy(), pp.pack_qnet(nets["seed1"]).numpy()
    key = lambda a: a[np.lexsort((a[:, 1], a[:, 0]))]
    for prec in ("f32", "f16"):
        call = lambda lo, hi: pp.host_selfplay_eval(cfg, hi - lo, quota, None, wa, wb, ep_log_cap=(hi - lo) * quota,
                                                    precision=prec, seed=seed, env_id_base=lo, device=0)
        c, log = call(0, n)
        with ThreadPoolExecutor(2) as ex:
            (c0, log0), (c1, log1) = ex.map(lambda s: call(*s), [(0, n // 3), (n // 3, n)])
        assert c["episodes"] == n * quota and all(c[k] == c0[k] + c1[k] for k in pp.COUNTER_NAMES)
        assert np.array_equal(key(log), key(np.concatenate([log0, log1])))
        if prec == "f32":
            serves = np.array([[po.philox_serve(seed, i, j, cfg) for i in range(n)] for j in range(quota)])   # [quota, n, 3]
            pool = tuple(np.ascontiguousarray(serves[:, :, k]) for k in range(3))
            b = po.EnvBatch(n, "f64")
            b.serve(pool[0][0], pool[1][0], pool[2][0])
            w = po.selfplay(po.make_params(cfg), b, _oracle_policy(po.POLICY_QNET, nets["seed0"]),
                            _oracle_policy(po.POLICY_QNET, nets["seed1"]), 4096, pool, quota=quota, log_cap=n * quota)
            assert np.array_equal(key(log), key(w["ep_log"]))
            assert all(c[k] == w["counters"][i] for i, k in enumerate(pp.COUNTER_NAMES))
    assert pp._lib.load().pp_host_release(-1) == 0
    c2, _ = pp.host_selfplay_eval(cfg, n, quota, None, wa, wb, precision="f16", seed=seed)      # buffers come back on demand
    assert c2 == c


def test_shard_invariance_philox(H, nets):
    """n envs on one slab == the same envs as two slabs with env_id_base offsets (serves, exploration and random
    players are keyed by the GLOBAL env id), so multi-GPU sharding cannot change any outcome."""
    cfg = H["env_config_rnn_yaml"]
    n, K = 4096, 256
    def run(lo, hi):
        env = pp.VecPongEnv2P(hi - lo, mode="f64", serve="philox", seed=77, env_id_base=lo, **cfg)
        env.reset()
        eng = pp.SelfPlayEngine(env, pp.Policy.random(), pp.Policy.qnet(nets["seed0"], eps=0.1), seed=9)
        eng.run(K)
        return gu.np_of(env._real[:, :hi - lo]), gu.np_of(env._int[:, :hi - lo]), gu.np_of(env.counters)
    r, i, c = run(0, n)
    r0, i0, c0 = run(0, n // 2)
    r1, i1, c1 = run(n // 2, n)
    assert np.array_equal(gu.bits(r), gu.bits(np.concatenate([r0, r1], 1)))
    assert np.array_equal(i, np.concatenate([i0, i1], 1)) and np.array_equal(c, c0 + c1)


@pytest.mark.parametrize("n,cap", [(1000, 4096), (777, 500), (64, 64)])
def test_replay_scatter_compacts_valid_rows_and_wraps(n, cap):
    rs = np.random.RandomState(n)
    obs = rs.rand(n, 7).astype(np.float32); nxt = rs.rand(n, 7).astype(np.float32)
    act = rs.randint(0, 3, n).astype(np.uint8); rew = rs.choice([-1.0, 0.0, 1.0], n).astype(np.float32)
    done = (rs.rand(n) < 0.1).astype(np.uint8); valid = (rs.rand(n) < 0.7).astype(np.uint8)
    ring = pp.ReplayRing(cap)
    ring.scatter(obs, act, rew, nxt, done, valid)
    # a batch larger than the ring keeps what sequential pushes would: among its last `cap` rows, the valid ones
    v = valid.astype(bool) & (np.arange(n) >= n - cap)
    k = int(v.sum())
    assert int(ring.head.item()) == k and k <= cap
    got = _sorted_rows(gu.np_of(ring.obs[:k]), gu.np_of(ring.act[:k]), gu.np_of(ring.rew[:k]),
                       gu.np_of(ring.next_obs[:k]), gu.np_of(ring.done[:k]))
    assert np.array_equal(got, _sorted_rows(obs[v], act[v], rew[v], nxt[v], done[v]))
    ring.scatter(obs, act, rew, nxt, done)                     # all rows valid; wraps when k + min(n, cap) > cap
    m = min(n, cap)
    assert int(ring.head.item()) == k + m and len(ring) == min(k + m, cap)
    rows = _sorted_rows(obs, act, rew, nxt, done)
    live = _sorted_rows(gu.np_of(ring.obs[:len(ring)]), gu.np_of(ring.act[:len(ring)]), gu.np_of(ring.rew[:len(ring)]),
                        gu.np_of(ring.next_obs[:len(ring)]), gu.np_of(ring.done[:len(ring)]))
    present = {r.tobytes() for r in rows}
    assert all(r.tobytes() in present for r in live)           # every slot holds one complete pushed row
    if n >= cap:                                               # the ring now holds exactly the last `cap` rows
        assert np.array_equal(live, _sorted_rows(obs[n - cap:], act[n - cap:], rew[n - cap:], nxt[n - cap:], done[n - cap:]))


def test_selfplay_ring_smaller_than_launch_keeps_last_steps(H, nets):
    """capacity < n * k: only the last capacity / n lock-step steps are written, every slot a complete row."""
    cfg = H["env_config_yaml"]
    n, K = 512, 40
    pool = gu.make_pool(3, n, 4, cfg, "f64")
    env = pp.VecPongEnv2P(n, mode="f64", serve=pool, **cfg)
    env.reset()
    b = gu.oracle_batch_like(env, "f64")
    ga, oa = _mk(("random", None, 0.0), nets, 0)
    gb, ob = _mk(("qnet", "seed0", 0.1), nets, 1)
    ring = pp.ReplayRing(n * 10 + 100)
    pp.SelfPlayEngine(env, ga, gb, seed=5).run(K, ring=ring)
    w = po.selfplay(po.make_params(cfg), b, oa, ob, K, pool, seed=5, replay_cap=n * K)["replay"]
    assert int(ring.head.item()) == n * 10
    last = {k: v[n * (K - 10):] for k, v in w.items()}
    got = _sorted_rows(gu.np_of(ring.obs[:n * 10]), gu.np_of(ring.act[:n * 10]), gu.np_of(ring.rew[:n * 10]),
                       gu.np_of(ring.next_obs[:n * 10]), gu.np_of(ring.done[:n * 10]))
    assert np.array_equal(got, _sorted_rows(last["obs"], last["act"], last["rew"], last["next"], last["done"]))
    with pytest.raises(pp.PongB200Error):
        pp.SelfPlayEngine(env, ga, gb).run(2, ring=pp.ReplayRing(n - 1))


def test_selfplay_lockstep_ring_is_time_major_per_env_and_wraps(H, nets):
    """ReplayRing(lockstep_envs=n): the row of env i at lock-step step t sits at slot (t % T) * n + i — the oracle's
    step-major replay order — over several launches and across the wrap; the host keeps the cursors."""
    cfg = H["env_config_yaml"]
    n, T = 300, 64
    pool = gu.make_pool(4, n, 8, cfg, "f64")
    env = pp.VecPongEnv2P(n, mode="f64", serve=pool, **cfg)
    env.reset()
    b = gu.oracle_batch_like(env, "f64")
    ga, oa = _mk(("qnet", "seed1", 0.0), nets, 0)
    gb, ob = _mk(("qnet", "seed0", 0.2), nets, 1)
    ring = pp.ReplayRing(n * T, lockstep_envs=n)
    eng = pp.SelfPlayEngine(env, ga, gb, seed=5)
    for k in (40, 25, 35):
        eng.run(k, ring=ring)
    K = 100
    assert ring.steps_written == K and int(ring.head.item()) == n * K
    w = po.selfplay(po.make_params(cfg), b, oa, ob, K, pool, seed=5, replay_cap=n * K)["replay"]
    for t in range(K - T, K):
        lo, slot = t * n, (t % T) * n
        assert np.array_equal(gu.np_of(ring.obs[slot:slot + n]), w["obs"][lo:lo + n]), t
        assert np.array_equal(gu.np_of(ring.next_obs[slot:slot + n]), w["next"][lo:lo + n])
        assert np.array_equal(gu.np_of(ring.act[slot:slot + n]), w["act"][lo:lo + n])
        assert np.array_equal(gu.np_of(ring.rew[slot:slot + n]), w["rew"][lo:lo + n])
        assert np.array_equal(gu.np_of(ring.done[slot:slot + n]), w["done"][lo:lo + n])
    with pytest.raises(pp.PongB200Error):                                       # the layout belongs to one slab size
        eng.run(2, ring=pp.ReplayRing(n * 2 * T, lockstep_envs=2 * n))
    with pytest.raises(pp.PongB200Error):                                       # appends have no (step, env) address
        ring.scatter(torch.zeros(4, 7), torch.zeros(4), torch.zeros(4), torch.zeros(4, 7), torch.zeros(4))


@pytest.mark.parametrize("kind", ["tc_lockstep", "tc_append", "rnn_tc_lockstep", "rnn_lockstep"])
def test_replay_rows_of_the_tensor_core_and_recurrent_kernels_match_an_oracle_replay(H, nets, kind):
    """Players on the reduced-precision paths may act differently from the oracle's, so the ring is checked against an
    ORACLE REPLAY OF THE KERNEL'S OWN ACTIONS: at every step B's observation, action, reward, terminal-or-next
    observation and done flag of every env must be the oracle's, bit for bit (scripts/train_iterative.py:243)."""
    cfg = H["env_config_yaml"]
    n, K, T = 300, 48, 64
    pool = gu.make_pool(9, n, 8, cfg, "f64")
    env = pp.VecPongEnv2P(n, mode="f64", serve=pool, **cfg)
    env.reset()
    b = gu.oracle_batch_like(env, "f64")
    if kind.startswith("tc"):
        pa, pb = pp.Policy.qnet(nets["seed1"], precision="f16"), pp.Policy.qnet(nets["seed0"], precision="f16", eps=0.3)
    else:
        prec = "f16" if kind.startswith("rnn_tc") else "f32"
        torch.manual_seed(3); ra = pp.QNetRNN()
        torch.manual_seed(4); rb = pp.QNetRNN()
        pa = pp.Policy.qnetrnn(ra, num_envs=n, precision=prec)
        pb = pp.Policy.qnetrnn(rb, num_envs=n, precision=prec, eps=0.3)
    lockstep = kind.endswith("lockstep")
    ring = pp.ReplayRing(n * T, lockstep_envs=n if lockstep else 0)
    eng = pp.SelfPlayEngine(env, pa, pb, seed=5)
    acts = np.concatenate([gu.np_of(eng.run(k, ring=ring, want_actions=True)["actions"]) for k in (20, K - 20)])
    assert int(ring.head.item()) == n * K
    p = po.make_params(cfg)
    got = {k: gu.np_of(getattr(ring, k)) for k in ("obs", "act", "rew", "next_obs", "done")}
    dones, want_rows = 0, []
    for t in range(K):
        _, ob = po.observe(b)
        tr = po.rollout(p, b, acts[t:t + 1], pool, trace=True)
        real, flags = tr["trace_real"][0], tr["trace_int"][0, 3]              # post-step, pre-reset state
        nxt = np.stack([real[0], real[1], real[2], real[3], real[6], real[5], real[4]], 1).astype(np.float32)   # _get_obs_for_B
        rew = np.where(flags & 4, 1.0, np.where(flags & 2, -1.0, 0.0)).astype(np.float32)
        done = (flags & 1).astype(np.uint8)
        dones += int(done.sum())
        rows = slice(t * n, (t + 1) * n)
        if lockstep:
            assert np.array_equal(got["obs"][rows], ob) and np.array_equal(got["next_obs"][rows], nxt), t
            assert np.array_equal(got["act"][rows], acts[t, :, 1]) and np.array_equal(got["rew"][rows], rew)
            assert np.array_equal(got["done"][rows], done)
        else:
            want_rows.append((ob, acts[t, :, 1].copy(), rew, nxt, done))
    if not lockstep:                                       # appends arrive in warp order, not step order: compare as a set
        g_rows = _sorted_rows(*[got[k][:n * K] for k in ("obs", "act", "rew", "next_obs", "done")])
        assert np.array_equal(g_rows, _sorted_rows(*[np.concatenate([w[j] for w in want_rows]) for j in range(5)]))
    assert dones > 0
    gu.assert_state_equal(env, b)


@pytest.mark.parametrize("prec,opp", [("f32", "rnn"), ("f16", "rnn"), ("f16", "qnet")])
def test_train_rnn_generation_sequence_replay_and_drqn_updates(H, prec, opp):
    """DRQN training mode on one slab (scripts/train_rnn_iterative.py:728-800): the recurrent kernel writes the lock-step
    ring, windows of 8 steps are drawn on the device, all 175 k parameters train, epsilon decays."""
    from pingpong_selfplay_ai_b200.train_rnn import DRQNTrainer, SequenceSampler, train_rnn_generation
    cfg = H["env_config_rnn_yaml"]
    n, steps, T = 1024, 96, 128
    torch.manual_seed(21); net_b = pp.QNetRNN()
    if opp == "rnn":
        torch.manual_seed(20)
        pa = pp.Policy.qnetrnn(pp.QNetRNN(), num_envs=n, precision=prec)
    else:
        torch.manual_seed(20)
        pa = pp.Policy.qnet(pp.QNet(), precision=prec)
    before = {k: v.clone() for k, v in net_b.state_dict().items()}
    env = pp.VecPongEnv2P(n, mode="f64", serve="philox", seed=6, **cfg)
    env.reset()
    trainer = DRQNTrainer(net_b, batch_size=64, target_update_interval=6, lr=1e-3)
    eng = pp.SelfPlayEngine(env, pa, pp.Policy.qnetrnn(trainer.model, num_envs=n, noisy=True, eps=1.0, precision=prec), seed=3)
    ring = pp.ReplayRing(n * T, lockstep_envs=n)
    sampler = SequenceSampler(ring, trace_length=8)
    out = train_rnn_generation(eng, trainer, ring, sampler, steps, chunk=16, updates_per_chunk=2, epsilon=1.0,
                               epsilon_decay=0.9, min_epsilon=0.05, precision=prec)
    assert out["env_steps"] == n * steps and ring.steps_written == steps and int(ring.head.item()) == n * steps
    assert out["stored_episodes"] > n // 2 and out["updates"] >= 8 and trainer.train_steps == out["updates"]
    assert out["mean_loss"] > 0 and np.isfinite(out["mean_loss"]) and 0.05 <= out["epsilon"] < 1.0
    after = trainer.model.state_dict()
    moved = [k for k in before if not k.endswith("epsilon") and not torch.equal(before[k].to(after[k].device), after[k])]
    assert {"features_extractor.0.weight", "lstm.weight_hh_l0", "fc_A.weight_mu", "fc_V.bias_sigma"} <= set(moved)
    # the ring is one time-ordered column per env: B's next observation is its observation one step later, except
    # across an episode end (terminal obs vs the new serve)
    obs = gu.np_of(ring.obs).reshape(T, n, 7)[:steps]
    nxt = gu.np_of(ring.next_obs).reshape(T, n, 7)[:steps]
    done = gu.np_of(ring.done).reshape(T, n)[:steps] != 0
    same = np.all(nxt[:-1] == obs[1:], axis=2)
    assert np.all(same[~done[:-1]]) and done.sum() == out["episodes"]
    assert not np.any(same[done[:-1]])
    rows = gu.np_of(sampler.sample_rows(512))
    t, i = rows // n, rows % n
    assert np.all(np.diff(t, axis=1) == 1) and np.all(i == i[:, :1]) and not done[t[:, :-1], i[:, :-1]].any()


# ------------------------------------------------------------------------------------------ tensor-core path
@pytest.mark.parametrize("mode", ["f64", "f32"])
@pytest.mark.parametrize("n", [1000, 5000])
def test_selfplay_tensor_core_env_bit_exact_under_its_own_actions(H, nets, mode, n):
    """PP_PREC_F16 fused kernel.  Actions come from reduced-precision Q-values, so they may differ from the oracle's
    at near-ties; everything else must not: replaying the kernel's own action stream through the oracle env gives
    the same state, counters and episode log bit for bit, and the actions agree with the fp32 oracle on the same
    observations except at near-ties."""
    cfg = H["env_config_yaml"]
    K, depth, seed = 200, 6, 99
    pool = gu.make_pool(21, n, depth, cfg, mode)
    env = pp.VecPongEnv2P(n, mode=mode, serve=pool, env_id_base=123, **cfg)
    env.reset()
    b = gu.oracle_batch_like(env, mode)
    b2 = b.copy()
    pa = pp.Policy.qnet(nets["seed0"], precision="f16")
    pb = pp.Policy.qnet(nets["ckpt_model5_1_fault_B"], eps=0.1, precision="f16")
    eng = pp.SelfPlayEngine(env, pa, pb, seed=seed)
    ring = pp.ReplayRing(n * K)
    got = eng.run(K, ring=ring, log_cap=1 << 16, want_actions=True)
    acts = gu.np_of(got["actions"])
    want = po.rollout(po.make_params(cfg), b, acts, pool, env_id_base=123, log_cap=1 << 16)
    gu.assert_state_equal(env, b)
    assert np.array_equal(gu.np_of(env.counters), want["counters"]) and want["counters"][1] > 50
    key = lambda a: a[np.lexsort((a[:, 1], a[:, 0]))]
    assert np.array_equal(key(gu.np_of(got["ep_log"])[:env.ep_log_count()]), key(want["ep_log"]))
    assert int(ring.head.item()) == n * K
    # closed-loop fp32 oracle from the same start: until the first disagreement per env the actions must be equal
    oa = _oracle_policy(po.POLICY_QNET, nets["seed0"]); ob = _oracle_policy(po.POLICY_QNET, nets["ckpt_model5_1_fault_B"], eps=0.1)
    w = po.selfplay(po.make_params(cfg), b2, oa, ob, K, pool, seed=seed, env_id_base=123, want_actions=True)
    same = (acts == w["actions"]).all(axis=2)                  # [K, n]
    first_diff = np.where(same.all(axis=0), K, np.argmin(same, axis=0))
    agree = first_diff.sum() / (K * n)
    print(f"tensor-core closed loop: {100 * agree:.2f}% of env-steps before the first action disagreement")
    assert agree >= 0.999                                       # measured 99.94 - 100 %


def test_tensor_core_rollout_head_tables_are_per_stream(H, nets):
    """The fused tensor-core rollout reads its head table from constant memory, one slot per (device, stream); streams
    beyond the 16 slots use the shared-memory table.  Twenty streams, alternating between two different player pairings,
    all in flight at once: every stream's actions and counters equal the same rollout run alone on the default stream
    (a shared table would make one pairing play with the other's heads)."""
    cfg = H["env_config_yaml"]
    n, K, depth = 2048, 96, 6
    pool = gu.make_pool(5, n, depth, cfg, "f64")
    pairs = [("seed0", "ckpt_model5_1_fault_B"), ("ckpt_model5_1_fault_B", "seed1")]

    def engine(pair):
        env = pp.VecPongEnv2P(n, mode="f64", serve=pool, env_id_base=7, **cfg)
        env.reset()
        return pp.SelfPlayEngine(env, pp.Policy.qnet(nets[pair[0]], precision="f16"), pp.Policy.qnet(nets[pair[1]], precision="f16"), seed=11)

    want = []
    for pair in pairs:
        eng = engine(pair)
        want.append((gu.np_of(eng.run(K, want_actions=True)["actions"]), gu.np_of(eng.env.counters)))
    assert not np.array_equal(want[0][0], want[1][0])
    engines = [engine(pairs[i % 2]) for i in range(20)]
    streams = [torch.cuda.Stream() for _ in engines]
    torch.cuda.synchronize()
    outs = []
    for eng, st in zip(engines, streams):
        with torch.cuda.stream(st):
            outs.append(eng.run(K, want_actions=True)["actions"])
    torch.cuda.synchronize()
    for i, (eng, out) in enumerate(zip(engines, outs)):
        assert np.array_equal(gu.np_of(out), want[i % 2][0]), f"stream {i}"
        assert np.array_equal(gu.np_of(eng.env.counters), want[i % 2][1]), f"stream {i}"


def test_tensor_core_rollout_can_be_captured_into_a_cuda_graph(H, nets):
    """The launch sequence of the fused tensor-core rollout (table kernel, device-to-device copy into constant memory,
    rollout kernel) is capturable: a replayed graph continues the rollout exactly like an eager launch (greedy players,
    Philox serves keyed by env and episode, so the baked-in step index does not matter)."""
    cfg = H["env_config_yaml"]
    n, K = 4096, 48

    def engine():
        env = pp.VecPongEnv2P(n, mode="f64", serve="philox", seed=3, **cfg)
        env.reset()
        return pp.SelfPlayEngine(env, pp.Policy.qnet(nets["seed0"], precision="f16"), pp.Policy.qnet(nets["seed1"], precision="f16"), seed=5)

    eager = engine()
    eager.run(K); eager.run(K)
    graphed = engine()
    graphed.run(K)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        graphed.run(K)
    g.replay()
    torch.cuda.synchronize()
    assert np.array_equal(gu.np_of(graphed.env.counters), gu.np_of(eager.env.counters))
    (r1, i1), (r2, i2) = gu.read_state(graphed.env), gu.read_state(eager.env)
    assert np.array_equal(gu.bits(r1), gu.bits(r2)) and np.array_equal(i1, i2)


def test_selfplay_tensor_core_mixed_players_and_win_rates(H, nets):
    """Follower / random opponents on the tensor-core path, and outcome statistics equal to the fp32 path's within
    sampling noise (same serves, greedy QNets)."""
    cfg = H["env_config_yaml"]
    n, quota = 4096, 4
    pool = gu.make_pool(6, n, quota, cfg, "f64")
    res = {}
    for prec in ("f32", "f16"):
        env = pp.VecPongEnv2P(n, mode="f64", serve=pool, **cfg)
        eng = pp.SelfPlayEngine(env, pp.Policy.qnet(nets["seed0"], precision=prec), pp.Policy.qnet(nets["seed1"], precision=prec), seed=3)
        res[prec] = eng.evaluate(quota, chunk=96)
        assert res[prec]["episodes"] == n * quota
    assert abs(res["f32"]["win_rate_b"] - res["f16"]["win_rate_b"]) <= 0.005
    assert abs(res["f32"]["env_steps"] - res["f16"]["env_steps"]) <= 0.005 * res["f32"]["env_steps"]
    for a in (pp.Policy.follower(), pp.Policy.random()):
        env = pp.VecPongEnv2P(n, mode="f64", serve=pool, **cfg)
        env.reset()
        b = gu.oracle_batch_like(env, "f64")
        got = pp.SelfPlayEngine(env, a, pp.Policy.qnet(nets["seed1"], precision="f16"), seed=4).run(50, want_actions=True)
        want = po.rollout(po.make_params(cfg), b, gu.np_of(got["actions"]), pool)
        gu.assert_state_equal(env, b)
        assert np.array_equal(gu.np_of(env.counters), want["counters"])


@pytest.mark.parametrize("prec", ["f32", "f16"])
def test_evaluate_serve_queue_equals_fixed_quota(H, nets, prec):
    """PP_SERVE_QUEUE (work stealing over the n x quota serves, one launch) gives the counters and per-episode records
    of the fixed per-env quota: each serve is played exactly once and an episode depends only on its serve."""
    cfg = H["env_config_yaml"]
    n, quota = 3000, 5
    pool = gu.make_pool(9, n, quota, cfg, "f64")
    res = {}
    for ws in (False, True):
        env = pp.VecPongEnv2P(n, mode="f64", serve=pool, **cfg)
        eng = pp.SelfPlayEngine(env, pp.Policy.qnet(nets["seed0"], precision=prec), pp.Policy.qnet(nets["seed1"], precision=prec))
        res[ws] = eng.evaluate(quota, chunk=128, work_stealing=ws, log_cap=n * quota)
        assert res[ws]["episodes"] == n * quota
    key = lambda t: gu.np_of(t)[np.lexsort((gu.np_of(t)[:, 1], gu.np_of(t)[:, 0]))]
    for k in pp.COUNTER_NAMES:
        assert res[True][k] == res[False][k], k
    assert np.array_equal(key(res[True]["ep_log"]), key(res[False]["ep_log"]))
    if prec == "f32":
        b = po.EnvBatch(n, "f64")
        b.serve(pool[0][0], pool[1][0], pool[2][0])
        w = po.selfplay(po.make_params(cfg), b, _oracle_policy(po.POLICY_QNET, nets["seed0"]),
                        _oracle_policy(po.POLICY_QNET, nets["seed1"]), 4096, pool, quota=quota, log_cap=n * quota)
        assert np.array_equal(key(res[True]["ep_log"]), w["ep_log"][np.lexsort((w["ep_log"][:, 1], w["ep_log"][:, 0]))])


# ------------------------------------------------------------------------------------------ recurrent players
def _rnn_teacher_forced_check(cfg, n, K, pol_kinds, mode="f64", quota=0, prec="f32"):
    """Fused QNetRNN self-play (config 4 shape): replay the kernel's own action stream through the oracle env (state,
    counters bit-exact) while the oracle QNetRNN, fed the same observations and carrying its own (h, c) with the
    reference's reset-at-episode-start rule, must pick the same greedy actions except at near-ties."""
    torch.manual_seed(11); net_a = pp.QNetRNN()
    torch.manual_seed(12); net_b = pp.QNetRNN()
    depth = 6
    pool = gu.make_pool(31, n, depth, cfg, mode)
    env = pp.VecPongEnv2P(n, mode=mode, serve=pool, **cfg)
    env.reset()
    b = gu.oracle_batch_like(env, mode)
    torch.manual_seed(13); net_q = pp.QNet()
    wq = po.qnet_weights_from_state_dict(net_q.state_dict())
    mk = {"qnet": lambda: pp.Policy.qnet(net_q, precision=prec),
          "rnn_a": lambda: pp.Policy.qnetrnn(net_a, num_envs=n, precision=prec),
          "rnn_b": lambda: pp.Policy.qnetrnn(net_b, num_envs=n, precision=prec),
          "follower": lambda: pp.Policy.follower(), "random": lambda: pp.Policy.random()}
    pa, pb = mk[pol_kinds[0]](), mk[pol_kinds[1]]()
    for p in (pa, pb):
        if p.h is not None:
            p.h.fill_(3.0); p.c.fill_(-2.0)                     # garbage: the first step of an episode must zero it
    eng = pp.SelfPlayEngine(env, pa, pb, seed=5)
    acts = np.concatenate([gu.np_of(eng.run(k, want_actions=True, quota=quota)["actions"]) for k in (K // 2, K - K // 2)])
    p = po.make_params(cfg)
    wts = {"rnn_a": po.qnetrnn_weights_from_state_dict(net_a.state_dict()), "rnn_b": po.qnetrnn_weights_from_state_dict(net_b.state_dict())}
    hc = {k: (np.zeros((n, 128), np.float32), np.zeros((n, 128), np.float32)) for k in wts}
    fresh = np.ones(n, bool)
    checked = agree = 0
    counters = np.zeros(8, np.int64)
    for t in range(K):
        oa, ob = po.observe(b)
        live = ~((quota > 0) & (b.ep_idx >= quota)) if quota else np.ones(n, bool)
        for side, (kind, obs) in enumerate(zip(pol_kinds, (oa, ob))):
            if kind == "qnet":                                   # fp32 fmaf chain on the CUDA cores: the oracle's bits
                _, a = po.qnet_forward(wq, obs)
                assert np.array_equal(acts[t, live, side], a[live]), (t, side)
            if kind not in wts:
                continue
            h, c = hc[kind]
            h[fresh] = 0; c[fresh] = 0
            hh, cc = h.copy(), c.copy()
            q, a = po.qnetrnn_forward(wts[kind], obs, hh, cc)
            h[live], c[live] = hh[live], cc[live]                # frozen envs do not advance
            srt = np.sort(q, axis=1)
            clear = ((srt[:, 2] - srt[:, 1]) > (1e-4 if prec == "f32" else 1e-3)) & live
            checked += clear.sum(); agree += (acts[t, clear, side] == a[clear]).sum()
        ep_before = b.ep_idx.copy()
        out = po.rollout(p, b, acts[t:t + 1], pool, quota=quota)
        counters += out["counters"]
        fresh = np.where(live, b.ep_idx != ep_before, fresh)
    gu.assert_state_equal(env, b)
    assert np.array_equal(gu.np_of(env.counters), counters) and counters[1] > 0
    assert checked > (0.9 if prec == "f32" else 0.8) * K * n * sum(k in wts for k in pol_kinds) * (0.5 if quota else 1.0)
    assert agree >= checked - max(2, checked // 2000), (agree, checked)      # fp32 exp/tanh differ in the last ulps only
    for kind, pol in (("rnn_a", pa), ("rnn_b", pb)):
        if pol.h is not None and kind in pol_kinds and not quota:
            assert np.abs(gu.np_of(pol.hidden()[0]) - hc[kind][0]).max() < (1e-4 if prec == "f32" else 1e-3)


@pytest.mark.parametrize("kinds", [("rnn_a", "rnn_b"), ("follower", "rnn_b"), ("rnn_a", "random")])
def test_selfplay_qnetrnn_rollout_teacher_forced(H, kinds):
    _rnn_teacher_forced_check(H["env_config_rnn_yaml"], 200, 70, kinds)


@pytest.mark.parametrize("prec", ["f32", "f16"])
@pytest.mark.parametrize("kinds", [("qnet", "rnn_b"), ("rnn_a", "qnet")])
def test_selfplay_qnet_meets_qnetrnn(H, kinds, prec):
    """tests/arena.py pairs QNet and QNetRNN agents: the QNet side runs in fp32 inside the recurrent kernel (both the
    CUDA-core and the tensor-core one) and picks exactly the oracle's actions."""
    _rnn_teacher_forced_check(H["env_config_yaml"], 230, 50, kinds, prec=prec)


def test_selfplay_qnetrnn_quota_and_f32_mode(H):
    _rnn_teacher_forced_check(H["env_config_rnn_yaml"], 130, 90, ("rnn_a", "rnn_b"), mode="f32", quota=2)


# ------------------------------------------------------------------------------------------ training mode
@pytest.mark.parametrize("prec", ["f32", "f16"])
def test_train_generation_rollout_replay_and_updates(H, prec):
    """Config 5 shape on one slab: epsilon-greedy rollout of the learning player B (train-mode NoisyNet weights) writes
    replay rows on the device, prioritised batches train the heads only, the target net syncs, epsilon decays."""
    cfg = H["env_config_yaml"]
    n, steps = 2048, 96
    torch.manual_seed(0); net_a = pp.QNet()
    torch.manual_seed(1); net_b = pp.QNet()
    before = {k: v.clone() for k, v in net_b.state_dict().items()}
    env = pp.VecPongEnv2P(n, mode="f64", serve="philox", seed=5, **cfg)
    env.reset()
    trainer = pp.DQNTrainer(net_b, batch_size=256, target_update_interval=8, lr=1e-3)
    eng = pp.SelfPlayEngine(env, pp.Policy.qnet(net_a, noisy=True, precision=prec),
                            pp.Policy.qnet(net_b, noisy=True, eps=1.0, precision=prec), seed=3)
    ring = pp.ReplayRing(n * 64)
    sampler = pp.PrioritizedSampler(ring)
    out = pp.train_generation(eng, trainer, ring, sampler, steps, chunk=16, updates_per_chunk=3, epsilon=1.0,
                              epsilon_decay=0.9, min_epsilon=0.02, precision=prec)
    assert out["env_steps"] == n * steps and int(ring.head.item()) == n * steps and len(sampler) == ring.capacity
    assert out["updates"] == 6 * 3 and trainer.train_steps == 18 and out["mean_loss"] > 0
    assert 0.02 <= out["epsilon"] < 1.0 and out["episodes"] > n
    after = trainer.model.state_dict()
    for k in before:
        same = torch.equal(before[k].to(after[k].device), after[k])
        assert same == (k.startswith("features") or k.endswith("epsilon") and False), k     # heads (and their noise) move, features do not
    for p, q in zip(trainer.model.parameters(), trainer.target.parameters()):
        assert torch.allclose(p, q, atol=5e-3)                                               # synced at update 16, two Adam steps ago
    # the ring holds B's view: rewards in {-1, 0, +1}, actions in {0, 1, 2}, done only with a non-zero reward
    rew, done, act = gu.np_of(ring.rew), gu.np_of(ring.done), gu.np_of(ring.act)
    assert set(np.unique(rew)) <= {-1.0, 0.0, 1.0} and act.max() <= 2 and np.all(rew[done != 0] != 0)
    # epsilon = 1.0 at the start: B's first actions are uniform over {0, 1, 2}
    assert 0.25 < (act == 1).mean() < 0.45


@pytest.mark.parametrize("kinds", [("rnn_a", "rnn_b"), ("follower", "rnn_b"), ("rnn_a", "random")])
def test_selfplay_qnetrnn_tensor_core_rollout_teacher_forced(H, kinds):
    """The tensor-core QNetRNN fused rollout (PP_PREC_F16): same checks as the fp32 kernel, several tiles, ragged."""
    _rnn_teacher_forced_check(H["env_config_rnn_yaml"], 300, 60, kinds, prec="f16")


def test_selfplay_qnetrnn_tensor_core_quota_and_f32_mode(H):
    _rnn_teacher_forced_check(H["env_config_rnn_yaml"], 130, 90, ("rnn_a", "rnn_b"), mode="f32", quota=2, prec="f16")
