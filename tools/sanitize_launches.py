"""One small launch of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_launches.py
Sizes are tiny (the tools slow kernels down 10-100x) but ragged, so tail warps, partial tiles and the quota path run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pingpong_selfplay_ai_b200 as pp

cfg = dict(pp.ENV_DEFAULTS)
which = sys.argv[1:] or ["tc", "rnn_tc", "fp32", "rnn_fp32", "train", "env"]
torch.manual_seed(0); qa = pp.QNet(); torch.manual_seed(1); qb = pp.QNet()
torch.manual_seed(2); ra = pp.QNetRNN(); torch.manual_seed(3); rb = pp.QNetRNN()


def fresh(n):
    env = pp.VecPongEnv2P(n, mode="f64", serve="philox", seed=1, **cfg)
    env.reset()
    return env


if "tc" in which:            # fused tensor-core self-play: several groups, ragged tail, replay ring, episode log, quota
    env = fresh(700)
    eng = pp.SelfPlayEngine(env, pp.Policy.qnet(qa, precision="f16"), pp.Policy.qnet(qb, precision="f16", eps=0.2), seed=3)
    eng.run(24, ring=pp.ReplayRing(700 * 24), log_cap=4096, want_actions=True)
    eng.run(40, quota=2)
    pp.qnet_act(torch.rand(300, 7, device="cuda"), pp.Policy.qnet(qa, precision="f16"), want_q=True)
    print("tc ok", env.read_counters())
if "rnn_tc" in which:        # fused recurrent tensor-core self-play: 3 tiles, ragged, TMA weight ring, ring rows
    env = fresh(300)
    eng = pp.SelfPlayEngine(env, pp.Policy.qnetrnn(ra, num_envs=300, precision="f16"),
                            pp.Policy.qnetrnn(rb, num_envs=300, precision="f16", eps=0.2), seed=3)
    eng.run(6, ring=pp.ReplayRing(300 * 8, lockstep_envs=300), want_actions=True)
    pp.qnetrnn_act(torch.rand(200, 7, device="cuda"), pp.Policy.qnetrnn(ra, num_envs=200, precision="f16"), want_q=True)
    print("rnn_tc ok", env.read_counters())
if "fp32" in which:
    env = fresh(1000)
    pp.SelfPlayEngine(env, pp.Policy.qnet(qa), pp.Policy.follower(), seed=3).run(20, ring=pp.ReplayRing(1000 * 20))
    print("fp32 ok", env.read_counters())
if "rnn_fp32" in which:
    env = fresh(200)
    pp.SelfPlayEngine(env, pp.Policy.qnetrnn(ra, num_envs=200), pp.Policy.qnet(qb), seed=3).run(5)
    print("rnn_fp32 ok", env.read_counters())
if "train" in which:
    env = fresh(1024)
    trainer = pp.DQNTrainer(qb, batch_size=256, use_graph=False)
    eng = pp.SelfPlayEngine(env, pp.Policy.qnet(qa, noisy=True), pp.Policy.qnet(qb, noisy=True, eps=1.0), seed=3)
    ring = pp.ReplayRing(1 << 15)
    out = pp.train_generation(eng, trainer, ring, pp.PrioritizedSampler(ring), 32, chunk=16, updates_per_chunk=2)
    print("train ok", out["updates"], out["mean_loss"])
if "env" in which:
    env = fresh(4099)
    a = torch.randint(0, 3, (4099,), dtype=torch.uint8, device="cuda")
    env.step(a, a)
    env.rollout(torch.randint(0, 3, (30, 4099, 2), dtype=torch.uint8, device="cuda"), log_cap=1024)
    pp.collide_batch([-0.04] * 5, [0.02] * 5, [0.03] * 5, [3.0] * 5, 1, 0.6, 1.0, 0.03)
    print("env ok", env.read_counters())
torch.cuda.synchronize()
