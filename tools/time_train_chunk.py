"""Where a training-mode chunk spends its time (config 5 shape on one GPU): python tools/time_train_chunk.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pingpong_selfplay_ai_b200 as pp

n, k, prec = 65536, 64, "f16"
cfg = dict(pp.ENV_DEFAULTS)
torch.manual_seed(0); net_a = pp.QNet()
torch.manual_seed(1); net_b = pp.QNet()
env = pp.VecPongEnv2P(n, mode="f64", serve="philox", seed=1, **cfg); env.reset()
tr = pp.DQNTrainer(net_b, batch_size=256)
eng = pp.SelfPlayEngine(env, pp.Policy.qnet(net_a, noisy=True, precision=prec), pp.Policy.qnet(net_b, noisy=True, eps=0.5, precision=prec), seed=7)
ring = pp.ReplayRing(1 << 22); sampler = pp.PrioritizedSampler(ring)
pp.train_generation(eng, tr, ring, sampler, k * 4, chunk=k, updates_per_chunk=4, epsilon=0.5, precision=prec)
torch.cuda.synchronize()
def timed(name, fn, reps=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); print(f"{name:28s} {1e3 * (time.perf_counter() - t0) / reps:8.3f} ms")
timed("rollout k=64 + ring", lambda: eng.run(k, ring=ring))
timed("reset_noise + pack (fused)", lambda: tr.reset_noise_and_pack(eng.pb.weights))
timed("note_new_rows", lambda: sampler.note_new_rows(n * k))
timed("update (graph)", lambda: tr.update(sampler))
timed("sampler.sample only", lambda: sampler.sample(256, 0.5))
timed("counters .item()", lambda: int(env.counters[1].item()))
timed("whole chunk", lambda: pp.train_generation(eng, tr, ring, sampler, k, chunk=k, updates_per_chunk=4, epsilon=0.5, precision=prec), reps=10)
