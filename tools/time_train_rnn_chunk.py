"""Where a DRQN training chunk spends its time (CUDA events + host wall per phase): python tools/time_train_rnn_chunk.py [envs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pingpong_selfplay_ai_b200 as pp
from pingpong_selfplay_ai_b200.train_rnn import DRQNTrainer, SequenceSampler
from pingpong_selfplay_ai_b200.policy import pack_qnetrnn_tc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
k, dev = 64, torch.device("cuda")
cfg = dict(pp.ENV_DEFAULTS)
env = pp.VecPongEnv2P(n, mode="f64", serve="philox", seed=1, **cfg); env.reset()
torch.manual_seed(0); a = pp.QNetRNN(); torch.manual_seed(1); b = pp.QNetRNN()
trainer = DRQNTrainer(b, batch_size=64, device=dev)
eng = pp.SelfPlayEngine(env, pp.Policy.qnetrnn(a, num_envs=n, precision="f16"),
                        pp.Policy.qnetrnn(trainer.model, num_envs=n, noisy=True, eps=0.5, precision="f16"), seed=7)
ring = pp.ReplayRing(n * k, lockstep_envs=n)
sampler = SequenceSampler(ring, trace_length=8)


def phase(name, fn, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    host = time.perf_counter() - t0
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    print(f"{name:34s} device {e0.elapsed_time(e1) / reps:8.3f} ms   host-issue {1e3 * host / reps:8.3f} ms   wall {1e3 * wall / reps:8.3f} ms", flush=True)
    return out


def pack():
    trainer.model.reset_noise()
    blob = pack_qnetrnn_tc(trainer.model, noisy=True).to(dev)
    eng.pb.weights.copy_(blob, non_blocking=True)


for _ in range(2):
    pack(); eng.run(k, ring=ring); sampler.refresh()
    for _ in range(4):
        trainer.update(sampler)
phase("reset_noise + pack_qnetrnn_tc (host)", pack)
phase("reset_noise + pack (device kernels)", lambda: trainer.reset_noise_and_pack_tc(eng.pb.weights))
phase(f"rollout {k} steps x {n} envs + ring", lambda: eng.run(k, ring=ring))
phase("sampler.refresh", sampler.refresh)
phase("trainer.update (graph replay)", lambda: trainer.update(sampler), reps=20)
phase("counters .item()", lambda: int(env.counters[1].item()))
