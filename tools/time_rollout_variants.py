"""Cost of replay rows and of epsilon-greedy exploration inside the fused tensor-core rollout."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pingpong_selfplay_ai_b200 as pp
n, k = 65536, 64
cfg = dict(pp.ENV_DEFAULTS)
torch.manual_seed(0); a = pp.QNet(); torch.manual_seed(1); b = pp.QNet()
def run(name, eps, ring, lockstep=False):
    env = pp.VecPongEnv2P(n, mode="f64", serve="philox", seed=1, **cfg); env.reset()
    eng = pp.SelfPlayEngine(env, pp.Policy.qnet(a, precision="f16"), pp.Policy.qnet(b, precision="f16", eps=eps), seed=7)
    r = pp.ReplayRing(1 << 22, lockstep_envs=n if lockstep else 0) if ring else None
    for _ in range(3): eng.run(k, ring=r)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): eng.run(k, ring=r)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:36s} {e0.elapsed_time(e1) / 30:7.3f} ms")
run("greedy, no ring", 0.0, False)
run("eps 0.5, no ring", 0.5, False)
run("greedy, append ring", 0.0, True)
run("eps 0.5, append ring", 0.5, True)
run("eps 0.5, lock-step ring", 0.5, True, True)
