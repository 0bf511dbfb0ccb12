"""Phase timers of the recurrent tensor-core rollout (needs a -DPP_TC_TIMING build):
   PP_EXTRA_NVCC_FLAGS=-DPP_TC_TIMING python -m pingpong_selfplay_ai_b200.build --force; python tools/rnn_phase_timers.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pingpong_selfplay_ai_b200 as pp
tiles = int(os.environ.get("RT_TILES", "148"))          # CTAs at work (one 128-env tile each)
n, k = tiles * 128, 16
env = pp.VecPongEnv2P(n, mode="f64", serve="philox", seed=1, **dict(pp.ENV_DEFAULTS)); env.reset()
torch.manual_seed(0); a = pp.QNetRNN(); torch.manual_seed(1); b = pp.QNetRNN()
eng = pp.SelfPlayEngine(env, pp.Policy.qnetrnn(a, num_envs=n, precision="f16"), pp.Policy.qnetrnn(b, num_envs=n, precision="f16"), seed=7)
eng.run(k); torch.cuda.synchronize()
lib = pp._lib.load()
t = (C.c_ulonglong * 32)()
lib.pp_debug_rt_timing(t, 1)
eng.run(k); torch.cuda.synchronize()
lib.pp_debug_rt_timing(t, 1)
ctas, ps = tiles, tiles * k * 2                       # player-steps timed by thread 0 of every CTA
names = {1: "issuer: wait for operand rows (ready)", 3: "issuer: wait for a weight stage (full)", 4: "issuer: wait for a drained accumulator",
         8: "worker: wait for accumulators (done)", 9: "worker: h_prev staging", 10: "worker: LSTM cells (4 quarters)",
         14: "worker:   of which tcgen05.ld + wait (8 per player-step)", 15: "worker:   of which cell arithmetic (8 batches of 8 units)",
         16: "worker:   wait for L1", 17: "worker:   wait for features.2", 18: "worker:   wait for gate quarter 0",
         19: "worker:   wait for gate quarter 1", 20: "worker:   wait for gate quarter 2", 21: "worker:   wait for gate quarter 3",
         22: "worker:   wait for the shared head", 23: "worker:   wait for the dueling heads",
         12: "worker: player-step total", 13: "worker: env step + bookkeeping (per lock-step step x2)"}
for slot, name in names.items():
    print(f"{name:56s} {t[slot] / ps:10.0f} cycles / player-step")
