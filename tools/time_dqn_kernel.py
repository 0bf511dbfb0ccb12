"""CUDA-event timing of the training-mode kernels alone: python tools/time_dqn_kernel.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pingpong_selfplay_ai_b200 as pp
from pingpong_selfplay_ai_b200 import _lib
from pingpong_selfplay_ai_b200.selfplay import _ptr, _stream_ptr

torch.manual_seed(0)
tr = pp.DQNTrainer(pp.QNet(), batch_size=256, fused=True, use_graph=False)
ring = pp.ReplayRing(1 << 22)
ring.obs.uniform_(-1, 1); ring.next_obs.uniform_(-1, 1); ring.head.fill_(1 << 22)
sampler = pp.PrioritizedSampler(ring); sampler.note_new_rows()
idx, iw = sampler.sample(256, 0.5)
lib, st = _lib.load(), _stream_ptr(torch.device("cuda"))
rs = ring.struct()
def grads():
    lib.pp_dqn_head_grads(C.byref(rs), _ptr(idx), _ptr(iw), 256, *tr._feature_ptrs(), C.byref(tr._on_v), C.byref(tr._on_a),
                          C.byref(tr._tg_v), C.byref(tr._tg_a), 1, 0, 0.99, _ptr(tr._td_buf), _ptr(tr._loss_buf), _ptr(sampler.prios),
                          _ptr(sampler.max_prio), _ptr(tr._workspace), st)
def noise():
    lib.pp_noisy_reset(tr._noise_all, 4, 0, _ptr(tr._noise_counter), st)
def timed(name, fn, reps=200):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:34s} {1e3 * e0.elapsed_time(e1) / reps:8.2f} us")
timed("pp_dqn_head_grads (batch 256)", grads)
timed("pp_noisy_reset (4 layers)", noise)
timed("opt.step (torch Adam, capturable)", tr.opt.step)
timed("sampler.sample (4 M rows, 3 launches)", lambda: sampler.sample(256, 0.5))
if hasattr(tr, "_adam_step"):
    timed("pp_adam_step", tr._adam_step)
g = torch.cuda.CUDAGraph()
tr2 = pp.DQNTrainer(pp.QNet(), batch_size=256, fused=True, use_graph=True)
for _ in range(5): tr2.update(sampler)
timed("update (graph replay)", lambda: tr2.update(sampler), reps=100)
