"""Two or more ranks: the fused NVLink all-reduce + Adam (pp_adam_step_allreduce) against an NCCL all-reduce + pp_adam_step.
   torchrun --nproc-per-node 2 tools/p2p_allreduce_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import pingpong_selfplay_ai_b200 as pp
from pingpong_selfplay_ai_b200 import dist as ppd

rank, world, local = ppd.init_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
torch.manual_seed(0)
a = pp.DQNTrainer(pp.QNet(), batch_size=64, use_graph=False, lr=3e-3, device=dev)       # p2p path (if available)
os.environ["PP_P2P_ALLREDUCE"] = "0"
b = pp.DQNTrainer(pp.QNet(), batch_size=64, use_graph=False, lr=3e-3, device=dev)       # NCCL path
b.model.load_state_dict(a.model.state_dict())
print(f"rank {rank}: a: {a._p2p_note} | b: {b._p2p_note}", flush=True)
assert a._p2p is not None, "symmetric memory path not available: " + a._p2p_note
g = torch.Generator(device=dev).manual_seed(100 + rank)
for step in range(50):
    grads = torch.randn(a._flat_grad.numel(), generator=g, device=dev)
    a._flat_grad.copy_(grads); b._flat_grad.copy_(grads)
    a._allreduce_grads(); a._post(None)
    b._allreduce_grads(); b._post(None)
    if step % 7 == 0:
        torch.cuda._sleep(int(2e6) * (rank + 1))            # skew the ranks: the kernel has to wait for its peers
torch.cuda.synchronize()
for p, q in zip(a.head_params, b.head_params):
    assert torch.allclose(p, q, rtol=1e-5, atol=1e-7), (p - q).abs().max()
flat = torch.cat([p.detach().flatten() for p in a.head_params])
others = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(others, flat)
assert all(torch.equal(o, others[0]) for o in others), "replicas diverged"
# timing: update tail with NCCL + Adam vs the fused launch
def timed(fn, reps=300):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps
t_p2p = timed(lambda: a._post(None))
t_nccl = timed(lambda: (b._allreduce_grads(), b._post(None)))
if rank == 0:
    print(f"ok: {world} ranks bit-identical replicas; fused all-reduce + Adam {t_p2p:.1f} us per update, NCCL all-reduce + Adam {t_nccl:.1f} us", flush=True)
dist.barrier(); dist.destroy_process_group()
