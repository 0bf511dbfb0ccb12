"""Env-slab data parallelism: one process per GPU, one contiguous slab of environments per rank.

The self-play path shards with no data-path collective (environments are independent:
envs/my_pong_env_2p.py:116-225 touches only its own state, weights are read-only during a rollout).  Device RNG
(serves, exploration, random players) is keyed by the GLOBAL env id, so results do not depend on the world size.
The only exchanges are latency-bound all-reduces: the per-generation counters (8 x int64) and, in training mode,
the flattened gradients of the trainable head parameters (520 floats for QNet: scripts/train_iterative.py:97,101-104).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def slab_bounds(num_envs: int, world_size: int, rank: int) -> tuple[int, int]:
    """[lo, hi) of this rank's slab: contiguous, sizes differ by at most one, lower ranks take the remainder."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    q, r = divmod(int(num_envs), int(world_size))
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from torchrun's environment; initialises the process group when WORLD_SIZE > 1."""
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def is_parallel() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def allreduce_counters(counters: torch.Tensor) -> torch.Tensor:
    """SUM of the int64 counter vector over all slabs (a copy; the per-rank counters stay local)."""
    total = counters.clone()
    if is_parallel():
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    return total


def allreduce_mean_grads(params) -> None:
    """Average the gradients of `params` over ranks with ONE all-reduce of the flattened vector."""
    if not is_parallel():
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(dist.get_world_size())
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


def flatten_grads_(params) -> torch.Tensor:
    """Make the .grad tensors of `params` views into ONE flat buffer (values preserved) and return it, so that the
    all-reduce needs no gather / scatter copies: allreduce_mean_flat_(flat) then updates every p.grad in place."""
    params = list(params)
    flat = torch.zeros(sum(p.numel() for p in params), dtype=params[0].dtype, device=params[0].device)
    off = 0
    for p in params:
        view = flat[off:off + p.numel()].view_as(p)
        if p.grad is not None:
            view.copy_(p.grad)
        p.grad = view
        off += p.numel()
    return flat


def allreduce_mean_flat_(flat: torch.Tensor) -> None:
    """Average a flat gradient buffer over ranks: one collective and one scaling kernel."""
    if not is_parallel():
        return
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(dist.get_world_size())


def max_over_ranks(value: float, device=None) -> float:
    """Timing helper: the slowest rank defines the step time."""
    if not is_parallel():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def broadcast_(t: torch.Tensor, src: int = 0) -> torch.Tensor:
    if is_parallel():
        dist.broadcast(t, src)
    return t


def all_ranks_ready(local_ready: bool, device=None) -> bool:
    """True iff EVERY rank is ready (all-reduce MIN of a flag).  Whether an update step runs — and with it the gradient
    all-reduce — must be ONE decision for all ranks: slabs of unequal size (slab_bounds) or different episode counts
    cross the `enough replay data` threshold at different chunks, and ranks that issue different numbers of collectives
    pair a gradient all-reduce with a counter all-reduce (hang or garbage)."""
    if not is_parallel():
        return bool(local_ready)
    t = torch.tensor([1 if local_ready else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(int(t.item()))


def broadcast_module_(module: torch.nn.Module, src: int = 0) -> None:
    """Replicas start from rank `src`'s parameters and buffers (data-parallel training averages gradients, which keeps
    replicas equal only if they start equal)."""
    if not is_parallel():
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src)
