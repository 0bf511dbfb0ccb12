#!/usr/bin/env python
"""Headline benchmark: self-play env-steps/sec (PongEnv2P step + both players' QNet action) — BASELINE.json `metric`.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the configuration the metric is quoted on that fits one GPU): 65 536 lock-step
envs per GPU, random-init QNet A (torch.manual_seed(0)) vs QNet B (seed 1), eval-mode weights, greedy, auto-reset
with device (Philox) serves, fp64 bit-exact env arithmetic.  One "step" = one launch of the fused self-play kernel
= `--lockstep` (256) lock-step env steps of every env.  Weak scaling: every rank owns its own 65 536-env slab; the
only collective is the all-reduce of the 8 counters at the end of the timed region.

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` goes through the
host-buffer C-ABI entry (numpy serves + weights in, counters out, copies inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

ENV_CFG = dict(  # the env: block of the reference's config.yaml (values restated; the file is not on the GPU box)
    render_size=400, paddle_width=0.2, paddle_speed=0.03, max_score=3, enable_render=False, enable_spin=True,
    magnus_factor=0.025, restitution=1, friction=0.6, ball_mass=1.0, world_ball_radius=0.03,
    ball_speed_range=[0.03, 0.05], spin_range=[-5, 5], ball_angle_intervals=[[-60, -30], [30, 60]],
    speed_scale_every=1, speed_increment=0.1)
METRIC = "self-play env-steps/sec (env + both players' QNet action)"
UNIT = "env-steps/s"
FLOP_PER_ENV_STEP = 19200            # 2 players x 2 x 4800 MAC
E2E_QUOTA = 32                       # episodes per env of one end-to-end call (eval_vs_model over 65536 x 32 = 2 M games)
BYTES_PER_STEP_F64 = 203             # K1 single step, all outputs materialised, fp64 mode
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (profiles/):
NCU_TRAFFIC = {("selfplay_tc_kernel", 256): 5.088512e6 + 95.744e3,     # r01_selfplay_tc256_final_metrics.txt, 65536 envs x 256 steps
               ("selfplay_tc_kernel", 64): 5.083904e6 + 27.136e3,      # r01_selfplay_r01b_metrics.txt, 65536 envs x 64 steps
               ("selfplay_kernel", 64): 5.094144e6 + 15.36e3}           # r01_selfplay_r01_metrics.txt, same shape
NCU_K1_TRAFFIC_PER_ENV = (293.624576e6 + 499.684352e6) / 4194304   # r01_k1_r01b_metrics.txt: 189.1 B per env-step at 4 M envs


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), bf16_burst=float(p["bf16_tflops"]),
                    bf16_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), src="measured")
    except Exception:
        return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (NVML; nvidia-smi as a fallback)."""

    BAD = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80}
    NOTE = {"sw_power_cap": 0x4}

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                try:
                    r = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for name, bit in {**self.BAD, **self.NOTE}.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def stop(self) -> dict:
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)
        if not self.samples:
            try:
                import subprocess
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                a, b = (float(v) for v in out.strip().split(","))
                return dict(sm_mhz=a, sm_max_mhz=b, reasons=["unsampled: NVML unavailable, one nvidia-smi reading after the run"])
            except Exception:
                return dict(sm_mhz=None, sm_max_mhz=None, reasons=["unavailable"])
        return dict(sm_mhz=float(np.median(self.samples)), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                    samples=len(self.samples))


# ------------------------------------------------------------------------------------------------ CPU baselines
def _port_selfplay_loop(n_steps: int, seed: int) -> int:
    """The reference's own per-step eval loop (scripts/train_iterative.py:171-181) restated on the CPU port:
    pure-Python env + torch batch-1 QNet forward + argmax().item() for A and B, reset on done."""
    import random

    from oracle import pong_port
    from oracle.policy_torch import QNetPort
    torch.set_num_threads(1)
    random.seed(seed)
    torch.manual_seed(0); net_a = QNetPort().eval()
    torch.manual_seed(1); net_b = QNetPort().eval()
    env = pong_port.PongPort(**ENV_CFG)
    oa, ob = env.reset()
    with torch.no_grad():
        for _ in range(n_steps):
            a = net_a(torch.tensor(oa, dtype=torch.float32).unsqueeze(0)).argmax(1).item()
            b = net_b(torch.tensor(ob, dtype=torch.float32).unsqueeze(0)).argmax(1).item()
            (oa, ob), _, done, _ = env.step(a, b)
            if done:
                oa, ob = env.reset()
    return n_steps


def _worker(args):
    n_steps, seed = args
    t0 = time.perf_counter()
    _port_selfplay_loop(n_steps, seed)
    return time.perf_counter() - t0


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class PortPool:
    """`procs` forked workers, each running the reference-shaped loop independently (no CUDA in the children; the
    parent forks BEFORE it touches the GPU)."""

    def __init__(self, procs: int):
        import multiprocessing as mp
        self.procs = procs
        self.pool = mp.get_context("fork").Pool(procs) if procs > 1 else None

    def run(self, steps_per_proc: int) -> tuple[float, float]:
        """env-steps/s over all workers and the wall time of this sample."""
        t0 = time.perf_counter()
        if self.pool is None:
            _worker((steps_per_proc, 0))
        else:
            self.pool.map(_worker, [(steps_per_proc, s) for s in range(self.procs)], chunksize=1)
        wall = time.perf_counter() - t0
        return self.procs * steps_per_proc / wall, wall

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()


def c_oracle_throughput(n: int, k: int) -> float:
    """Single-thread C oracle (fmaf-chain QNet + fp64 env) on the same workload shape — an optimised CPU point."""
    from oracle import pong_oracle as po
    from oracle.policy_torch import QNetPort
    pool = tuple(np.ascontiguousarray(a) for a in po.serve_pool_from_reference_rng(1, n, 4, ENV_CFG))
    torch.manual_seed(0); wa = po.qnet_weights_from_state_dict(QNetPort().state_dict())
    torch.manual_seed(1); wb = po.qnet_weights_from_state_dict(QNetPort().state_dict())
    b = po.EnvBatch(n, "f64")
    b.serve(pool[0][0], pool[1][0], pool[2][0])
    pa, pb = po.make_policy(po.POLICY_QNET, wa), po.make_policy(po.POLICY_QNET, wb)
    p = po.make_params(ENV_CFG)
    po.selfplay(p, b, pa, pb, 2, pool)
    t0 = time.perf_counter()
    out = po.selfplay(p, b, pa, pb, k, pool)
    return float(out["counters"][0]) / (time.perf_counter() - t0)


def run_reference(args, rank: int):
    """`--impl reference`: the reference's CPU implementation of the path (its per-step Python loop, restated by the
    oracle port because /root/reference cannot travel to the GPU box), one process per host core."""
    if rank != 0:
        return
    cores = host_cores()
    procs = max(1, min(cores, int(os.environ.get("PP_REF_PROCS", cores))))
    per_proc = args.ref_steps
    pool = PortPool(procs)
    rate = None
    for _ in range(max(args.warmup, 1)):
        v, _ = pool.run(max(per_proc // 8, 50))
        rate = v / procs                                   # env-steps/s of one process
    # keep the whole timed run within ~2.5 minutes whatever --steps is: a step is a BOUNDED sample of the workload
    per_proc = int(min(per_proc, max(50, 150.0 * rate / max(args.steps, 1))))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pool.run(per_proc)
    wall = time.perf_counter() - t0
    pool.close()
    value = procs * per_proc * args.steps / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 env state + f32 QNet", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": f"{procs} processes x {per_proc} env-steps per step of the reference-shaped loop "
                                   "(pure-Python PongEnv2P port + torch batch-1 QNet A and B, argmax().item()), "
                                   "torch.set_num_threads(1) per process"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    if getattr(args, "workload", "qnet") == "rnn":
        wl = ("configs[3] shape: QNetRNN (LSTM) A vs B rollout with per-env hidden state, fused obs->QNetRNN->argmax + env "
              "step, greedy, auto-reset (Philox serves), config.yaml env params")
    else:
        wl = ("configs[2]: self-play eval, random-init QNet A vs QNet B, 65536 lock-step PongEnv2P envs per GPU, "
              "fused obs->QNet->argmax + env step, greedy, auto-reset (Philox serves), config.yaml env params")
    return {"workload": wl,
            "envs_per_gpu": args.envs, "lockstep_steps_per_launch": args.lockstep, "env_mode": args.mode,
            "qnet_precision": args.precision, "parallelism": f"env-slab dp{world}",
            "l2": "flushed between timed steps (256 MiB memset outside the per-step CUDA-event brackets); "
                  "state is register-resident inside a launch"}


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=65536, help="envs per GPU")
    ap.add_argument("--lockstep", type=int, default=None,
                    help="lock-step env steps per launch (default: 20480 for the headline workload = 0.11 s per launch, so "
                         "that the timed region of --steps 20 is seconds of sustained load; 64 for the others)")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the config4_rnn / config5_train objects (BASELINE.json configs[3] and configs[4])")
    ap.add_argument("--mode", default="f64", choices=["f64", "f32"])
    ap.add_argument("--precision", default="f16", choices=["f32", "f16"],
                    help="QNet path: f16 = tcgen05 tensor cores (fp16 hi/lo operands, fp32 accumulate, ~1e-6 of fp32); "
                         "f32 = CUDA-core fmaf chain, bit-identical to the oracle")
    ap.add_argument("--workload", default="qnet", choices=["qnet", "rnn", "train", "train_rnn", "arena"],
                    help="qnet = configs[2] (the headline); rnn = configs[3] shape: QNetRNN A vs B with per-env (h, c); "
                         "train = configs[4] shape: epsilon-greedy rollout + replay scatter + PER Double-DQN updates + grad all-reduce; "
                         "train_rnn = DRQN training mode: recurrent rollout + lock-step ring + sequence updates; "
                         "arena = a tests/arena.py-shaped round robin (10 agents, 100 games per pairing) on the batched engine")
    ap.add_argument("--ref-steps", type=int, default=1500, help="reference arm: env-steps per process per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-calls", type=int, default=50, help="timed calls of the host-buffer entry")
    ap.add_argument("--no-k1", action="store_true", help="skip the single-step env kernel HBM roofline measurement")
    ap.add_argument("--k1-envs", type=int, default=16 << 20, help="envs of the single-step HBM roofline measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.lockstep is None:
        args.lockstep = 20480 if args.workload == "qnet" else 64

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    if args.workload == "arena":
        run_arena_workload(args)
        return
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:      # before CUDA is touched: the workers are forked
        cpu_baseline = measure_cpu_baseline()

    import pingpong_selfplay_ai_b200 as pp
    from pingpong_selfplay_ai_b200 import dist as ppd

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path (use --impl reference for the CPU arm)")
    rank, world, local = ppd.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    peaks = load_peaks()

    n, k = args.envs, args.lockstep
    lo = rank * n
    env = pp.VecPongEnv2P(n, device=dev, mode=args.mode, serve="philox", seed=2026, env_id_base=lo, **ENV_CFG)
    env.reset()
    if args.workload == "rnn":
        torch.manual_seed(0); net_a = pp.QNetRNN()
        torch.manual_seed(1); net_b = pp.QNetRNN()
        pa = pp.Policy.qnetrnn(net_a, num_envs=n, device=dev, precision=args.precision)
        pb = pp.Policy.qnetrnn(net_b, num_envs=n, device=dev, precision=args.precision)
        args.no_e2e = args.no_k1 = True
    else:
        torch.manual_seed(0); net_a = pp.QNet()
        torch.manual_seed(1); net_b = pp.QNet()
        pa = pp.Policy.qnet(net_a, device=dev, precision=args.precision)
        pb = pp.Policy.qnet(net_b, device=dev, precision=args.precision)
    eng = pp.SelfPlayEngine(env, pa, pb, seed=7)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    if args.workload == "train":
        run_train_workload(pp, ppd, args, env, net_a, net_b, dev, rank, world)
        return
    if args.workload == "train_rnn":
        run_train_rnn_workload(pp, ppd, args, env, dev, rank, world)
        return
    for _ in range(args.warmup):
        eng.run(k)
    torch.cuda.synchronize()
    env.counters.zero_()
    sampler = ClockSampler(local).start()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    for s, e in ev:
        flush.zero_()                                   # L2 flush, outside the event bracket
        s.record()
        eng.run(k)                                      # ONE launch of the fused self-play kernel
        e.record()
    total = ppd.allreduce_counters(env.counters)        # the path's only collective (8 x int64)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    dev_ms = sum(s.elapsed_time(e) for s, e in ev)
    dev_ms = ppd.max_over_ranks(dev_ms, dev)
    env_steps = int(total[0].item())
    assert env_steps == world * n * k * args.steps, (env_steps, world, n, k, args.steps)
    value = env_steps / (dev_ms * 1e-3)

    line = {
        "metric": METRIC if args.workload == "qnet" else METRIC.replace("QNet", "QNetRNN"),
        "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": f"{args.mode} env state + {args.precision} QNet", "data": "synthetic",
        "config": workload_config(args, world), "clocks": clocks,
        # per timed step on every rank: the fused rollout kernel, plus for the tensor-core QNet path the one-block kernel
        # that builds the head tables (its device-to-device copy into constant memory is a memcpy node, not counted)
        "gpu_launches": args.steps * (2 if (args.workload == "qnet" and args.precision == "f16"
                                            and os.environ.get("PP_CONST_HEADS", "1")[:1] != "0") else 1),
        "wall_s_timed_region_incl_flush": wall,
        "outcomes": {"episodes": int(total[1].item()), "wins_a": int(total[2].item()), "wins_b": int(total[3].item()),
                     "paddle_hits": int(total[6].item())},
    }
    flop = 626432 if args.workload == "rnn" else FLOP_PER_ENV_STEP
    tf = value / world * flop / 1e12
    kname = ("selfplay_rnn_tc_kernel" if args.precision == "f16" else "selfplay_rnn_kernel") if args.workload == "rnn" else \
        ("selfplay_tc_kernel" if args.precision == "f16" else "selfplay_kernel")
    traffic = NCU_TRAFFIC.get((kname, args.lockstep)) if (args.envs, args.mode) == (65536, "f64") else None
    line["roofline"] = {"bound": "tensor", "kernel": ("selfplay_rnn_tc_kernel" if args.precision == "f16" else "selfplay_rnn_kernel") if args.workload == "rnn" else ("selfplay_tc_kernel" if args.precision == "f16" else "selfplay_kernel"), "achieved": tf, "peak": peaks["bf16_sustained"],
                        "unit": "TFLOP/s", "frac": tf / peaks["bf16_sustained"], "traffic": traffic,
                        "traffic_note": "DRAM bytes per launch from the committed ncu --set full capture of this shape (profiles/); "
                                        "state lives in registers, so it is ~0.1 % of what a materialising step would move",
                        "peak_source": f"{peaks['src']} bf16 sustained (kernel timed inside a long step)",
                        "note": f"algorithmic {flop} FLOP per env-step (both players' net) x env-steps per launch / "
                                "CUDA-event launch time, per GPU"}

    line["timed_region_s"] = dev_ms * 1e-3
    if rank == 0 and not args.no_k1:
        line["roofline_env_step"] = measure_k1(pp, dev, peaks, args.mode, args.k1_envs)
    if not args.no_e2e:
        line["e2e"] = measure_e2e(pp, net_a, net_b, args, rank, world, local, dev)
    if args.workload == "qnet" and not args.no_secondary:
        del eng, env, flush
        torch.cuda.empty_cache()
        line["config4_rnn"] = measure_config4_rnn(pp, ppd, args, dev, rank, world, local, peaks)
        torch.cuda.empty_cache()
        line["config5_train"] = measure_config5_train(pp, ppd, args, dev, rank, world, local, peaks)
    if cpu_baseline is not None:
        line.update(cpu_baseline)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def run_train_workload(pp, ppd, args, env, net_a, net_b, dev, rank, world):
    """configs[4] shape per GPU: one training generation chunk loop (rollout with replay rows, prioritised batches,
    Double-DQN on the heads, gradient and counter all-reduce).  One bench step = one chunk of `--lockstep` steps followed
    by 4 updates of batch 256 per rank.  Prints its own JSON line (not the headline metric)."""
    n, k = args.envs, args.lockstep
    trainer = pp.DQNTrainer(net_b, batch_size=256, device=dev)
    eng = pp.SelfPlayEngine(env, pp.Policy.qnet(net_a, noisy=True, precision=args.precision, device=dev),
                            pp.Policy.qnet(net_b, noisy=True, eps=0.5, precision=args.precision, device=dev), seed=7)
    ring = pp.ReplayRing(max(1 << 20, n * k), device=dev)
    sampler = pp.PrioritizedSampler(ring)
    pp.train_generation(eng, trainer, ring, sampler, k * args.warmup, chunk=k, updates_per_chunk=4, epsilon=0.5, precision=args.precision)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    out = pp.train_generation(eng, trainer, ring, sampler, k * args.steps, chunk=k, updates_per_chunk=4, epsilon=0.5, precision=args.precision)
    torch.cuda.synchronize()
    wall = ppd.max_over_ranks(time.perf_counter() - t0, dev)
    if rank == 0:
        print(json.dumps({"metric": "training-mode env-steps/sec (rollout + replay + DQN updates)", "value": out["env_steps"] / wall,
                          "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
                          "updates_per_s": out["updates"] / wall, "mean_loss": out["mean_loss"], "epsilon": out["epsilon"],
                          "config": {"workload": "configs[4] shape: train generation chunk loop", "envs_per_gpu": n,
                                     "lockstep_steps_per_chunk": k, "updates_per_chunk": 4, "batch_per_rank": 256,
                                     "replay_capacity": ring.capacity, "qnet_precision": args.precision}}), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def _port_arena_games(games: int) -> tuple[int, float]:
    """The match loop of tests/arena.py:294-308 for one QNetRNN x QNetRNN pairing on the CPU port (pure-Python env +
    torch batch-1 recurrent forwards) — what the reference spends per game.  -> (env-steps, seconds)."""
    import random

    from oracle import pong_port
    from oracle.policy_torch import QNetRNNPort
    torch.set_num_threads(1)
    random.seed(1)
    torch.manual_seed(0); net_a = QNetRNNPort().eval()
    torch.manual_seed(1); net_b = QNetRNNPort().eval()
    env = pong_port.PongPort(**ENV_CFG)
    steps, t0 = 0, time.perf_counter()
    with torch.no_grad():
        for _ in range(games):
            oa, ob = env.reset()
            ha, hb, done = net_a.init_hidden(1, "cpu"), net_b.init_hidden(1, "cpu"), False
            while not done:
                qa, ha = net_a(torch.tensor(oa, dtype=torch.float32).unsqueeze(0).unsqueeze(0), ha)
                qb, hb = net_b(torch.tensor(ob, dtype=torch.float32).unsqueeze(0).unsqueeze(0), hb)
                (oa, ob), _, done, _ = env.step(int(qa.argmax(1).item()), int(qb.argmax(1).item()))
                steps += 1
    return steps, time.perf_counter() - t0


def run_arena_workload(args):
    """A tests/arena.py-shaped tournament (ARENA_CONFIG: 2 QNet + 7 QNetRNN + the ball follower, 100 games per pairing,
    45 pairings) on the batched engine, one launch per pairing, pairings overlapped on CUDA streams.  One bench step = one
    whole tournament incl. the host logic (database records, result read-back).  Prints its own JSON line."""
    cpu = None
    if not args.no_cpu_baseline:
        st, sec = _port_arena_games(300)
        cpu = {"games_per_s": 300 / sec, "env_steps_per_s": st / sec, "cores": 1, "kind": "port",
               "sample": "300 QNetRNN x QNetRNN games, reference-shaped match loop (tests/arena.py:294-308), one process"}
    import pingpong_selfplay_ai_b200 as pp
    from pingpong_selfplay_ai_b200 import arena, checkpoint as ck
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    models, agents = [], {}
    for k in range(10):
        torch.manual_seed(k)
        typ = "QNet" if k < 2 else ("QNetRNN" if k < 9 else "HardcodedBallFollower")
        net = pp.QNet() if typ == "QNet" else (pp.QNetRNN() if typ == "QNetRNN" else None)
        info = {"id": f"agent{k}", "type": typ, "path": "random-init"}
        models.append(info)
        agents[info["id"]] = ck.Agent(info, net.eval() if net is not None else None)
    games = 100

    def tournament(seed):
        db = {"models": list(models), "match_history": []}
        plan = arena.create_match_plan(db, games)
        res = arena.run_tournament(ENV_CFG, db, None, plan, agents=agents, seed=seed, precision=args.precision, concurrent=8)
        return len(db["match_history"]), int(sum(int(r[2].sum()) for r in res.values()))
    for w in range(max(args.warmup, 1)):
        tournament(1000 * w)
    torch.cuda.synchronize()
    reps = max(1, min(args.steps, 20))
    t0, n_games, n_steps = time.perf_counter(), 0, 0
    for r in range(reps):
        g, st = tournament(77 + 1000 * r)
        n_games += g; n_steps += st
    wall = time.perf_counter() - t0
    print(json.dumps({"metric": "arena round robin games/sec (45 pairings x 100 games, host logic included)",
                      "value": n_games / wall, "unit": "games/s", "env_steps_per_s": n_steps / wall, "n_gpus": 1,
                      "steps": reps, "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * wall / reps, "higher_is_better": True,
                      "config": {"workload": "tests/arena.py ARENA_CONFIG shape: 2 QNet + 7 QNetRNN + ball follower, random init",
                                 "pairings": 45, "games_per_pairing": games, "qnet_precision": args.precision,
                                 "concurrent_pairings": 8},
                      "cpu_baseline": cpu}), flush=True)


def run_train_rnn_workload(pp, ppd, args, env, dev, rank, world):
    """DRQN training mode per GPU (scripts/train_rnn_iterative.py:728-800 shape): recurrent B learns against recurrent A;
    one bench step = one chunk of `--lockstep` steps written to the lock-step ring + 4 sequence updates of batch 64 per
    rank (trace length 8, all 175 k parameters, gradient all-reduce).  Prints its own JSON line (not the headline)."""
    from pingpong_selfplay_ai_b200.train_rnn import DRQNTrainer, SequenceSampler, train_rnn_generation
    n, k = args.envs, args.lockstep
    torch.manual_seed(0); net_a = pp.QNetRNN()
    torch.manual_seed(1); net_b = pp.QNetRNN()
    trainer = DRQNTrainer(net_b, batch_size=64, device=dev)
    eng = pp.SelfPlayEngine(env, pp.Policy.qnetrnn(net_a, num_envs=n, precision=args.precision, device=dev),
                            pp.Policy.qnetrnn(trainer.model, num_envs=n, noisy=True, eps=0.5, precision=args.precision, device=dev), seed=7)
    steps_in_ring = max(64, k)
    ring = pp.ReplayRing(n * steps_in_ring, device=dev, lockstep_envs=n)
    sampler = SequenceSampler(ring, trace_length=8)
    kw = dict(chunk=k, updates_per_chunk=4, epsilon=0.5, precision=args.precision)
    train_rnn_generation(eng, trainer, ring, sampler, k * args.warmup, **kw)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    out = train_rnn_generation(eng, trainer, ring, sampler, k * args.steps, **kw)
    torch.cuda.synchronize()
    wall = ppd.max_over_ranks(time.perf_counter() - t0, dev)
    if rank == 0:
        print(json.dumps({"metric": "DRQN training-mode env-steps/sec (recurrent rollout + sequence replay + updates)",
                          "value": out["env_steps"] / wall, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * wall / args.steps, "updates_per_s": out["updates"] / wall,
                          "mean_loss": out["mean_loss"], "epsilon": out["epsilon"], "stored_episodes": out["stored_episodes"],
                          "config": {"workload": "DRQN train generation chunk loop (config_rnn.yaml training block shape)",
                                     "envs_per_gpu": n, "lockstep_steps_per_chunk": k, "updates_per_chunk": 4,
                                     "batch_per_rank": 64, "trace_length": 8, "ring_steps": steps_in_ring,
                                     "qnet_precision": args.precision}}), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def measure_cpu_baseline() -> dict:
    procs = max(1, min(host_cores(), 64))
    per_proc = 120000
    pool = PortPool(procs)
    pool.run(200)
    v, w = pool.run(per_proc)
    pool.close()
    out = {"cpu_baseline": {"value": v, "unit": UNIT, "cores": procs, "kind": "port",
                            "sample": f"{procs} processes x {per_proc} env-steps of the reference-shaped loop (pure-Python "
                                      f"PongEnv2P port + torch batch-1 QNet A and B, argmax().item()), {w:.1f} s wall, "
                                      f"{host_cores()} host cores available"}}
    try:
        out["cpu_baseline_c_oracle"] = {"value": c_oracle_throughput(4096, 64), "unit": UNIT, "cores": 1, "kind": "port",
                                        "sample": "C oracle closed loop (fp64 env + fmaf-chain QNet), 4096 envs x 64 steps, 1 thread"}
    except Exception as e:  # an extra data point, not part of the contract
        out["cpu_baseline_c_oracle"] = {"error": str(e)}
    return out


RNN_ENV_CFG = dict(ENV_CFG, speed_scale_every=5, speed_increment=0.2)      # config_rnn.yaml:27-28 (restated)
RNN_FLOP_PER_ENV_STEP = 626432       # 2 players x 2 x 156 608 MAC
CONFIG4_ENVS, CONFIG5_ENVS = 262144, 1048576


def _port_rnn_loop(n_steps: int) -> float:
    """QNetRNN A vs B on the CPU port (tests/arena.py:294-308 shaped: batch-1 recurrent forwards) -> env-steps/s."""
    import random

    from oracle import pong_port
    from oracle.policy_torch import QNetRNNPort
    torch.set_num_threads(1)
    random.seed(1)
    torch.manual_seed(0); net_a = QNetRNNPort().eval()
    torch.manual_seed(1); net_b = QNetRNNPort().eval()
    env = pong_port.PongPort(**RNN_ENV_CFG)
    oa, ob = env.reset()
    ha, hb = net_a.init_hidden(1, "cpu"), net_b.init_hidden(1, "cpu")
    t0 = time.perf_counter()
    with torch.no_grad():
        for _ in range(n_steps):
            qa, ha = net_a(torch.tensor(oa, dtype=torch.float32).unsqueeze(0).unsqueeze(0), ha)
            qb, hb = net_b(torch.tensor(ob, dtype=torch.float32).unsqueeze(0).unsqueeze(0), hb)
            (oa, ob), _, done, _ = env.step(int(qa.argmax(1).item()), int(qb.argmax(1).item()))
            if done:
                oa, ob = env.reset()
                ha, hb = net_a.init_hidden(1, "cpu"), net_b.init_hidden(1, "cpu")
    return n_steps / (time.perf_counter() - t0)


def _port_train_loop(n_steps: int) -> float:
    """The reference's training loop (scripts/train_iterative.py:238-261: env step + both forwards + memory.push +
    train_step of batch 256 per env step) on the CPU port -> env-steps/s."""
    from oracle import pong_port
    from oracle.policy_torch import QNetPort
    from oracle.train_port import TrainLoopPort
    import random
    torch.set_num_threads(1)
    random.seed(3); np.random.seed(3)
    torch.manual_seed(0); a = QNetPort()
    torch.manual_seed(1); b = QNetPort()
    loop = TrainLoopPort(pong_port.PongPort(**ENV_CFG), a, b, memory_size=100000)
    loop.run(300)                                       # fills the first batch: train_step runs from step 256 on
    t0 = time.perf_counter()
    loop.run(n_steps)
    return n_steps / (time.perf_counter() - t0)


def _sum_over_ranks(values, dev, world):
    if world == 1:
        return [int(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.int64, device=dev)
    torch.distributed.all_reduce(t)
    return [int(v) for v in t.tolist()]


def measure_config4_rnn(pp, ppd, args, dev, rank, world, local, peaks):
    """BASELINE.json configs[3]: QNetRNN (LSTM) A vs B rollout with per-env hidden state, 262 144 envs SPLIT over the
    ranks (strong scaling), fused obs -> QNetRNN -> argmax + env step on the tensor-core recurrent kernel."""
    lo, hi = ppd.slab_bounds(CONFIG4_ENVS, world, rank)
    n, k, launches, warm = hi - lo, 32, 16, 3
    env = pp.VecPongEnv2P(n, device=dev, mode=args.mode, serve="philox", seed=2026, env_id_base=lo, **RNN_ENV_CFG)
    env.reset()
    torch.manual_seed(0); net_a = pp.QNetRNN()
    torch.manual_seed(1); net_b = pp.QNetRNN()
    eng = pp.SelfPlayEngine(env, pp.Policy.qnetrnn(net_a, num_envs=n, device=dev, precision=args.precision),
                            pp.Policy.qnetrnn(net_b, num_envs=n, device=dev, precision=args.precision), seed=7)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(warm):
        eng.run(k)
    torch.cuda.synchronize()
    env.counters.zero_()
    sampler = ClockSampler(local).start()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(launches)]
    for s, e in ev:
        flush.zero_()
        s.record(); eng.run(k); e.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = ppd.max_over_ranks(sum(s.elapsed_time(e) for s, e in ev), dev)
    steps, episodes = _sum_over_ranks([env.counters[0].item(), env.counters[1].item()], dev, world)
    assert steps == CONFIG4_ENVS * k * launches, (steps, CONFIG4_ENVS, k, launches)
    value = steps / (ms * 1e-3)
    tf = value / world * RNN_FLOP_PER_ENV_STEP / 1e12
    out = {"metric": "QNetRNN self-play env-steps/sec (env + both players' LSTM action)", "value": value, "unit": UNIT,
           "n_gpus": world, "scaling": "strong", "envs_total": CONFIG4_ENVS, "envs_per_gpu": n, "lockstep_steps_per_launch": k,
           "launches": launches, "warmup": warm, "ms_per_launch": ms / launches, "env_steps_per_s_per_gpu": value / world,
           "dtype": f"{args.mode} env state + {args.precision} QNetRNN", "clocks": clocks, "episodes": episodes,
           "roofline": {"bound": "tensor", "kernel": "selfplay_rnn_tc_kernel" if args.precision == "f16" else "selfplay_rnn_kernel",
                        "achieved": tf, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": tf / peaks["bf16_sustained"],
                        "traffic": None, "peak_source": f"{peaks['src']} bf16 sustained",
                        "note": f"algorithmic {RNN_FLOP_PER_ENV_STEP} FLOP per env-step (both players' net) x env-steps per "
                                "launch / CUDA-event launch time, per GPU"},
           "config": {"workload": "configs[3]: QNetRNN (LSTM) A vs B rollout with per-env hidden state, 262144 envs split over "
                                  "the GPUs, greedy, auto-reset (Philox serves), config_rnn.yaml env params",
                      "l2": "flushed between timed launches"}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v = _port_rnn_loop(2500)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": "2500 env-steps of the reference-shaped recurrent loop (pure-Python env port + torch "
                                         "batch-1 QNetRNN A and B), one process"}
    return out


def measure_config5_train(pp, ppd, args, dev, rank, world, local, peaks):
    """BASELINE.json configs[4]: a train_iterative generation chunk loop, 1 048 576 envs SPLIT over the ranks:
    epsilon-greedy rollout of the learner B (train-mode NoisyNet weights) + replay rows + PER Double-DQN updates of
    batch 256 per rank with the NCCL all-reduce of the 520 head gradients inside every update + counter all-reduce."""
    lo, hi = ppd.slab_bounds(CONFIG5_ENVS, world, rank)
    n, k, upc = hi - lo, 64, 4
    chunks, warm = 24 * world, 6                        # every rank times the same number of chunks (collectives inside)
    env = pp.VecPongEnv2P(n, device=dev, mode=args.mode, serve="philox", seed=2027, env_id_base=lo, **ENV_CFG)
    env.reset()
    torch.manual_seed(0); net_a = pp.QNet()
    torch.manual_seed(1); net_b = pp.QNet()
    trainer = pp.DQNTrainer(net_b, batch_size=256, device=dev)
    eng = pp.SelfPlayEngine(env, pp.Policy.qnet(net_a, noisy=True, precision=args.precision, device=dev),
                            pp.Policy.qnet(net_b, noisy=True, eps=0.5, precision=args.precision, device=dev), seed=7)
    ring = pp.ReplayRing(max(1 << 20, n * k), device=dev)
    sampler = pp.PrioritizedSampler(ring)
    kw = dict(chunk=k, updates_per_chunk=upc, epsilon=0.5, precision=args.precision)
    pp.train_generation(eng, trainer, ring, sampler, k * warm, **kw)
    torch.cuda.synchronize()
    clock = ClockSampler(local).start()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = pp.train_generation(eng, trainer, ring, sampler, k * chunks, **kw)
    e1.record()
    torch.cuda.synchronize()
    clocks = clock.stop()
    sec = ppd.max_over_ranks(e0.elapsed_time(e1) * 1e-3, dev)
    assert out["env_steps"] == CONFIG5_ENVS * k * chunks and out["updates"] == upc * chunks
    # the collective alone: the flat 520-float gradient buffer, 200 all-reduces back to back
    ar_us = None
    if world > 1:                                       # the NCCL collective alone, for reference (the update may not use it)
        flat = trainer._flat_grad.clone()
        for _ in range(10):
            ppd.allreduce_mean_flat_(flat)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(200):
            ppd.allreduce_mean_flat_(flat)
        a1.record(); torch.cuda.synchronize()
        ar_us = ppd.max_over_ranks(a0.elapsed_time(a1) * 1e3 / 200, dev)
    value = out["env_steps"] / sec
    tf = value / world * FLOP_PER_ENV_STEP / 1e12
    res = {"metric": "training-mode env-steps/sec (epsilon-greedy rollout + replay rows + PER Double-DQN updates)",
           "value": value, "unit": UNIT, "n_gpus": world, "scaling": "strong", "envs_total": CONFIG5_ENVS, "envs_per_gpu": n,
           "lockstep_steps_per_chunk": k, "updates_per_chunk": upc, "batch_per_rank": 256, "chunks": chunks, "warmup_chunks": warm,
           "ms_per_chunk": 1e3 * sec / chunks, "updates_per_s": out["updates"] / sec, "grad_allreduce_us": ar_us,
           "grad_allreduce": ("single rank: none" if world == 1 else
                              (trainer._p2p_note if getattr(trainer, "_p2p", None) is not None else
                               ("NCCL all-reduce captured inside the update's CUDA graph" if getattr(trainer, "_split", True) is False
                                else "eager NCCL all-reduce between two CUDA graphs") + "; " + getattr(trainer, "_p2p_note", ""))),
           "mean_loss": out["mean_loss"], "epsilon": out["epsilon"], "episodes": out["episodes"], "replay_capacity": ring.capacity,
           "dtype": f"{args.mode} env state + {args.precision} QNet rollout, fp32 update", "clocks": clocks,
           "roofline": {"bound": "tensor", "kernel": "selfplay_tc_kernel (rollout with replay rows)", "achieved": tf,
                        "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": tf / peaks["bf16_sustained"], "traffic": None,
                        "peak_source": f"{peaks['src']} bf16 sustained",
                        "note": f"algorithmic {FLOP_PER_ENV_STEP} FLOP per env-step x env-steps / chunk-loop time incl. the "
                                "updates, per GPU; replay rows add 62 B of HBM writes per env-step"},
           "config": {"workload": "configs[4]: train_iterative generation chunk loop, 1048576 envs split over the GPUs, "
                                  "epsilon-greedy rollout + replay scatter + batched PER Double-DQN + NCCL all-reduce of "
                                  "gradients and counters"}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v = _port_train_loop(400)
        res["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": "400 env-steps of the reference's training loop restated on the CPU port (env step + "
                                         "both forwards + PER push + one train_step of batch 256 per env step), one process"}
    return res


def measure_k1(pp, dev, peaks, mode, n):
    """K1 single-step kernel (all outputs materialised) on a working set far beyond L2: HBM roofline."""
    env = pp.VecPongEnv2P(n, device=dev, mode=mode, serve="philox", seed=1, **ENV_CFG)
    env.reset()
    g = torch.Generator(device=dev).manual_seed(0)
    aa = torch.randint(0, 3, (n,), dtype=torch.uint8, device=dev, generator=g)
    ab = torch.randint(0, 3, (n,), dtype=torch.uint8, device=dev, generator=g)
    for _ in range(3):
        env.step(aa, ab)
    torch.cuda.synchronize()
    reps = 20
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        env.step(aa, ab)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    per = BYTES_PER_STEP_F64 if mode == "f64" else 147
    gbs = n * per / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "step_kernel", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s",
            "frac": gbs / peaks["hbm"], "traffic": NCU_K1_TRAFFIC_PER_ENV * n if mode == "f64" else None,
            "traffic_note": "ncu dram bytes per env-step measured at 4 M envs (189.1 B vs 203 B algorithmic) x envs",
            "envs": n, "ms_per_launch": ms,
            "env_steps_per_s": n / (ms * 1e-3), "peak_source": f"{peaks['src']} copy bandwidth",
            "note": f"{per} algorithmic B per env-step (SURVEY.md 8d), working set {n * per / 2**20:.0f} MiB >> L2"}


def measure_e2e(pp, net_a, net_b, args, rank=0, world=1, local=0, dev=None):
    """The same metric through the reference-facing host-buffer call (pp_host_selfplay_eval = eval_vs_model,
    scripts/train_iterative.py:171-181, for `envs` envs x `E2E_QUOTA` episodes each): packed weights in PINNED host memory
    in, counters out; the H2D / D2H copies and every sync are inside the timed region (wall clock around the synchronous
    C-ABI call).  Serves are drawn on the device from a host-supplied seed (the counterpart of random.seed()).  Every
    rank calls the entry for ITS device and slab (device ordinal + env_id_base); value = all ranks' env-steps / slowest
    rank's wall time."""
    from pingpong_selfplay_ai_b200 import dist as ppd
    n, quota, reps = args.envs, E2E_QUOTA, args.e2e_calls
    keep = []

    def pinned(a):                                      # the step's inputs live in PINNED host memory (bench contract)
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        keep.append(t)
        return t.numpy()
    wa, wb = pinned(pp.pack_qnet(net_a).cpu().numpy()), pinned(pp.pack_qnet(net_b).cpu().numpy())
    call = lambda s: pp.host_selfplay_eval(ENV_CFG, n, quota, None, wa, wb, mode=args.mode, precision=args.precision,
                                           device=local, seed=s, env_id_base=rank * n)
    for s in range(3):
        call(1000 + s)                                  # warm-up (staging buffers, kernel attributes)
    if world > 1:
        torch.distributed.barrier()
    steps_total, episodes, t0 = 0, 0, time.perf_counter()
    for s in range(reps):
        c, _ = call(s)
        steps_total += c["env_steps"]; episodes += c["episodes"]
    wall = time.perf_counter() - t0
    wall_max = ppd.max_over_ranks(wall, dev)
    if world > 1:
        t = torch.tensor([steps_total, episodes], dtype=torch.int64, device=dev)
        torch.distributed.all_reduce(t)
        steps_total, episodes = int(t[0].item()), int(t[1].item())
    in_bytes = 2 * ((wa.nbytes + 255) // 256 * 256) + 256
    return {"value": steps_total / wall_max, "unit": UNIT, "h2d_bytes_per_step": int(in_bytes), "d2h_bytes_per_step": 72,
            "call": f"pp_host_selfplay_eval on every rank (eval_vs_model for {n} envs x {quota} episodes per GPU; pinned host "
                    "weight blobs in, counters out; serves drawn on the device from a host seed; serve queue, one launch)",
            "calls": reps, "ranks": world, "episodes_per_call_per_gpu": episodes // (reps * world),
            "env_steps_per_call_per_gpu": steps_total // (reps * world), "ms_per_call": 1e3 * wall_max / reps}


if __name__ == "__main__":
    main()
