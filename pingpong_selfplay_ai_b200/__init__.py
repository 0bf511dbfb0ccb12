"""pingpong_selfplay_ai_b200 — B200-native batched self-play engine for the PongEnv2P hot path of
MaxChen228/pingpong-selfplay-ai (envs/my_pong_env_2p.py + envs/physics.py stepped in lock step with both
paddles' actions chosen by QNet / QNetRNN on the device).

Host side: Python/PyTorch (buffers, streams, torch.distributed).  Compute: hand-written sm_100a CUDA in
`csrc/`, reached through the C ABI of `include/pong_b200.h` (libpong_b200.so).  There is no CPU fallback.
"""
from . import _lib
from ._lib import PongB200Error
from .env import (COUNTER_NAMES, PongEnv2P, ServePool, VecPongEnv2P, collide_batch,
                  collide_sphere_with_moving_plane)
from .params import ENV_DEFAULTS, make_params, resolve_env_config
from .policy import NoisyLinear, Policy, QNet, QNetRNN, pack_qnet, pack_qnetrnn
from .selfplay import (ReplayRing, SelfPlayEngine, eval_vs_model, eval_vs_pool, host_selfplay_eval, qnet_act,
                       qnetrnn_act)
from .train import DQNTrainer, PrioritizedSampler, train_generation
from . import arena, checkpoint
from .checkpoint import Agent, load_agent

__all__ = [
    "PongB200Error", "PongEnv2P", "VecPongEnv2P", "ServePool", "COUNTER_NAMES", "ENV_DEFAULTS", "make_params",
    "resolve_env_config", "collide_batch", "collide_sphere_with_moving_plane", "NoisyLinear", "QNet", "QNetRNN", "Policy", "pack_qnet", "pack_qnetrnn", "ReplayRing",
    "SelfPlayEngine", "host_selfplay_eval", "qnet_act", "qnetrnn_act", "DQNTrainer", "PrioritizedSampler", "train_generation",
    "arena", "checkpoint", "Agent", "load_agent", "eval_vs_model", "eval_vs_pool",
]
