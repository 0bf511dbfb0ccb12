"""`from envs.my_pong_env_2p import PongEnv2P` (scripts/train_iterative.py:18, tests/arena.py:42) -> the n = 1 device
adaptor with the reference's constructor keywords, reset / step signatures, observation layout and reward convention
(envs/my_pong_env_2p.py:19-39,83,116,235-263).  `VecPongEnv2P` is the same interface over n lock-step envs."""
from pingpong_selfplay_ai_b200.env import PongEnv2P, VecPongEnv2P  # noqa: F401

__all__ = ["PongEnv2P", "VecPongEnv2P"]
