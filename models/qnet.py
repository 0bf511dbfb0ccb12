"""`from models.qnet import QNet` (scripts/train_iterative.py:19, tests/arena.py:43) — models/qnet.py:6-75."""
from pingpong_selfplay_ai_b200.policy import NoisyLinear, QNet  # noqa: F401

__all__ = ["NoisyLinear", "QNet"]
