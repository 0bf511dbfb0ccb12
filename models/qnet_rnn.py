"""`from models.qnet_rnn import QNetRNN` (tests/arena.py:44, scripts/train_rnn_iterative.py) — models/qnet_rnn.py:8-152."""
from pingpong_selfplay_ai_b200.policy import NoisyLinear, QNetRNN  # noqa: F401

__all__ = ["NoisyLinear", "QNetRNN"]
