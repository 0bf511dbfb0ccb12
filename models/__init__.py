"""Import-compatible drop-in for the reference's `models` package (models/qnet.py, models/qnet_rnn.py): modules with
the reference's architecture, state_dict keys and initialisation, so its checkpoints load unchanged; the engine packs
their weights for the device kernels (pingpong_selfplay_ai_b200.policy)."""
