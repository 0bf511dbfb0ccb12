"""Import-compatible drop-in for the reference's `envs` package (envs/my_pong_env_2p.py, envs/physics.py): the same
names, backed by libpong_b200.so.  Put this repository's root on PYTHONPATH ahead of the reference's and
`scripts/train_iterative.py` / `tests/arena.py` import the device environment unchanged."""
