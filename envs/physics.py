"""`from envs.physics import collide_sphere_with_moving_plane` (envs/my_pong_env_2p.py:8) -> the device routine the
env kernels use for a paddle impact (pp_collide), same argument order and return triple as envs/physics.py:3-23."""
from pingpong_selfplay_ai_b200.env import collide_sphere_with_moving_plane  # noqa: F401

__all__ = ["collide_sphere_with_moving_plane"]
