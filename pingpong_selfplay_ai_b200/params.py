"""Environment parameters: the `env:` block of the reference's config.yaml -> PPParams.

The keyword names and defaults are those of `PongEnv2P.__init__` (envs/my_pong_env_2p.py:19-39), so
`cfg['env']` from the reference's YAML files is accepted unchanged.  Derived constants are computed
here with the reference's own Python expressions (envs/physics.py:7-11, my_pong_env_2p.py:152,230):
YAML ints (`restitution: 1`) and libm `pow` (`R ** 2`) then behave exactly as in the reference, and
the kernels only ever see doubles.
"""
from __future__ import annotations

from . import _lib

ENV_DEFAULTS = dict(
    render_size=400, paddle_width=0.2, paddle_speed=0.02, max_score=3, enable_render=False,
    enable_spin=True, magnus_factor=0.01, restitution=0.9, friction=0.2, ball_mass=1.0,
    world_ball_radius=0.03, ball_speed_range=(0.01, 0.05), spin_range=(-10, 10),
    ball_angle_intervals=None, speed_scale_every=3, speed_increment=0.2,
)
DEFAULT_ANGLE_INTERVALS = [[-60, -30], [30, 60]]      # my_pong_env_2p.py:56


def resolve_env_config(kwargs: dict) -> dict:
    """Constructor keywords -> full config dict; unknown keys raise TypeError like the reference ctor."""
    unknown = set(kwargs) - set(ENV_DEFAULTS)
    if unknown:
        raise TypeError(f"PongEnv2P.__init__() got an unexpected keyword argument '{sorted(unknown)[0]}'")
    cfg = dict(ENV_DEFAULTS)
    cfg.update(kwargs)
    if not cfg["ball_angle_intervals"]:
        cfg["ball_angle_intervals"] = [list(v) for v in DEFAULT_ANGLE_INTERVALS]
    return cfg


def make_params(cfg: dict) -> _lib.PPParams:
    e, m, R = cfg["restitution"], cfg["ball_mass"], cfg["world_ball_radius"]
    p = _lib.PPParams()
    p.paddle_speed = float(cfg["paddle_speed"])
    p.half_width = cfg["paddle_width"] / 2
    p.magnus_factor = float(cfg["magnus_factor"])
    p.neg_e = float(-e)
    p.m_1pe = float(m * (1 + e))
    p.inertia = float((2 / 5) * m * R ** 2)
    p.two_m_over_7 = float(2 * m / 7.0)
    p.mu = float(cfg["friction"])
    p.mass = float(m)
    p.radius = float(R)
    p.speed_scale = 1.0 + cfg["speed_increment"]
    p.speed_lo, p.speed_hi = (float(v) for v in cfg["ball_speed_range"])
    ang = cfg["ball_angle_intervals"] or DEFAULT_ANGLE_INTERVALS
    for w in range(2):
        p.angle_lo[w], p.angle_hi[w] = float(ang[w][0]), float(ang[w][1])
    p.spin_lo, p.spin_hi = (float(v) for v in cfg["spin_range"])
    p.enable_spin = int(bool(cfg["enable_spin"]))
    p.max_score = int(cfg["max_score"])
    p.speed_scale_every = int(cfg["speed_scale_every"])
    if p.speed_scale_every <= 0 or p.max_score <= 0:
        raise ValueError("speed_scale_every and max_score must be positive")
    return p
