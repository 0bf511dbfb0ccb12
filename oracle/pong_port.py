"""CPU restatement (pure Python, scalar fp64) of the reference PongEnv2P.

TEST INFRASTRUCTURE — only `tests/`, `oracle/gen_golden.py`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` leg may import this.  It is the
checker (and the timed CPU baseline "port"), never a product path: the product in
`pingpong_selfplay_ai_b200/` fails loudly when its CUDA library is missing.

Parity status: PINNED in the build container against the unmodified reference
(`tests/test_oracle_vs_reference.py`, bit-exact over >1e6 env-steps incl. forced quirk
cases) and everywhere against `tests/golden/*` generated from the reference by
`oracle/gen_golden.py`.  The reference's own repository holds no tests or golden
vectors for this path, so those generated vectors are the pin.

What is restated (reference file:line, all under /root/reference):
  * constructor defaults / parameter names .... envs/my_pong_env_2p.py:19-39
  * reset: 4 draws speed, coin, angle, spin .... envs/my_pong_env_2p.py:83-114
  * step: paddles, Magnus, advance, X walls,
    top / bottom paddle or score, done ......... envs/my_pong_env_2p.py:116-225
  * speed scaling every n-th paddle hit ......... envs/my_pong_env_2p.py:227-232
  * the two 7-D fp32 observations .............. envs/my_pong_env_2p.py:235-263
  * sphere vs moving plane impulse ............. envs/physics.py:3-23

Every arithmetic expression keeps the reference's association order; CPython evaluates
each binary op as one IEEE-754 double operation (no FMA), which is what the C oracle
(`pong_oracle.c`, built with -ffp-contract=off) and the CUDA kernels (`__dadd_rn` /
`__dmul_rn` / `__ddiv_rn`) reproduce.
"""
from __future__ import annotations

import math
import random as _global_random

import numpy as np

ENV_DEFAULTS = dict(  # envs/my_pong_env_2p.py:19-39
    render_size=400, paddle_width=0.2, paddle_speed=0.02, max_score=3, enable_render=False,
    enable_spin=True, magnus_factor=0.01, restitution=0.9, friction=0.2, ball_mass=1.0,
    world_ball_radius=0.03, ball_speed_range=(0.01, 0.05), spin_range=(-10, 10),
    ball_angle_intervals=None, speed_scale_every=3, speed_increment=0.2,
)

# config.yaml:1-17 and config_rnn.yaml:6-28 of the reference (values only; the two files
# differ in speed_scale_every / speed_increment, and restitution 1 vs 1.0).
CONFIG_YAML_ENV = dict(
    render_size=400, paddle_width=0.2, paddle_speed=0.03, max_score=3, enable_render=False,
    enable_spin=True, magnus_factor=0.025, restitution=1, friction=0.6, ball_mass=1.0,
    world_ball_radius=0.03, ball_speed_range=[0.03, 0.05], spin_range=[-5, 5],
    ball_angle_intervals=[[-60, -30], [30, 60]], speed_scale_every=1, speed_increment=0.1,
)
CONFIG_RNN_YAML_ENV = dict(CONFIG_YAML_ENV, restitution=1.0, speed_scale_every=5, speed_increment=0.2)

# Constructor keyword sets beyond the two YAML blocks (the reference accepts all of them, my_pong_env_2p.py:19-39):
# parity is also pinned on these through tests/golden/env_extra_cfgs.npz (oracle/gen_golden.py, from the reference).
EXTRA_ENV_CONFIGS = [
    dict(),                                                                    # the constructor defaults
    dict(enable_spin=False, restitution=0.9, friction=0.2, max_score=1),
    dict(paddle_width=0.35, paddle_speed=0.07, max_score=5, speed_scale_every=2, speed_increment=0.35,
         ball_speed_range=(0.02, 0.12), spin_range=(-25, 25), magnus_factor=0.06, world_ball_radius=0.05, ball_mass=2.5),
    dict(restitution=1, friction=0.0, speed_scale_every=1, speed_increment=0.0, ball_angle_intervals=[[-80, -10], [10, 80]]),
    dict(restitution=0.5, friction=1.5, paddle_width=0.05, ball_speed_range=(0.2, 0.6), max_score=2),   # very fast balls
]


def collide(vn, vt, u, omega, e, mu, m, R):
    """envs/physics.py:3-23 — returns (vn', vt', omega')."""
    vn_out = -e * vn                                   # :7   (-e)*vn
    j_n = m * (1 + e) * abs(vn)                        # :8   (m*(1+e))*|vn|
    inertia = (2 / 5) * m * R ** 2                     # :9   ((2/5)*m)*R**2
    j_t = (2 * m / 7.0) * (u + R * omega - vt)         # :10  ((2m)/7)*((u+R*w)-vt)
    cap = mu * j_n                                     # :11
    if not (abs(j_t) <= cap):                          # :13-18 slip: Coulomb cap, sign of vrel
        v_rel = (vt - u) - R * omega
        j_t = -cap * math.copysign(1, v_rel)
    return vn_out, vt + (j_t / m), omega - (R * j_t) / inertia   # :20-23


def draw_serve(rng, speed_range, angle_intervals, spin_range):
    """The 4 Mersenne-Twister draws of reset() in reference order (:98-111) -> (vx, vy, spin)."""
    speed = rng.uniform(*speed_range)
    which = 0 if rng.random() < 0.5 else 1
    rad = math.radians(rng.uniform(*angle_intervals[which]))
    vx = speed * math.cos(rad)
    vy = speed * math.sin(rad)
    spin = rng.uniform(*spin_range)
    return vx, vy, spin


class PongPort:
    """Scalar fp64 restatement with the reference's public surface (reset/step/attrs)."""

    def __init__(self, rng=None, **kw):
        unknown = set(kw) - set(ENV_DEFAULTS)
        if unknown:
            raise TypeError(f"unexpected env parameters: {sorted(unknown)}")
        p = dict(ENV_DEFAULTS)
        p.update(kw)
        if not p["ball_angle_intervals"]:
            p["ball_angle_intervals"] = [[-60, -30], [30, 60]]      # :58
        for k, v in p.items():
            setattr(self, k, v)
        self._rng = rng if rng is not None else _global_random      # reference uses the global module
        self.bounce_count = 0
        self.reset()                                                # :81 (consumes 4 draws)

    # ------------------------------------------------------------------ reset
    def reset(self, seed=None, options=None):
        self.scoreA = self.scoreB = 0
        self.bounce_count = 0
        self.top_paddle_x = self.bottom_paddle_x = 0.5
        self.ball_x = self.ball_y = 0.5
        self.ball_vx, self.ball_vy, self.spin = draw_serve(
            self._rng, self.ball_speed_range, self.ball_angle_intervals, self.spin_range)
        return self.observe()

    def serve(self, vx, vy, spin):
        """reset() with an injected serve instead of RNG draws (used by parity harnesses)."""
        self.scoreA = self.scoreB = 0
        self.bounce_count = 0
        self.top_paddle_x = self.bottom_paddle_x = 0.5
        self.ball_x = self.ball_y = 0.5
        self.ball_vx, self.ball_vy, self.spin = vx, vy, spin
        return self.observe()

    # ------------------------------------------------------------------- step
    @staticmethod
    def _move(pos, action, speed):
        if action == 0:
            pos = pos - speed
        elif action == 2:
            pos = pos + speed
        return min(max(pos, 0.0), 1.0)          # np.clip(pos, 0, 1) for finite pos (:122,:128)

    def _paddle_event(self, paddle_x, action, top_side):
        """Ball crossed y<0 (top_side) or y>1: returns True on a hit (state updated), False on a miss."""
        half = self.paddle_width / 2
        if not (paddle_x - half <= self.ball_x <= paddle_x + half):
            return False
        u = -self.paddle_speed if action == 0 else (self.paddle_speed if action == 2 else 0.0)
        vn_in = self.ball_vy if top_side else -self.ball_vy
        vn, vt, om = collide(vn_in, self.ball_vx, u, self.spin, self.restitution, self.friction,
                             self.ball_mass, self.world_ball_radius)
        self.ball_vy = vn if top_side else -vn
        self.ball_vx = vt
        self.spin = om
        self.ball_y = 0.0 if top_side else 1.0
        self.bounce_count += 1
        if self.bounce_count % self.speed_scale_every == 0:        # :227-232
            s = 1.0 + self.speed_increment
            self.ball_vx *= s
            self.ball_vy *= s
        return True

    def step(self, actionA, actionB):
        self.top_paddle_x = self._move(self.top_paddle_x, actionA, self.paddle_speed)
        self.bottom_paddle_x = self._move(self.bottom_paddle_x, actionB, self.paddle_speed)
        if self.enable_spin:                                        # :135-136
            self.ball_vx = self.ball_vx + self.magnus_factor * self.spin * self.ball_vy
        self.ball_x = self.ball_x + self.ball_vx                    # :139-140
        self.ball_y = self.ball_y + self.ball_vy
        if self.ball_x < 0:                                         # :143-148
            self.ball_x = -self.ball_x
            self.ball_vx = -self.ball_vx
        elif self.ball_x > 1:
            self.ball_x = 2 - self.ball_x
            self.ball_vx = -self.ball_vx
        rA = rB = 0.0
        done = False
        if self.ball_y < 0:                                         # :151-186
            if not self._paddle_event(self.top_paddle_x, actionA, True):
                rA, rB = -1.0, 1.0
                self.scoreB += 1
                done = self.scoreB >= self.max_score
        elif self.ball_y > 1:                                       # :189-223
            if not self._paddle_event(self.bottom_paddle_x, actionB, False):
                rA, rB = 1.0, -1.0
                self.scoreA += 1
                done = self.scoreA >= self.max_score
        return self.observe(), (rA, rB), done, {}

    # -------------------------------------------------------------------- obs
    def observe(self):
        a = np.array([self.ball_x, 1.0 - self.ball_y, self.ball_vx, -self.ball_vy,
                      self.top_paddle_x, self.bottom_paddle_x, self.spin], dtype=np.float32)
        b = np.array([self.ball_x, self.ball_y, self.ball_vx, self.ball_vy,
                      self.bottom_paddle_x, self.top_paddle_x, self.spin], dtype=np.float32)
        return a, b

    def state_tuple(self):
        return (float(self.ball_x), float(self.ball_y), float(self.ball_vx), float(self.ball_vy),
                float(self.spin), float(self.top_paddle_x), float(self.bottom_paddle_x),
                int(self.scoreA), int(self.scoreB), int(self.bounce_count))

    def close(self):
        pass
