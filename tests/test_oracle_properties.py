"""Property tests of the oracle (hypothesis, CPU only, no reference tree needed): the C restatement and the pure-Python
port must agree bit for bit on RANDOM constructor keyword sets, serves and action streams — not only on the few configs
the golden trajectories cover — and both must keep the invariants the reference's step() has by construction
(envs/my_pong_env_2p.py:116-225)."""
import random

import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import pong_oracle as po
from oracle.pong_port import ENV_DEFAULTS, PongPort, draw_serve

cfgs = st.fixed_dictionaries({
    "paddle_width": st.sampled_from([0.05, 0.2, 0.35, 0.9]),
    "paddle_speed": st.sampled_from([0.01, 0.03, 0.08]),
    "max_score": st.integers(1, 5),
    "enable_spin": st.booleans(),
    "magnus_factor": st.sampled_from([0.0, 0.025, 0.1]),
    "restitution": st.sampled_from([0.5, 0.9, 1.0, 1.1]),
    "friction": st.sampled_from([0.0, 0.3, 0.6, 1.5]),
    "ball_mass": st.sampled_from([0.5, 1.0, 2.0]),
    "world_ball_radius": st.sampled_from([0.01, 0.03, 0.05]),
    "speed_scale_every": st.integers(1, 6),
    "speed_increment": st.sampled_from([0.0, 0.1, 0.2, 0.5]),
})


def _bits(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64)).view(np.uint64)


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(cfg_kw=cfgs, seed=st.integers(0, 2 ** 31 - 1), out_of_range=st.booleans())
def test_c_oracle_equals_python_port_on_random_configs(cfg_kw, seed, out_of_range):
    cfg = dict(ENV_DEFAULTS, **cfg_kw)
    if not cfg["ball_angle_intervals"]:
        cfg["ball_angle_intervals"] = [[-60, -30], [30, 60]]
    rng = random.Random(seed)
    port = PongPort(rng=random.Random(seed ^ 0x5a5a), **cfg_kw)
    p = po.make_params(cfg)
    b = po.EnvBatch(1, "f64")
    serve = draw_serve(rng, cfg["ball_speed_range"], cfg["ball_angle_intervals"], cfg["spin_range"])
    port.serve(*serve)
    b.serve(np.array([serve[0]]), np.array([serve[1]]), np.array([serve[2]]))
    hi = 5 if out_of_range else 2                         # values outside {0, 1, 2} mean "stay" (:118-128)
    prev_scores, episodes = (0, 0), 0
    for t in range(400):
        aa, ab = rng.randint(0, hi), rng.randint(0, hi)
        (oa_p, ob_p), (ra_p, rb_p), done_p, _ = port.step(aa, ab)
        oa, ob, ra, rb, done = po.step(p, b, np.array([aa], np.uint8), np.array([ab], np.uint8))
        state_p = (port.ball_x, port.ball_y, port.ball_vx, port.ball_vy, port.spin, port.top_paddle_x, port.bottom_paddle_x)
        sr, si = b.state_matrix()
        assert np.array_equal(_bits(state_p), _bits(sr[:, 0])), t
        assert tuple(int(v) for v in si[:, 0]) == (port.scoreA, port.scoreB, port.bounce_count)
        assert np.array_equal(np.asarray(oa_p, np.float32).view(np.uint32), oa[0].view(np.uint32))
        assert np.array_equal(np.asarray(ob_p, np.float32).view(np.uint32), ob[0].view(np.uint32))
        assert (float(ra[0]), float(rb[0]), bool(done[0])) == (float(ra_p), float(rb_p), bool(done_p))
        # invariants of the reference's step()
        assert 0.0 <= port.top_paddle_x <= 1.0 and 0.0 <= port.bottom_paddle_x <= 1.0          # np.clip :123,128
        assert ra_p == -rb_p and ra_p in (-1.0, 0.0, 1.0)                                       # zero-sum rewards
        gained = (port.scoreA - prev_scores[0], port.scoreB - prev_scores[1])
        assert gained == ((1, 0) if ra_p > 0 else (0, 1) if ra_p < 0 else (0, 0))               # a reward is a point
        assert bool(done_p) == (max(port.scoreA, port.scoreB) >= cfg["max_score"] and ra_p != 0)   # :180-186,216-223
        # observations are the two mirrored views of one state (:235-257)
        assert oa_p[0] == ob_p[0] and oa_p[2] == ob_p[2] and oa_p[6] == ob_p[6]
        assert oa_p[3] == -ob_p[3] and oa_p[4] == ob_p[5] and oa_p[5] == ob_p[4]
        prev_scores = (port.scoreA, port.scoreB)
        if done_p:                                            # no re-serve inside an episode; reset() only on done
            episodes += 1
            serve = draw_serve(rng, cfg["ball_speed_range"], cfg["ball_angle_intervals"], cfg["spin_range"])
            port.serve(*serve)
            b.serve(np.array([serve[0]]), np.array([serve[1]]), np.array([serve[2]]))
            prev_scores = (0, 0)
    assert episodes >= 0
