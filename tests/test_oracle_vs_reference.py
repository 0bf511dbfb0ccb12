"""Oracle against the LIVE unmodified reference (build container only: needs /root/reference)."""
import random

import numpy as np
import pytest

from oracle import pong_oracle as po
from oracle import pong_port, ref_shim

pytestmark = pytest.mark.reference


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


@pytest.mark.parametrize("cfgname", ["config.yaml", "config_rnn.yaml"])
def test_port_and_c_oracle_bit_exact_over_a_million_steps(cfgname):
    """>= 1e6 env-steps: reference vs python port vs C oracle, identical serves and actions, every step."""
    PongEnv2P, _, _, _ = ref_shim.load_reference()
    cfg = ref_shim.load_reference_config(cfgname)["env"]
    steps = 520_000
    random.seed(99)
    ref = PongEnv2P(**cfg)
    port = pong_port.PongPort(rng=random.Random(0), **cfg)
    p = po.make_params(cfg)
    arng = np.random.RandomState(5)
    acts = arng.randint(0, 3, size=(steps, 2)).astype(np.uint8)
    # phase 1: reference vs port, in lock step, collecting the reference's serves and states
    ref.reset()
    serves = [(float(ref.ball_vx), float(ref.ball_vy), float(ref.spin))]
    port.serve(*serves[0])
    states = np.zeros((steps, 7)); ints = np.zeros((steps, 3), np.int32); dones = np.zeros(steps, np.uint8)
    for t in range(steps):
        a, b = int(acts[t, 0]), int(acts[t, 1])
        (ra_o, rb_o), rr, dr, _ = ref.step(a, b)
        (pa_o, pb_o), pr, dp, _ = port.step(a, b)
        st = port.state_tuple()
        assert st == (float(ref.ball_x), float(ref.ball_y), float(ref.ball_vx), float(ref.ball_vy), float(ref.spin),
                      float(ref.top_paddle_x), float(ref.bottom_paddle_x), ref.scoreA, ref.scoreB, ref.bounce_count), t
        assert rr == pr and dr == dp and ra_o.tobytes() == pa_o.tobytes() and rb_o.tobytes() == pb_o.tobytes(), t
        states[t] = st[:7]; ints[t] = st[7:]; dones[t] = dr
        if dr:
            ref.reset()
            serves.append((float(ref.ball_vx), float(ref.ball_vy), float(ref.spin)))
            port.serve(*serves[-1])
    # phase 2: the C rollout replays the same thing in one call
    sv = np.asarray(serves)
    b = po.EnvBatch(1, "f64"); b.serve(*sv[0])
    out = po.rollout(p, b, acts.reshape(steps, 1, 2), tuple(sv[:, i].reshape(-1, 1) for i in range(3)), trace=True)
    assert np.array_equal(_bits(out["trace_real"][:, :, 0]), _bits(states))
    assert np.array_equal(out["trace_int"][:, :3, 0], ints)
    assert np.array_equal(out["trace_int"][:, 3, 0] & 1, dones)
    assert dones.sum() > 5000


def test_serve_draws_match_reference_reset():
    PongEnv2P, _, _, _ = ref_shim.load_reference()
    cfg = ref_shim.load_reference_config("config.yaml")["env"]
    random.seed(4242)
    env = PongEnv2P(**cfg)                     # constructor's reset() consumes the first 4 draws
    got = [(env.ball_vx, env.ball_vy, env.spin)]
    for _ in range(63):
        env.reset(); got.append((env.ball_vx, env.ball_vy, env.spin))
    vx, vy, sp = po.serve_pool_from_reference_rng(4242, 8, 8, cfg)   # env-major order
    want = [(vx[j, i], vy[j, i], sp[j, i]) for i in range(8) for j in range(8)]
    assert got == want


def test_qnet_torch_port_state_dict_is_interchangeable():
    """oracle/policy_torch.py mirrors the reference modules key-for-key (used by the CPU baseline loop)."""
    import torch
    from oracle import policy_torch
    _, _, QNet, QNetRNN = ref_shim.load_reference()
    torch.manual_seed(3)
    ref = QNet(7, 3); mine = policy_torch.QNetPort()
    mine.load_state_dict(ref.state_dict())
    x = torch.randn(33, 7)
    for train in (False, True):
        ref.train(train); mine.train(train)
        assert torch.equal(ref(x), mine(x))
    refr = QNetRNN(); miner = policy_torch.QNetRNNPort()
    miner.load_state_dict(refr.state_dict())
    xs = torch.randn(5, 1, 7); hc = refr.init_hidden(5, "cpu")
    refr.eval(); miner.eval()
    q1, hc1 = refr(xs, hc); q2, hc2 = miner(xs, miner.init_hidden(5, "cpu"))
    assert torch.equal(q1, q2) and torch.equal(hc1[0], hc2[0]) and torch.equal(hc1[1], hc2[1])
