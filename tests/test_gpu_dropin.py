"""The import-compatible shims (`envs/`, `models/` at the repository root) under reference-shaped caller code (`-m gpu`).

The loops below are the reference's own, restated verbatim in shape: eval_vs_model (scripts/train_iterative.py:171-181)
and the arena match loop with select_action_universal (tests/arena.py:199-219,294-308).  They import ONLY the names the
reference imports — `envs.my_pong_env_2p.PongEnv2P`, `envs.physics`, `models.qnet.QNet`, `models.qnet_rnn.QNetRNN` —
and run against the device environment; the oracle port (pure-Python env + torch CPU nets) driven by the same
`random.seed` must produce the same games."""
import random
import struct
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import pong_port
from oracle.policy_torch import QNetPort, QNetRNNPort
import pp_testutil as gu

pytestmark = pytest.mark.gpu


def _shims():
    # a fresh interpreter state for the four names: in the build container an earlier test may have imported the
    # UNMODIFIED reference under the same module names (oracle/ref_shim.py)
    for name in [m for m in sys.modules if m == "envs" or m.startswith("envs.") or m == "models" or m.startswith("models.")]:
        del sys.modules[name]
    from envs.my_pong_env_2p import PongEnv2P
    from envs.physics import collide_sphere_with_moving_plane
    from models.qnet import QNet
    from models.qnet_rnn import QNetRNN
    import pingpong_selfplay_ai_b200 as pp
    assert PongEnv2P is pp.PongEnv2P and QNet is pp.QNet and QNetRNN is pp.QNetRNN
    return PongEnv2P, collide_sphere_with_moving_plane, QNet, QNetRNN


def eval_vs_model(env, A, B, episodes, device):
    """scripts/train_iterative.py:171-181, plus the per-game scores for the comparison."""
    wins, games = 0, []
    for _ in range(episodes):
        oA, oB = env.reset(); done = False
        while not done:
            aA = A(torch.tensor(oA, dtype=torch.float32, device=device).unsqueeze(0)).argmax(1).item()
            aB = B(torch.tensor(oB, dtype=torch.float32, device=device).unsqueeze(0)).argmax(1).item()
            (nA, nB), (rA, rB), done, _ = env.step(aA, aB)
            oA, oB = nA, nB
        if rB > rA: wins += 1
        games.append((env.scoreA, env.scoreB, env.bounce_count))
    return wins / episodes, games


def select_action_universal(obs, model, model_type, hidden_state, device):
    """tests/arena.py:199-219"""
    with torch.no_grad():
        if model_type == "QNetRNN":
            obs_tensor = torch.tensor(obs, dtype=torch.float32, device=device).unsqueeze(0).unsqueeze(0)
            q_values, next_hidden_state = model(obs_tensor, hidden_state)
            return int(q_values.argmax(1).item()), next_hidden_state
        elif model_type == "QNet":
            obs_tensor = torch.tensor(obs, dtype=torch.float32, device=device).unsqueeze(0)
            return int(model(obs_tensor).argmax(1).item()), None
        ball_x, my_paddle_x = float(obs[0]), float(obs[4])          # numpy 1.24.3: float64 arithmetic (requirements.txt:2)
        tolerance = 0.02
        if ball_x < my_paddle_x - tolerance: action = 0
        elif ball_x > my_paddle_x + tolerance: action = 2
        else: action = 1
        return action, None


def arena_match(env, net_A, type_A, net_B, type_B, episodes, device):
    """tests/arena.py:294-308 -> [(scoreA, scoreB)] per game."""
    out = []
    for _ in range(episodes):
        obs_A, obs_B = env.reset()
        done = False
        hidden_A = net_A.init_hidden(1, device) if type_A == "QNetRNN" else None
        hidden_B = net_B.init_hidden(1, device) if type_B == "QNetRNN" else None
        while not done:
            act_A, hidden_A = select_action_universal(obs_A, net_A, type_A, hidden_A, device)
            act_B, hidden_B = select_action_universal(obs_B, net_B, type_B, hidden_B, device)
            (obs_A, obs_B), _, done, _ = env.step(act_A, act_B)
        out.append((env.scoreA, env.scoreB))
    return out


def test_reference_shaped_eval_vs_model_runs_on_the_shims():
    PongEnv2P, _, QNet, _ = _shims()
    cfg = gu.hashes()["env_config_yaml"]
    device = torch.device("cpu")                                     # the reference's device when CUDA torch is absent
    torch.manual_seed(0); A = QNet().to(device); pA = QNetPort().to(device)
    torch.manual_seed(1); B = QNet().to(device); pB = QNetPort().to(device)
    pA.load_state_dict(A.state_dict()); pB.load_state_dict(B.state_dict())
    random.seed(2024)
    env = PongEnv2P(**cfg)
    got = eval_vs_model(env, A, B, 12, device)
    random.seed(2024)
    want = eval_vs_model(pong_port.PongPort(**cfg), pA, pB, 12, device)
    assert got == want and len(got[1]) == 12
    assert isinstance(got[1][0][0], int) and all(max(a, b) == cfg["max_score"] for a, b, _ in got[1])


@pytest.mark.parametrize("types", [("QNetRNN", "QNet"), ("HardcodedBallFollower", "QNetRNN")])
def test_reference_shaped_arena_match_runs_on_the_shims(types):
    PongEnv2P, _, QNet, QNetRNN = _shims()
    cfg = gu.hashes()["env_config_rnn_yaml"]
    device = torch.device("cpu")

    def make(kind, seed):
        torch.manual_seed(seed)
        if kind == "QNet":
            m, p = QNet().eval(), QNetPort().eval()
        elif kind == "QNetRNN":
            m, p = QNetRNN().eval(), QNetRNNPort().eval()
        else:
            return "bot", "bot"
        p.load_state_dict(m.state_dict())
        return m, p
    (a, pa), (b, pb) = make(types[0], 3), make(types[1], 4)
    random.seed(77)
    got = arena_match(PongEnv2P(**cfg), a, types[0], b, types[1], 8, device)
    random.seed(77)
    want = arena_match(pong_port.PongPort(**cfg), pa, types[0], pb, types[1], 8, device)
    assert got == want and len(got) == 8


def test_physics_shim_reproduces_the_collision_known_answers():
    """envs.physics.collide_sphere_with_moving_plane on the device (pp_collide) == the reference's known answers
    (tests/golden/env_hashes.json, generated from the unmodified envs/physics.py), bit for bit."""
    _, collide, _, _ = _shims()
    import pingpong_selfplay_ai_b200 as pp
    H = gu.hashes()
    cfg = H["env_config_yaml"]
    args = (cfg["restitution"], cfg["friction"], cfg["ball_mass"], cfg["world_ball_radius"])
    for kat in H["collision_kat"]:
        inp = [float.fromhex(v) for v in kat["inp"]]
        got = collide(*inp, *args)
        assert isinstance(got, tuple) and all(isinstance(v, float) for v in got)
        assert [struct.pack("<d", v) for v in got] == [struct.pack("<d", float.fromhex(v)) for v in kat["out"]]
    # a batch against the Python restatement, stick / slip / +-0.0 relative velocity included
    rs = np.random.RandomState(5)
    n = 20000
    vn, vt = -rs.uniform(1e-4, 0.3, n), rs.uniform(-0.3, 0.3, n)
    u, om = rs.choice([-0.03, 0.0, 0.03], n), rs.uniform(-60, 60, n)
    vt[:50] = u[:50] + 0.03 * om[:50]                                # vrel = +-0.0 -> copysign(1, +-0.0)
    out = [gu.np_of(t) for t in pp.collide_batch(vn, vt, u, om, *args)]
    want = np.array([pong_port.collide(a, b, c, d, *args) for a, b, c, d in zip(vn, vt, u, om)])
    for k in range(3):
        assert np.array_equal(gu.bits(out[k]), gu.bits(want[:, k]))
