"""Oracle policy nets (C fmaf-chain definition) against torch outputs of the reference models (golden)."""
import os

import numpy as np
import pytest

from oracle import pong_oracle as po


def _sd(g, name):
    pre = name + "/"
    return {k[len(pre):]: v for k, v in g.items() if k.startswith(pre) and not k[len(pre):].startswith(("q_", "h_", "c_"))}


@pytest.fixture(scope="module")
def qg(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "qnet_golden.npz")))


@pytest.mark.parametrize("name", ["seed0", "seed1", "ckpt_model5_1_fault_B"])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_qnet_oracle_matches_torch_reference(qg, name, mode):
    w = po.qnet_weights_from_state_dict(_sd(qg, name), noisy=(mode == "train"))
    q, a = po.qnet_forward(w, qg["obs"])
    ref = qg[f"{name}/q_{mode}"]
    assert np.abs(q - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())      # north_star fp32 tolerance
    # greedy action agrees wherever the reference's top-2 gap is not a rounding-level tie
    srt = np.sort(ref, axis=1)
    clear = (srt[:, 2] - srt[:, 1]) > 1e-5
    assert clear.mean() > 0.95
    assert np.array_equal(a[clear], ref.argmax(1)[clear])


@pytest.mark.parametrize("fname,name", [("qnetrnn_golden.npz", "seed0"), ("qnetrnn_ckpt_golden.npz", "rnn_agent_4_B")])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_qnetrnn_oracle_matches_torch_reference(golden_dir, fname, name, mode):
    g = dict(np.load(os.path.join(golden_dir, fname)))
    w = po.qnetrnn_weights_from_state_dict(_sd(g, name), noisy=(mode == "train"))
    seq = g["seq"]                                     # [B,T,7]
    B, T = seq.shape[:2]
    H = w["Whh"].shape[1]
    h = np.zeros((B, H), np.float32); c = np.zeros((B, H), np.float32)      # init_hidden: zeros (qnet_rnn.py:146-152)
    for t in range(T):
        q, _ = po.qnetrnn_forward(w, seq[:, t], h, c)
        ref = g[f"{name}/q_{mode}"][t]
        assert np.abs(q - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), t
    assert np.abs(h - g[f"{name}/h_{mode}"]).max() < 1e-5
    assert np.abs(c - g[f"{name}/c_{mode}"]).max() < 1e-5 * max(1.0, np.abs(g[f"{name}/c_{mode}"]).max())


def test_philox4x32_10_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    assert [hex(v) for v in po.philox4x32(0, 0, 0, 0, 0, 0)] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    f = 0xFFFFFFFF
    assert [hex(v) for v in po.philox4x32(f, f, f, f, f, f)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    got = po.philox4x32(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0)
    assert [hex(v) for v in got] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_philox_serve_distribution_matches_reference_formula():
    from oracle.pong_port import CONFIG_YAML_ENV as cfg
    s = np.array([po.philox_serve(7, i, 0, cfg) for i in range(4000)])
    speed = np.hypot(s[:, 0], s[:, 1])
    ang = np.degrees(np.arctan2(s[:, 1], s[:, 0]))
    assert speed.min() >= 0.03 - 1e-12 and speed.max() <= 0.05 + 1e-12
    assert np.all((np.abs(ang) >= 30 - 1e-9) & (np.abs(ang) <= 60 + 1e-9)) and np.all(s[:, 0] > 0)
    assert 0.45 < (ang > 0).mean() < 0.55
    assert s[:, 2].min() >= -5 and s[:, 2].max() <= 5 and abs(s[:, 2].mean()) < 0.2
    assert abs(speed.mean() - 0.04) < 5e-4
