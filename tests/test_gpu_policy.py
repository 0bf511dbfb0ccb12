"""K2a / K2b parity (`-m gpu`): device action selection against the oracle's fmaf-chain definition (bit-exact for
QNet) and against the torch outputs of the reference modules stored in tests/golden (1e-5, north_star fp32)."""
import os

import numpy as np
import pytest
import torch

import pingpong_selfplay_ai_b200 as pp
from oracle import pong_oracle as po
import pp_testutil as gu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qg():
    return dict(np.load(os.path.join(gu.GOLDEN, "qnet_golden.npz")))


@pytest.mark.parametrize("name", ["seed0", "seed1", "ckpt_model5_1_fault_B"])
@pytest.mark.parametrize("noisy", [False, True])
def test_qnet_act_matches_oracle_bitwise_and_torch_reference(qg, name, noisy):
    sd = gu.golden_sd(qg, name)
    pol = pp.Policy.qnet(sd, noisy=noisy)
    obs = torch.from_numpy(qg["obs"]).cuda()
    act, q = pp.qnet_act(obs, pol, want_q=True)
    wq, wa = po.qnet_forward(po.qnet_weights_from_state_dict(sd, noisy=noisy), qg["obs"])
    assert np.array_equal(gu.bits(gu.np_of(q)), gu.bits(wq))                   # same fmaf chain -> same bits
    assert np.array_equal(gu.np_of(act), wa)
    ref = qg[f"{name}/q_{'train' if noisy else 'eval'}"]
    assert np.abs(gu.np_of(q) - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("n", [1, 127, 128, 129, 5000])
def test_qnet_act_ragged_sizes(qg, n):
    sd = gu.golden_sd(qg, "seed1")
    pol = pp.Policy.qnet(sd)
    rs = np.random.RandomState(n)
    obs = rs.uniform(-1, 1, size=(n, 7)).astype(np.float32)
    obs[:, 6] *= 5
    act, q = pp.qnet_act(torch.from_numpy(obs).cuda(), pol, want_q=True)
    wq, wa = po.qnet_forward(po.qnet_weights_from_state_dict(sd), obs)
    assert np.array_equal(gu.bits(gu.np_of(q)), gu.bits(wq)) and np.array_equal(gu.np_of(act), wa)


def test_follower_random_and_epsilon_greedy_match_oracle_rng(qg):
    """HardcodedBallFollower (tests/arena.py:211-217), the uniform random player and the epsilon-greedy overlay
    (scripts/train_iterative.py:124-130) use Philox keyed by (seed; env id, step, player): same actions as the
    oracle's restatement of that generator."""
    n, seed, step, base = 3000, 0xDEADBEEF12345, 17, 1000
    rs = np.random.RandomState(3)
    obs = rs.uniform(0, 1, size=(n, 7)).astype(np.float32)
    obs[::5, 0] = obs[::5, 4]                                                  # inside the tolerance band
    dobs = torch.from_numpy(obs).cuda()
    a, _ = pp.qnet_act(dobs, pp.Policy.follower(tol=0.02), seed=seed, step_index=step, env_id_base=base)
    obs[1::5, 0] = obs[1::5, 4] - np.float32(0.02)                             # on the band's edge: float32 and float64
    obs[2::5, 0] = obs[2::5, 4] + np.float32(0.02)                             # arithmetic decide differently here
    dobs = torch.from_numpy(obs).cuda()
    a, _ = pp.qnet_act(dobs, pp.Policy.follower(tol=0.02), seed=seed, step_index=step, env_id_base=base)
    # the reference pins numpy 1.24.3: `np.float32 - 0.02` and the compares are float64 (tests/arena.py:211-217)
    x, pad = obs[:, 0].astype(np.float64), obs[:, 4].astype(np.float64)
    want = np.where(x < pad - 0.02, 0, np.where(x > pad + 0.02, 2, 1))
    f32 = np.where(obs[:, 0] < obs[:, 4] - np.float32(0.02), 0, np.where(obs[:, 0] > obs[:, 4] + np.float32(0.02), 2, 1))
    assert (want != f32).any()                                                 # the edge cases do tell the two apart
    assert np.array_equal(gu.np_of(a), want) and len(set(want.tolist())) == 3
    for stream in (1, 2):
        a, _ = pp.qnet_act(dobs, pp.Policy.random(), seed=seed, step_index=step, env_id_base=base, stream_id=stream)
        r = np.array([po.philox4x32(base + i, step, stream, 0, seed & 0xFFFFFFFF, seed >> 32) for i in range(n)])
        assert np.array_equal(gu.np_of(a), (r[:, 1].astype(np.uint64) * 3) >> 32)
    sd = gu.golden_sd(qg, "seed0")
    greedy, _ = pp.qnet_act(dobs, pp.Policy.qnet(sd), seed=seed, step_index=step, env_id_base=base, stream_id=2)
    eg, _ = pp.qnet_act(dobs, pp.Policy.qnet(sd, eps=0.3), seed=seed, step_index=step, env_id_base=base, stream_id=2)
    r = np.array([po.philox4x32(base + i, step, 2, 0, seed & 0xFFFFFFFF, seed >> 32) for i in range(n)])
    explore = r[:, 0].astype(np.uint64) < po.eps_threshold(0.3)
    want = np.where(explore, (r[:, 1].astype(np.uint64) * 3) >> 32, gu.np_of(greedy))
    assert np.array_equal(gu.np_of(eg), want) and 0.25 < explore.mean() < 0.35


@pytest.mark.parametrize("fname,name", [("qnetrnn_golden.npz", "seed0"), ("qnetrnn_ckpt_golden.npz", "rnn_agent_4_B")])
@pytest.mark.parametrize("noisy", [False, True])
def test_qnetrnn_act_carried_state_vs_oracle_and_torch_reference(fname, name, noisy):
    """12 carried steps for 48 envs (B = 48 exercises a partial 64-env tile): Q within 1e-5 of the torch reference
    and within 2e-6 of the oracle chain (expf/tanhf differ in the last ulps), (h, c) after the last step too."""
    g = dict(np.load(os.path.join(gu.GOLDEN, fname)))
    sd = gu.golden_sd(g, name)
    seq = g["seq"]
    B, T = seq.shape[:2]
    pol = pp.Policy.qnetrnn(sd, num_envs=B, noisy=noisy)
    w = po.qnetrnn_weights_from_state_dict(sd, noisy=noisy)
    h = np.zeros((B, 128), np.float32); c = np.zeros((B, 128), np.float32)
    mode = "train" if noisy else "eval"
    pol.h.fill_(7.0); pol.c.fill_(-3.0)                                         # reset_mask must zero these at t = 0
    for t in range(T):
        mask = torch.ones(B, dtype=torch.uint8, device="cuda") if t == 0 else None
        act, q = pp.qnetrnn_act(torch.from_numpy(seq[:, t].copy()).cuda(), pol, reset_mask=mask, want_q=True)
        wq, wa = po.qnetrnn_forward(w, seq[:, t], h, c)
        ref = g[f"{name}/q_{mode}"][t]
        assert np.abs(gu.np_of(q) - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), t
        assert np.abs(gu.np_of(q) - wq).max() <= 2e-6 * max(1.0, np.abs(wq).max()), t
        srt = np.sort(wq, axis=1)
        clear = (srt[:, 2] - srt[:, 1]) > 1e-5
        assert np.array_equal(gu.np_of(act)[clear], wa[clear]), t
    assert np.abs(gu.np_of(pol.h).T - g[f"{name}/h_{mode}"]).max() < 1e-5
    assert np.abs(gu.np_of(pol.c).T - g[f"{name}/c_{mode}"]).max() < 1e-5 * max(1.0, np.abs(g[f"{name}/c_{mode}"]).max())


def test_qnetrnn_partial_reset_mask_and_larger_batch():
    torch.manual_seed(4)
    net = pp.QNetRNN()
    n = 200
    pol = pp.Policy.qnetrnn(net, num_envs=n)
    w = po.qnetrnn_weights_from_state_dict(net.state_dict())
    rs = np.random.RandomState(0)
    h = np.zeros((n, 128), np.float32); c = np.zeros((n, 128), np.float32)
    for t in range(5):
        obs = rs.uniform(-1, 1, size=(n, 7)).astype(np.float32)
        mask = (rs.rand(n) < 0.3).astype(np.uint8) if t else np.ones(n, np.uint8)
        h[mask.astype(bool)] = 0; c[mask.astype(bool)] = 0
        _, q = pp.qnetrnn_act(torch.from_numpy(obs).cuda(), pol, reset_mask=torch.from_numpy(mask).cuda(), want_q=True)
        wq, _ = po.qnetrnn_forward(w, obs, h, c)
        assert np.abs(gu.np_of(q) - wq).max() <= 2e-6 * max(1.0, np.abs(wq).max()), t
    assert np.abs(gu.np_of(pol.h).T - h).max() < 2e-6 and np.abs(gu.np_of(pol.c).T - c).max() < 4e-6


# ------------------------------------------------------------------------------------------ tensor-core path
@pytest.mark.parametrize("name", ["seed0", "seed1", "ckpt_model5_1_fault_B"])
@pytest.mark.parametrize("noisy", [False, True])
def test_qnet_act_tensor_core_path_within_1e3_of_reference(qg, name, noisy):
    """PP_PREC_F16 (tcgen05, fp16 operands, fp32 accumulation): Q within 1e-3 of the torch reference (north_star
    tolerance for the reduced-precision path) and of the fp32 oracle; greedy action equal wherever the reference's
    top-2 gap exceeds the tolerance."""
    sd = gu.golden_sd(qg, name)
    pol = pp.Policy.qnet(sd, noisy=noisy, precision="f16")
    obs = torch.from_numpy(qg["obs"]).cuda()
    act, q = pp.qnet_act(obs, pol, want_q=True)
    q, act = gu.np_of(q), gu.np_of(act)
    ref = qg[f"{name}/q_{'train' if noisy else 'eval'}"]
    wq, wa = po.qnet_forward(po.qnet_weights_from_state_dict(sd, noisy=noisy), qg["obs"])
    err = np.abs(q - ref).max()
    print(f"tensor-core QNet {name} noisy={noisy}: max |dQ| = {err:.3e} (|Q| max {np.abs(ref).max():.2f})")
    assert err <= 1e-3 and np.abs(q - wq).max() <= 1e-3
    srt = np.sort(ref, axis=1)
    clear = (srt[:, 2] - srt[:, 1]) > 2e-3
    assert clear.mean() > 0.9 and np.array_equal(act[clear], ref.argmax(1)[clear])
    assert (act == wa).mean() > 0.99


@pytest.mark.parametrize("n", [1, 127, 129, 1000, 70001])
def test_qnet_act_tensor_core_ragged_sizes_and_large_inputs(qg, n):
    sd = gu.golden_sd(qg, "seed1")
    pol = pp.Policy.qnet(sd, precision="f16")
    rs = np.random.RandomState(n)
    obs = rs.uniform(-1, 1, size=(n, 7)).astype(np.float32)
    obs[:, 6] *= 40                                            # spin after hard hits
    if n >= 1000:
        obs[:7] = 0; obs[:7, :][np.arange(7), np.arange(7)] = 3e5        # beyond fp16 range: saturates, stays finite
    act, q = pp.qnet_act(torch.from_numpy(obs).cuda(), pol, want_q=True)
    q = gu.np_of(q)
    wq, wa = po.qnet_forward(po.qnet_weights_from_state_dict(sd), obs)
    assert np.isfinite(q).all()
    lo = 7 if n >= 1000 else 0
    assert np.abs(q[lo:] - wq[lo:]).max() <= 1e-3 * max(1.0, np.abs(wq[lo:]).max())


@pytest.mark.parametrize("fname,name", [("qnetrnn_golden.npz", "seed0"), ("qnetrnn_ckpt_golden.npz", "rnn_agent_4_B")])
@pytest.mark.parametrize("noisy", [False, True])
def test_qnetrnn_act_tensor_core_path_carried_state(fname, name, noisy):
    """PP_PREC_F16 QNetRNN (every layer on tcgen05, fp16 hi/lo activations, fp32 accumulation): 12 carried steps for
    48 envs, Q within 1e-3 of the torch reference, (h, c) within 1e-3, reset mask honoured."""
    g = dict(np.load(os.path.join(gu.GOLDEN, fname)))
    sd = gu.golden_sd(g, name)
    seq = g["seq"]
    B, T = seq.shape[:2]
    pol = pp.Policy.qnetrnn(sd, num_envs=B, noisy=noisy, precision="f16")
    mode = "train" if noisy else "eval"
    pol.h.fill_(7.0); pol.c.fill_(-3.0)
    worst = 0.0
    for t in range(T):
        mask = torch.ones(B, dtype=torch.uint8, device="cuda") if t == 0 else None
        act, q = pp.qnetrnn_act(torch.from_numpy(seq[:, t].copy()).cuda(), pol, reset_mask=mask, want_q=True)
        ref = g[f"{name}/q_{mode}"][t]
        err = np.abs(gu.np_of(q) - ref).max()
        worst = max(worst, err / max(1.0, np.abs(ref).max()))
        assert err <= 1e-3 * max(1.0, np.abs(ref).max()), (t, err)
        srt = np.sort(ref, axis=1)
        clear = (srt[:, 2] - srt[:, 1]) > 2e-3
        assert np.array_equal(gu.np_of(act)[clear], ref.argmax(1)[clear]), t
    print(f"tensor-core QNetRNN {name} noisy={noisy}: worst relative |dQ| = {worst:.3e}")
    h_dev, c_dev = (gu.np_of(t) for t in pol.hidden())
    assert np.abs(h_dev - g[f"{name}/h_{mode}"]).max() < 1e-3
    assert np.abs(c_dev - g[f"{name}/c_{mode}"]).max() < 1e-3 * max(1.0, np.abs(g[f"{name}/c_{mode}"]).max())


def test_qnetrnn_act_tensor_core_many_tiles_and_partial_reset():
    torch.manual_seed(4)
    net = pp.QNetRNN()
    n = 128 * 150 + 37                                          # more tiles than SMs, ragged last tile
    pol = pp.Policy.qnetrnn(net, num_envs=n, precision="f16")
    w = po.qnetrnn_weights_from_state_dict(net.state_dict())
    rs = np.random.RandomState(0)
    h = np.zeros((n, 128), np.float32); c = np.zeros((n, 128), np.float32)
    for t in range(3):
        obs = rs.uniform(-1, 1, size=(n, 7)).astype(np.float32)
        mask = (rs.rand(n) < 0.3).astype(np.uint8) if t else np.ones(n, np.uint8)
        h[mask.astype(bool)] = 0; c[mask.astype(bool)] = 0
        _, q = pp.qnetrnn_act(torch.from_numpy(obs).cuda(), pol, reset_mask=torch.from_numpy(mask).cuda(), want_q=True)
        sub = slice(0, None, 97)                                 # the serial oracle on a sample of the envs
        hs, cs = h[sub].copy(), c[sub].copy()
        wq, _ = po.qnetrnn_forward(w, obs[sub], hs, cs)
        assert np.abs(gu.np_of(q)[sub] - wq).max() <= 1e-3, t
        h[:], c[:] = (gu.np_of(t) for t in pol.hidden())         # carry the device state forward for the next step
        assert np.abs(h[sub] - hs).max() < 1e-3 and np.abs(c[sub] - cs).max() < 1e-3
