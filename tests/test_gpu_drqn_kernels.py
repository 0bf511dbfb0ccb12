"""The hand-written DRQN update (csrc/drqn_kernels.cu: pp_drqn_grads, pp_clip_grad_norm, pp_adam_step_multi) against
torch autograd on the CPU (`-m gpu`): train_step_rnn of scripts/train_rnn_iterative.py:400-531 restated in
oracle/train_port.py, run on the reference's own module (torch port, state_dict-compatible) in fp32."""
import copy
import os

import numpy as np
import pytest
import torch

import pingpong_selfplay_ai_b200 as pp
from pingpong_selfplay_ai_b200.train_rnn import DRQNTrainer, SequenceSampler
from oracle.policy_torch import QNetRNNPort
from oracle.train_port import TorchDRQNTrainer, drqn_loss
import pp_testutil as gu

pytestmark = pytest.mark.gpu


def _net(which):
    if which == "ckpt":                                  # the reference's trained checkpoint (large head weights, saturated gates)
        g = dict(np.load(os.path.join(gu.GOLDEN, "qnetrnn_ckpt_golden.npz")))
        net = pp.QNetRNN()
        net.load_state_dict({k: torch.as_tensor(v) for k, v in gu.golden_sd(g, "rnn_agent_4_B").items()})
    else:
        torch.manual_seed(7)
        net = pp.QNetRNN()
    net.reset_noise()
    return net


def _filled_ring(n, T, seed):
    g = torch.Generator().manual_seed(seed)
    ring = pp.ReplayRing(n * T, lockstep_envs=n)
    ring.obs.copy_(torch.rand(n * T, 7, generator=g) * 2 - 1)
    ring.next_obs.copy_(torch.rand(n * T, 7, generator=g) * 2 - 1)
    ring.act.copy_(torch.randint(0, 3, (n * T,), generator=g).to(torch.uint8))
    ring.rew.copy_(torch.randint(-1, 2, (n * T,), generator=g).float())
    ring.done.copy_((torch.rand(n * T, generator=g) < 0.3).to(torch.uint8))
    return ring


def _reference(net, target, ring, rows, gamma, noisy):
    """Loss and gradients of train_step_rnn through torch autograd on the CPU."""
    ref, tgt = QNetRNNPort(), QNetRNNPort()
    ref.load_state_dict({k: v.detach().cpu() for k, v in net.state_dict().items()})
    tgt.load_state_dict({k: v.detach().cpu() for k, v in target.state_dict().items()})
    ref.train(noisy); tgt.eval()
    r = rows.cpu()
    take = lambda t: t.detach().cpu()[r]
    loss = drqn_loss(ref, tgt, take(ring.obs), take(ring.act).long(), take(ring.rew), take(ring.next_obs), take(ring.done) != 0, gamma)
    ref.zero_grad()
    loss.backward()
    return float(loss.detach()), {k: p.grad.clone() for k, p in ref.named_parameters()}, ref


@pytest.mark.parametrize("which,batch,trace", [("seed", 64, 8), ("ckpt", 64, 8), ("seed", 16, 5), ("ckpt", 128, 12)])
def test_drqn_grads_equal_torch_autograd(which, batch, trace):
    net = _net(which)
    n, T = 64, 32
    ring = _filled_ring(n, T, seed=batch + trace)
    g = torch.Generator().manual_seed(3)
    t_end = torch.randint(trace - 1, T, (batch,), generator=g)
    env = torch.randint(0, n, (batch,), generator=g)
    rows = ((t_end.unsqueeze(1) + torch.arange(-(trace - 1), 1).unsqueeze(0)) * n + env.unsqueeze(1)).cuda()
    tr = DRQNTrainer(copy.deepcopy(net), gamma=0.97, batch_size=batch, use_graph=False)
    with torch.no_grad():                                # a target that differs from the online net
        for p in tr.target.parameters():
            p.mul_(0.9)
    assert tr.fused
    loss = float(tr.grads_on_rows(ring, rows))
    want_loss, want, _ = _reference(tr.model, tr.target, ring, rows, 0.97, noisy=True)
    assert loss == pytest.approx(want_loss, rel=2e-5, abs=1e-7)
    worst = 0.0
    for name, p in tr.model.named_parameters():
        got, w = p.grad.detach().cpu(), want[name]
        scale = max(float(w.abs().max()), 1e-8)
        err = float((got - w).abs().max()) / scale
        worst = max(worst, err)
        assert err < 1e-4, (name, err, scale)
    print(f"{which} batch {batch} trace {trace}: loss {loss:.6f}, worst relative gradient error {worst:.2e}")


def test_drqn_clip_and_adam_equal_torch():
    """clip_grad_norm_(1.0) + Adam on the flat gradient buffer == torch's, over several steps (large and small norms)."""
    net = _net("seed")
    tr = DRQNTrainer(copy.deepcopy(net), lr=1e-3, batch_size=64, use_graph=False)
    ref = copy.deepcopy(tr.model).cpu()
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    g = torch.Generator().manual_seed(1)
    for step in range(4):
        scale = [30.0, 1e-3, 3.0, 0.2][step]
        for p, q in zip(tr.model.parameters(), ref.parameters()):
            gr = torch.randn(p.shape, generator=g) * scale / 400
            p.grad.copy_(gr)
            q.grad = gr.clone()
        norm = torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm=1.0)
        opt.step()
        tr.clip_and_step()
        assert float(tr._norm_out[0]) == pytest.approx(float(norm), rel=1e-5)
        assert float(tr._norm_out[1]) == pytest.approx(min(1.0, 1.0 / (float(norm) + 1e-6)), rel=1e-5)
        for (name, p), q in zip(tr.model.named_parameters(), ref.parameters()):
            assert torch.allclose(p.detach().cpu(), q.detach(), rtol=2e-5, atol=2e-7), (step, name)
    assert all(float(tr.opt.state[p]["step"]) == 4.0 for p in tr.params)


def test_drqn_fused_update_tracks_the_autograd_update_in_a_generation():
    """Same windows, same noise: the product trainer and the autograd / cuDNN formulation (oracle TorchDRQNTrainer) stay together over the
    eager updates; then the fused update runs from its captured CUDA graph."""
    torch.manual_seed(5)
    net = pp.QNetRNN()
    n, T = 256, 32
    ring = _filled_ring(n, T, seed=9)
    ring.done.copy_((torch.rand(n * T, generator=torch.Generator().manual_seed(2)) < 0.06).to(torch.uint8))
    ring.steps_written = T
    ring.head.fill_(n * T)
    sampler = SequenceSampler(ring, trace_length=8)
    assert sampler.refresh() > 64
    a = DRQNTrainer(copy.deepcopy(net), lr=1e-3, batch_size=64, use_graph=True)
    b = TorchDRQNTrainer(copy.deepcopy(net), lr=1e-3, batch_size=64, use_graph=False, device="cuda")
    b.model.load_state_dict(a.model.state_dict())        # the same epsilon buffers
    torch.backends.cudnn.allow_tf32 = False
    for step in range(3):                                # the eager updates, on the same windows for both
        rows = sampler.sample_rows(64, generator=torch.Generator(device="cuda").manual_seed(100 + step)).clone()
        sampler.sample_rows = lambda *args, **kw: rows   # instance override: both trainers draw through it
        la = a.update(sampler)
        lb = b.update(sampler)
        del sampler.sample_rows
        assert float(la) == pytest.approx(float(lb), rel=1e-3, abs=1e-6), step
    for (name, p), q in zip(a.model.named_parameters(), b.model.parameters()):
        assert torch.allclose(p, q, rtol=1e-2, atol=2e-4), name
    before = {k: v.clone() for k, v in a.model.state_dict().items()}
    losses = [float(a.update(sampler)) for _ in range(5)]         # captured at the 4th update, replayed afterwards
    assert a._graph is not None and a.train_steps == 8 and all(np.isfinite(l) and l > 0 for l in losses)
    assert len(set(losses)) == 5                                    # every replay draws fresh windows
    assert not torch.equal(before["lstm.weight_hh_l0"], a.model.state_dict()["lstm.weight_hh_l0"])


@pytest.mark.parametrize("noisy", [False, True])
def test_device_pack_of_the_tensor_core_weight_image_equals_the_host_pack(noisy):
    """pp_pack_qnetrnn_tc (one launch) == policy.pack_qnetrnn_tc (the torch formulation), byte for byte, for a random-init
    net and the reference's trained checkpoint."""
    import ctypes as C
    from pingpong_selfplay_ai_b200 import _lib
    from pingpong_selfplay_ai_b200.policy import pack_qnetrnn_tc
    from pingpong_selfplay_ai_b200.selfplay import _ptr, _stream_ptr
    for which in ("seed", "ckpt"):
        net = _net(which).cuda()
        want = pack_qnetrnn_tc(net, noisy=noisy).cpu().numpy()
        st = DRQNTrainer._net_struct(net)
        img = torch.zeros(_lib.RNNTC_BLOB_BYTES, dtype=torch.uint8, device="cuda")
        _lib.check(_lib.load().pp_pack_qnetrnn_tc(C.byref(st), int(noisy), _ptr(img), _stream_ptr(torch.device("cuda", 0))))
        got = img.cpu().numpy()
        bad = np.flatnonzero(got != want)
        assert bad.size == 0, (which, bad[:8], bad.size)


@pytest.mark.parametrize("steps", [25, 40, 41, 97, 200])
def test_sequence_sampler_kernels_equal_the_reference_episode_then_window_distribution(steps):
    """pp_seq_window_weights / pp_per_sample / pp_seq_expand_rows on the device == SequenceReplayBuffer's distribution
    (scripts/train_rnn_iterative.py:115-141, restated in tests/test_train_rnn_logic.py): the weight of every window end,
    the stored-episode count, and 61 440 drawn windows (contiguous, inside one stored episode, frequencies = weights)."""
    import test_train_rnn_logic as host
    n, T, L = 50, 40, host.L
    ring_cpu, done = host._lockstep_ring(n, T, steps, seed=steps, p_done=0.07)
    ring = pp.ReplayRing(n * T, lockstep_envs=n)
    for name in ("obs", "next_obs", "act", "rew", "done"):
        getattr(ring, name).copy_(getattr(ring_cpu, name))
    ring.steps_written = steps
    want, episodes = host._reference_buffer(done, n, T, steps)
    s = SequenceSampler(ring, trace_length=L)
    assert s.refresh() == episodes == len(s)
    w = gu.np_of(s._k_w).reshape(T, n)
    lo = max(0, steps - T)
    got = {}
    for r, i in zip(*np.nonzero(w)):
        t = next(t for t in range(lo, steps) if t % T == r)                     # the absolute step held by ring row r
        got[(t, int(i))] = float(w[r, i])
    assert got.keys() == want.keys() and all(abs(got[k] - want[k]) < 1e-6 for k in want)
    if not episodes:
        return
    draws = 15 * 4096
    rows = np.concatenate([gu.np_of(s.sample_rows(4096, seed=3)).copy() for _ in range(15)])    # static buffer: copy each draw
    assert rows.shape == (draws, L)
    t_abs = gu.np_of(ring.obs)[rows, 0].astype(np.int64)                         # obs[.., 0] = absolute step, obs[.., 1] = env
    env = gu.np_of(ring.obs)[rows, 1].astype(np.int64)
    assert np.all(np.diff(t_abs, axis=1) == 1) and np.all(env == env[:, :1])
    assert not done[t_abs[:, :-1], env[:, :-1]].any()                            # no episode end inside a window
    ends = list(zip(t_abs[:, -1].tolist(), env[:, -1].tolist()))
    assert set(ends) <= set(want)
    freq = {k: 0 for k in want}
    for e in ends:
        freq[e] += 1
    total = sum(want.values())
    for k, wv in want.items():
        p = wv / total
        assert abs(freq[k] / draws - p) < 5 * np.sqrt(p * (1 - p) / draws) + 1e-4, (k, freq[k] / draws, p)
    again = gu.np_of(s.sample_rows(64, seed=3))
    assert not np.array_equal(again, rows[:64])                                  # the draw counter advances
