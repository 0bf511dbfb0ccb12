"""Host-side logic on the CPU: parameter precomputation, weight packing, network mirrors, slab partitioning and the
world_size-2 gloo path of the collectives."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import pingpong_selfplay_ai_b200 as pp
from pingpong_selfplay_ai_b200 import _lib, dist as ppdist
from oracle import pong_oracle as po
import pp_testutil as gu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("key", ["env_config_yaml", "env_config_rnn_yaml"])
def test_params_equal_oracle_params_bitwise(key):
    cfg = gu.hashes()[key]
    p, o = pp.make_params(pp.resolve_env_config(cfg)), po.make_params(cfg)
    for name, _ in po.OracleParams._fields_:
        if name != "pad_":
            assert getattr(p, name) == getattr(o, name), name
    assert float(p.inertia).hex() == "0x1.797cc39ffd60fp-12" and float(p.two_m_over_7).hex() == "0x1.2492492492492p-2"
    assert (p.speed_lo, p.speed_hi, p.spin_lo, p.spin_hi) == (0.03, 0.05, -5.0, 5.0)
    assert list(p.angle_lo) == [-60.0, 30.0] and list(p.angle_hi) == [-30.0, 60.0]


def test_env_kwargs_follow_reference_constructor():
    cfg = pp.resolve_env_config({})
    assert cfg["paddle_speed"] == 0.02 and cfg["speed_scale_every"] == 3 and cfg["ball_angle_intervals"] == [[-60, -30], [30, 60]]
    with pytest.raises(TypeError):
        pp.resolve_env_config({"paddle_height": 1})


def _blob_forward(blob, obs):
    """Numpy forward straight from the packed k-major blob (checks the packing, not the kernel)."""
    o = _lib.QNET_OFF
    w1t, b1 = blob[o["W1T"]:o["B1"]].reshape(7, 64), blob[o["B1"]:o["W2T"]]
    w2t, b2 = blob[o["W2T"]:o["B2"]].reshape(64, 64), blob[o["B2"]:o["WHT"]]
    wht, bh = blob[o["WHT"]:o["BH"]].reshape(64, 4), blob[o["BH"]:]
    h = np.maximum(obs @ w1t + b1, 0)
    h = np.maximum(h @ w2t + b2, 0)
    y = h @ wht + bh
    return y[:, :1] + (y[:, 1:] - y[:, 1:].mean(1, keepdims=True))


@pytest.mark.parametrize("noisy", [False, True])
def test_pack_qnet_layout_and_mirror_outputs(noisy):
    g = dict(np.load(os.path.join(gu.GOLDEN, "qnet_golden.npz")))
    for name in ("seed0", "ckpt_model5_1_fault_B"):
        sd = gu.golden_sd(g, name)
        blob = pp.pack_qnet(sd, noisy=noisy).numpy()
        ref = g[f"{name}/q_{'train' if noisy else 'eval'}"]
        assert np.abs(_blob_forward(blob, g["obs"]) - ref).max() < 1e-5
        net = pp.QNet()
        net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)     # reference keys, unchanged
        net.train(noisy)
        with torch.no_grad():
            assert np.abs(net(torch.from_numpy(g["obs"])).numpy() - ref).max() < 1e-6
        assert torch.equal(pp.pack_qnet(net, noisy=noisy), torch.from_numpy(blob))


def test_pack_qnetrnn_layout_and_mirror_outputs():
    g = dict(np.load(os.path.join(gu.GOLDEN, "qnetrnn_golden.npz")))
    sd = gu.golden_sd(g, "seed0")
    blob = pp.pack_qnetrnn(sd).numpy()
    o = _lib.RNN_OFF
    wg = blob[o["WGT"]:o["BG"]].reshape(256, 128, 4)                         # [k][unit][gate]
    assert np.array_equal(wg[:128, 5, 2], sd["lstm.weight_ih_l0"][2 * 128 + 5])      # gate g, unit 5, over features
    assert np.array_equal(wg[128:, 9, 3], sd["lstm.weight_hh_l0"][3 * 128 + 9])
    bg = blob[o["BG"]:o["WST"]].reshape(128, 4)
    assert np.array_equal(bg[7, 1], (sd["lstm.bias_ih_l0"] + sd["lstm.bias_hh_l0"])[128 + 7])
    assert np.array_equal(blob[o["WST"]:o["BS"]].reshape(128, 128), sd["fc_shared_head.0.weight_mu"].T)
    net = pp.QNetRNN()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    net.eval()
    seq = torch.from_numpy(g["seq"])
    hc = net.init_hidden(seq.shape[0], "cpu")
    with torch.no_grad():
        for t in range(seq.shape[1]):
            q, hc = net(seq[:, t:t + 1], hc)
            assert np.abs(q.numpy() - g["seed0/q_eval"][t]).max() < 1e-6
    with pytest.raises(ValueError):
        pp.pack_qnetrnn(pp.QNetRNN(feature_dim=64))


def test_noisy_linear_statistics_and_reset():
    torch.manual_seed(0)
    lin = pp.NoisyLinear(64, 3)
    assert float(lin.weight_sigma[0, 0].detach()) == pytest.approx(0.017) and lin.weight_mu.abs().max() <= 1 / 8
    e0 = lin.weight_epsilon.clone()
    lin.reset_noise()
    assert not torch.equal(e0, lin.weight_epsilon)
    assert torch.allclose(lin.weight_epsilon, torch.outer(lin.bias_epsilon, lin.weight_epsilon[0] / lin.bias_epsilon[0]), atol=1e-5)


def test_eps_threshold_and_slab_bounds():
    from pingpong_selfplay_ai_b200.policy import eps_threshold
    assert eps_threshold(0.0) == 0 and eps_threshold(1.0) == 1 << 32 and eps_threshold(0.5) == 1 << 31
    assert eps_threshold(0.3) == po.eps_threshold(0.3)
    for n, w in [(10, 3), (262144, 8), (7, 8), (1 << 20, 8)]:
        b = [ppdist.slab_bounds(n, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        ppdist.slab_bounds(8, 2, 2)


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["PP_ROOT"])
from pingpong_selfplay_ai_b200 import dist as ppd, QNet
rank, world, local = ppd.init_from_env("gloo")
assert world == 2 and ppd.is_parallel()
lo, hi = ppd.slab_bounds(1001, world, rank)
counters = torch.tensor([hi - lo, rank + 1, 0, 0, 0, 0, 0, 7], dtype=torch.int64)
tot = ppd.allreduce_counters(counters)
assert tot.tolist() == [1001, 3, 0, 0, 0, 0, 0, 14] and counters[0] == hi - lo
torch.manual_seed(0)
net = QNet()
heads = list(net.fc_V.parameters()) + list(net.fc_A.parameters())
assert sum(p.numel() for p in heads) == 520
x = torch.full((4, 7), float(rank + 1))
net(x).sum().backward()
local_g = [p.grad.clone() for p in heads]
ppd.allreduce_mean_grads(heads)
gathered = [torch.zeros(520) for _ in range(2)]
dist.all_gather(gathered, torch.cat([g.reshape(-1) for g in local_g]))
want = (gathered[0] + gathered[1]) / 2
assert torch.allclose(torch.cat([p.grad.reshape(-1) for p in heads]), want, atol=1e-7)
for p, g in zip(heads, local_g):                      # the same through one flat buffer that the .grad tensors view
    p.grad = g.clone()
flat = ppd.flatten_grads_(heads)
assert flat.numel() == 520 and all(p.grad.data_ptr() >= flat.data_ptr() for p in heads)
ppd.allreduce_mean_flat_(flat)
assert torch.allclose(torch.cat([p.grad.reshape(-1) for p in heads]), want, atol=1e-7) and torch.allclose(flat, want, atol=1e-7)
assert ppd.max_over_ranks(float(rank)) == 1.0
t = torch.full((3,), float(rank)); ppd.broadcast_(t, 0); assert t.tolist() == [0.0, 0.0, 0.0]
dist.destroy_process_group()
sys.stdout.write(f"ok {rank}\n"); sys.stdout.flush()
"""


def test_gloo_world_size_2_counters_and_grad_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, PP_ROOT=ROOT, OMP_NUM_THREADS="1")
    import socket
    for attempt in range(3):                               # the rendezvous (port grab, store start-up) can lose a race
        with socket.socket() as sock:                      # a free rendezvous port (parallel test runs must not collide)
            sock.bind(("127.0.0.1", 0))
            port = sock.getsockname()[1]
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                            "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                           capture_output=True, text=True, env=env, timeout=240)
        if r.returncode == 0 or "AssertionError" in r.stderr:          # a worker's own assertion is a real failure
            break
    assert r.returncode == 0, r.stdout + r.stderr
    # the two workers share the pipe: their lines may interleave character-wise
    assert r.stdout.count("ok") == 2 and sorted(c for c in r.stdout if c.isdigit()) == ["0", "1"], r.stdout


_ARENA_WORKER = r"""
import copy, os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["PP_ROOT"])
from pingpong_selfplay_ai_b200 import arena, checkpoint as ck, dist as ppd
rank, world, _ = ppd.init_from_env("gloo")

class StubMatch:                                   # stands in for a pairing on the device: scores are a function of the seed
    def __init__(self, env_cfg, a, b, episodes, seed, precision, mode, device, stream, first_game=0):
        self.a, self.b, self.n, self.seed, self.stream = a, b, episodes, seed + 1000 * first_game, stream
    def launch(self, max_steps):
        pass
    def results(self):
        rs = np.random.RandomState(self.seed)
        win_a = rs.rand(self.n) < 0.5
        lose = rs.randint(0, 3, self.n)
        return np.where(win_a, 3, lose), np.where(win_a, lose, 3), rs.randint(5, 50, self.n)

models = [{"id": f"bot{k}", "type": "HardcodedBallFollower", "path": "N/A"} for k in range(5)]
agents = {m["id"]: ck.Agent(m) for m in models}
db = {"models": list(models), "match_history": []}
plan = arena.create_match_plan(db, 7)
assert len(plan) == 10
path = os.path.join(os.environ["PP_TMP"], f"db_rank{rank}.json")
res = arena.run_tournament({}, db, path, plan, agents=agents, seed=5, device="cpu", match_factory=StubMatch, concurrent=3)
one = {"models": list(models), "match_history": []}
res1 = arena.run_tournament({}, one, None, plan, agents=agents, seed=5, device="cpu", match_factory=StubMatch, shard=(0, 1))
strip = lambda h: [{k: v for k, v in r.items() if k != "timestamp"} for r in h]
assert strip(db["match_history"]) == strip(one["match_history"]) and len(db["match_history"]) == 70
assert res.keys() == res1.keys() and all(np.array_equal(res[k][0], res1[k][0]) for k in res)
assert os.path.exists(path) == (rank == 0)          # rank 0 alone writes the file
assert arena.create_match_plan(db, 7) == []
# a top-up continues each pairing's serve sequence (first_game = games already recorded) with the pairing's own seed
plan2 = arena.create_match_plan(db, 9)
seen = []
class Spy(StubMatch):
    def __init__(self, *a, first_game=0, **k):
        seen.append((a[1].id, a[2].id, a[4], first_game))
        super().__init__(*a, first_game=first_game, **k)
arena.run_tournament({}, db, None, plan2[3:], agents=agents, seed=5, device="cpu", match_factory=Spy, shard=(0, 1))
assert [s[3] for s in seen] == [7] * 7 and [s[2] for s in seen] == [5 + j for j in range(3, 10)]
dist.barrier()
dist.destroy_process_group()
sys.stdout.write(f"ok {rank}\n"); sys.stdout.flush()
"""


def test_gloo_world_size_2_tournament_pairings_shard_over_ranks(tmp_path):
    """arena.run_tournament with two ranks: pairing k runs on rank k % 2 with the single-GPU seed, both ranks end with
    the single-process history, rank 0 writes the database (the matches themselves are stubbed: no GPU here)."""
    script = tmp_path / "arena_worker.py"
    script.write_text(_ARENA_WORKER)
    env = dict(os.environ, PP_ROOT=ROOT, PP_TMP=str(tmp_path), OMP_NUM_THREADS="1")
    import socket
    for attempt in range(3):
        with socket.socket() as sock:
            sock.bind(("127.0.0.1", 0))
            port = sock.getsockname()[1]
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                            "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                           capture_output=True, text=True, env=env, timeout=240)
        if r.returncode == 0 or "AssertionError" in r.stderr:
            break
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2 and sorted(c for c in r.stdout if c.isdigit()) == ["0", "1"], r.stdout
