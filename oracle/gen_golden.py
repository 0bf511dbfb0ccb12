"""Generate tests/golden/* from the UNMODIFIED reference (container-only; needs /root/reference).

TEST INFRASTRUCTURE.  Run:  python -m oracle.gen_golden
The reference repository has no tests/fixtures for the hot path, so these
vectors — outputs of the reference itself, executed here through oracle/ref_shim.py — are what
pins the oracle and, through it, the CUDA kernels.  Files written:

  env_hashes.json        sha256 of 10 000-step runs (config.yaml, config_rnn.yaml env blocks; the
                         protocol of SURVEY.md section 8c) + collision known answers (hex floats)
  env_traj_config.npz    the config.yaml run step by step: actions, post-step state, obs, rewards,
                         done, and the serve (vx,vy,spin) of every reset in order
  env_traj_rnncfg.npz    same for the config_rnn.yaml env block (first 4 000 steps)
  env_random_steps.npz   8 192 single steps from random (incl. out-of-range / quirk) states
  qnet_golden.npz        QNet weights (seed 0, seed 1, checkpoints/model5-1_fault.pth) + Q on real obs
  qnetrnn_golden.npz     QNetRNN weights (seed 0) + 12-step carried-(h,c) Q sequences
  qnetrnn_ckpt_golden.npz  same for checkpoints_rnn/rnn_agent_4.pth
"""
from __future__ import annotations

import hashlib
import json
import os
import random
import struct

import numpy as np
import torch

from . import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _state(env):
    return (float(env.ball_x), float(env.ball_y), float(env.ball_vx), float(env.ball_vy), float(env.spin),
            float(env.top_paddle_x), float(env.bottom_paddle_x))


def run_trajectory(PongEnv2P, env_cfg, steps, seed=12345, action_seed=777):
    random.seed(seed)
    env = PongEnv2P(**env_cfg)
    env.reset()
    serves = [(float(env.ball_vx), float(env.ball_vy), float(env.spin))]
    arng = random.Random(action_seed)
    h_state, h_obs = hashlib.sha256(), hashlib.sha256()
    acts = np.zeros((steps, 2), np.uint8)
    st = np.zeros((steps, 7), np.float64)
    si = np.zeros((steps, 3), np.int32)
    obs = np.zeros((steps, 2, 7), np.float32)
    rew = np.zeros((steps, 2), np.float32)
    done_arr = np.zeros(steps, np.uint8)
    episodes = wins_a = wins_b = 0
    for t in range(steps):
        a, b = arng.randint(0, 2), arng.randint(0, 2)
        (oa, ob), (ra, rb), done, _ = env.step(a, b)
        acts[t] = (a, b)
        st[t] = _state(env)
        si[t] = (env.scoreA, env.scoreB, env.bounce_count)
        obs[t, 0], obs[t, 1] = oa, ob
        rew[t] = (ra, rb)
        done_arr[t] = done
        h_state.update(struct.pack("<7d3i?", *st[t].tolist(), int(env.scoreA), int(env.scoreB),
                                   int(env.bounce_count), bool(done)))
        h_obs.update(oa.tobytes() + ob.tobytes() + struct.pack("<2d", ra, rb))
        if done:
            episodes += 1
            wins_a += env.scoreA > env.scoreB
            wins_b += env.scoreB > env.scoreA
            env.reset()
            serves.append((float(env.ball_vx), float(env.ball_vy), float(env.spin)))
    hits = int((np.diff(np.concatenate([[0], si[:, 2]])) == 1).sum())
    summary = dict(steps=steps, episodes=episodes, wins_a=int(wins_a), wins_b=int(wins_b), paddle_hits=hits,
                   state_sha256=h_state.hexdigest(), obs_sha256=h_obs.hexdigest(),
                   first_serve=[float(v).hex() for v in serves[0]])
    arrays = dict(actions=acts, state=st, ints=si, obs=obs, rew=rew, done=done_arr,
                  serves=np.asarray(serves, np.float64))
    return summary, arrays


def random_single_steps(PongEnv2P, env_cfg, n, seed):
    """Single steps from injected states that cover the quirk list of SURVEY.md section 8a."""
    rs = np.random.RandomState(seed)
    random.seed(seed)
    env = PongEnv2P(**env_cfg)
    hw = env_cfg["paddle_width"] / 2
    pre = np.zeros((n, 7), np.float64); pre_i = np.zeros((n, 3), np.int32); acts = np.zeros((n, 2), np.uint8)
    post = np.zeros((n, 7), np.float64); post_i = np.zeros((n, 3), np.int32)
    obs = np.zeros((n, 2, 7), np.float32); rew = np.zeros((n, 2), np.float32); done = np.zeros(n, np.uint8)
    for i in range(n):
        kind = i % 8
        x, y = rs.uniform(0, 1), rs.uniform(0, 1)
        vx, vy = rs.uniform(-0.08, 0.08), rs.uniform(-0.08, 0.08)
        spin = rs.uniform(-40, 40)
        top, bot = rs.uniform(0, 1), rs.uniform(0, 1)
        a, b = rs.randint(0, 4), rs.randint(0, 4)       # 3 = "any other value: no move"
        sa, sb, bc = rs.randint(0, 3), rs.randint(0, 3), rs.randint(0, 12)
        if kind == 1:      # about to cross the top line, paddle roughly under the ball
            y, vy = rs.uniform(0, 0.03), -rs.uniform(0.03, 0.2)
            top = float(np.clip(x + vx + rs.uniform(-1.5, 1.5) * hw, 0, 1))
        elif kind == 2:    # about to cross the bottom line
            y, vy = 1 - rs.uniform(0, 0.03), rs.uniform(0.03, 0.2)
            bot = float(np.clip(x + vx + rs.uniform(-1.5, 1.5) * hw, 0, 1))
        elif kind == 3:    # already out of bounds, flying outward (re-hit quirk: no direction check)
            y, vy = -rs.uniform(0.0, 0.5), -rs.uniform(0.0, 0.1)
            top = float(np.clip(x + rs.uniform(-1.2, 1.2) * hw, 0, 1))
        elif kind == 4:    # paddle pinned at a wall while its action still gives u != 0
            top, a = (0.0, 0) if rs.rand() < 0.5 else (1.0, 2)
            x, vx, spin = top, 0.0, 0.0
            y, vy = 0.01, -0.05
        elif kind == 5:    # very fast ball: single reflection leaves x outside [0,1]
            vx = rs.uniform(-2.5, 2.5)
        elif kind == 6:    # exact paddle edge / vrel == +-0.0 cases
            bot = 0.5; b = 1; x = 0.5 + (hw if rs.rand() < 0.5 else -hw); vx = 0.0; spin = 0.0
            y, vy = 0.99, 0.05
            if rs.rand() < 0.5:
                x = 0.5; vx = 0.0 if rs.rand() < 0.5 else -0.0
        elif kind == 7:    # match point
            sa = sb = env_cfg["max_score"] - 1
            y, vy = (0.01, -0.3) if rs.rand() < 0.5 else (0.99, 0.3)
        env.ball_x, env.ball_y, env.ball_vx, env.ball_vy, env.spin = x, y, vx, vy, spin
        env.top_paddle_x, env.bottom_paddle_x = top, bot
        env.scoreA, env.scoreB, env.bounce_count = int(sa), int(sb), int(bc)
        pre[i] = (x, y, vx, vy, spin, top, bot); pre_i[i] = (sa, sb, bc); acts[i] = (a, b)
        (oa, ob), (ra, rb), d, _ = env.step(int(a), int(b))
        post[i] = _state(env); post_i[i] = (env.scoreA, env.scoreB, env.bounce_count)
        obs[i, 0], obs[i, 1] = oa, ob; rew[i] = (ra, rb); done[i] = d
    return dict(pre=pre, pre_i=pre_i, actions=acts, post=post, post_i=post_i, obs=obs, rew=rew, done=done)


def _sd_np(sd, prefix):
    return {f"{prefix}/{k}": v.detach().cpu().numpy() for k, v in sd.items()}


def gen_extra_cfgs():
    """tests/golden/env_extra_cfgs.npz: 3 000-step reference trajectories for pong_port.EXTRA_ENV_CONFIGS."""
    from .pong_port import EXTRA_ENV_CONFIGS
    PongEnv2P, _, _, _ = ref_shim.load_reference()
    out = {}
    for ci, kw in enumerate(EXTRA_ENV_CONFIGS):
        summary, arr = run_trajectory(PongEnv2P, kw, 3000, seed=500 + ci, action_seed=900 + ci)
        for k, v in arr.items():
            out[f"c{ci}/{k}"] = v
        print(ci, {k: summary[k] for k in ("episodes", "wins_a", "wins_b", "paddle_hits")})
    np.savez_compressed(os.path.join(OUT, "env_extra_cfgs.npz"), **out)


def gen_checkpoint_golden():
    """tests/golden/ckpt_golden.npz: what the reference's own loader (tests/test_round_robin.py:116-185) makes of the
    three on-disk formats — Q-values of the loaded nets on seeded observations.  The legacy (`fc.*`) case is a synthetic
    state_dict stored in the fixture itself, so the test needs no reference file; the three real files are addressed by
    their path under the reference root and checked only where that tree exists (this container)."""
    import tempfile

    import torch
    rr = ref_shim.load_reference_round_robin()
    rs = np.random.RandomState(2468)
    obs = np.concatenate([rs.uniform(-1, 1, size=(48, 7)), rs.uniform(-0.1, 1.1, size=(48, 7))]).astype(np.float32)
    out = {"obs": obs}
    g = torch.Generator().manual_seed(97)
    legacy = {"fc.0.weight": torch.randn(64, 7, generator=g) * 0.6, "fc.0.bias": torch.randn(64, generator=g) * 0.2,
              "fc.2.weight": torch.randn(64, 64, generator=g) * 0.25, "fc.2.bias": torch.randn(64, generator=g) * 0.2,
              "fc.4.weight": torch.randn(3, 64, generator=g) * 0.3, "fc.4.bias": torch.randn(3, generator=g) * 0.2}
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "legacy.pth")
        torch.save({"model": legacy, "epsilon": 0.1, "episode": 7}, path)
        net = rr.load_model_universal({"name": "legacy", "path": path, "type": "QNet"}, {}, torch.device("cpu"))
    with torch.no_grad():
        out["legacy/q"] = net(torch.from_numpy(obs)).numpy()
    for k, v in legacy.items():
        out["legacy/sd/" + k] = v.numpy()
    real = {"legacy_file": ("QNet", "checkpoints/model4-12.pth"), "dueling_file": ("QNet", "checkpoints/model5-3_fault.pth"),
            "rnn_file": ("QNetRNN", "checkpoints_rnn/rnn_pong_soul_2.pth")}
    for tag, (typ, rel) in real.items():
        net = rr.load_model_universal({"name": tag, "path": os.path.join(ref_shim.REFERENCE_ROOT, rel), "type": typ}, {},
                                      torch.device("cpu"))
        with torch.no_grad():
            if typ == "QNet":
                q = net(torch.from_numpy(obs))
            else:                                           # two steps, so (h, c) of the first one matter
                h = net.init_hidden(obs.shape[0], torch.device("cpu"))
                _, h = net(torch.from_numpy(obs[::-1].copy()).unsqueeze(1), h)
                q, _ = net(torch.from_numpy(obs).unsqueeze(1), h)
        out[tag + "/q"] = q.numpy()
        out[tag + "/path"] = np.array(rel)
        out[tag + "/type"] = np.array(typ)
    np.savez_compressed(os.path.join(OUT, "ckpt_golden.npz"), **out)
    print({k: (v.shape if v.ndim else str(v)) for k, v in out.items() if not k.startswith("legacy/sd")})


def synthetic_arena_database():
    """A small arena_database.json-shaped history: 5 models, uneven pair counts, swapped sides, a draw."""
    import random
    rng = random.Random(4242)
    models = [{"id": f"m{k}", "type": "QNet", "path": f"checkpoints/m{k}.pth", "description": f"model {k}"} for k in range(4)]
    models.append({"id": "BallFollowerBot", "type": "HardcodedBallFollower", "path": "N/A"})
    ids = [m["id"] for m in models]
    hist = []
    for a in range(len(ids)):
        for b in range(a + 1, len(ids)):
            for _ in range(rng.randint(0, 7)):
                p1, p2 = (ids[a], ids[b]) if rng.random() < 0.6 else (ids[b], ids[a])
                sa, sb = (3, rng.randint(0, 2)) if rng.random() < 0.55 else (rng.randint(0, 2), 3)
                hist.append({"p1": p1, "p2": p2, "winner": p1 if sa > sb else p2, "p1_score": sa, "p2_score": sb,
                             "timestamp": "2025-01-01T00:00:00Z"})
    hist.append({"p1": "m0", "p2": "m1", "winner": "draw", "p1_score": 2, "p2_score": 2, "timestamp": "2025-01-01T00:00:00Z"})
    return {"models": models, "match_history": hist}


def gen_arena_golden():
    """tests/golden/arena_golden.json: the reference's own create_match_plan / generate_summary_report / register_models
    (tests/arena.py:147-155,222-245,323-352) on the synthetic database above."""
    import copy
    arena = ref_shim.load_reference_round_robin("arena.py")
    db = synthetic_arena_database()
    plan = arena.create_match_plan(copy.deepcopy(db), 5)
    summary = arena.generate_summary_report(copy.deepcopy(db)).reset_index().to_dict(orient="records")
    db2 = copy.deepcopy(db)
    added = arena.register_models(db2, [{"id": "m1", "type": "QNet", "path": "x"}, {"id": "new", "type": "QNetRNN", "path": "y"}])
    with open(os.path.join(OUT, "arena_golden.json"), "w") as f:
        json.dump({"database": db, "plan_5": plan, "summary": summary, "register_added": bool(added),
                   "register_ids": [m["id"] for m in db2["models"]]}, f, indent=1)
    print(len(db["match_history"]), "records;", len(plan), "pairings to play;", summary[0])


def main():
    import sys
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "arena":
        gen_arena_golden()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "ckpt":
        gen_checkpoint_golden()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "extra":
        gen_extra_cfgs()
        return
    PongEnv2P, collide, QNet, QNetRNN = ref_shim.load_reference()
    cfg = ref_shim.load_reference_config("config.yaml")["env"]
    cfg_rnn = ref_shim.load_reference_config("config_rnn.yaml")["env"]

    # ---- environment
    s1, a1 = run_trajectory(PongEnv2P, cfg, 10000)
    s2, a2 = run_trajectory(PongEnv2P, cfg_rnn, 10000)
    np.savez_compressed(os.path.join(OUT, "env_traj_config.npz"), **a1)
    np.savez_compressed(os.path.join(OUT, "env_traj_rnncfg.npz"),
                        **{k: (v if k == "serves" else v[:4000]) for k, v in a2.items()})
    kat_in = [(-0.04, 0.02, 0.03, 3.0), (-0.04, 0.02, -0.03, -5.0), (-0.05, 0.0, 0.0, 0.0),
              (-0.03, 0.05, 0.0, 40.0), (-0.001, 0.3, 0.03, -50.0), (0.07, -0.02, -0.03, 12.5),
              (-0.05, 0.0, 0.0, -0.0), (-0.05, -0.0, 0.0, 0.0)]
    kats = []
    for vn, vt, u, om in kat_in:
        out = collide(vn, vt, u, om, cfg["restitution"], cfg["friction"], cfg["ball_mass"], cfg["world_ball_radius"])
        kats.append(dict(inp=[float(v).hex() for v in (vn, vt, u, om)], out=[float(v).hex() for v in out]))
    with open(os.path.join(OUT, "env_hashes.json"), "w") as f:
        json.dump(dict(protocol="random.seed(12345); env=PongEnv2P(**cfg); env.reset(); actions from "
                                "random.Random(777).randint(0,2) x2 per step; hash post-step state then reset on done",
                       config_yaml=s1, config_rnn_yaml=s2, env_config_yaml=cfg, env_config_rnn_yaml=cfg_rnn,
                       collision_kat=kats), f, indent=1)
    np.savez_compressed(os.path.join(OUT, "env_random_steps.npz"),
                        **{f"cfg/{k}": v for k, v in random_single_steps(PongEnv2P, cfg, 8192, 2024).items()},
                        **{f"rnn/{k}": v for k, v in random_single_steps(PongEnv2P, cfg_rnn, 2048, 2025).items()})

    # ---- QNet: observations taken from the real trajectory (both players' views)
    obs = np.concatenate([a1["obs"][:1024, 0], a1["obs"][:1024, 1]]).astype(np.float32)
    out = dict(obs=obs)
    nets = {}
    for seed in (0, 1):
        torch.manual_seed(seed)
        nets[f"seed{seed}"] = QNet(input_dim=7, output_dim=3)
    ck = torch.load(os.path.join(ref_shim.REFERENCE_ROOT, "checkpoints", "model5-1_fault.pth"),
                    map_location="cpu", weights_only=True)
    net = QNet(7, 3); net.load_state_dict(ck["modelB"]); nets["ckpt_model5_1_fault_B"] = net
    with torch.no_grad():
        for name, net in nets.items():
            out.update(_sd_np(net.state_dict(), name))
            net.eval(); out[f"{name}/q_eval"] = net(torch.from_numpy(obs)).numpy()
            net.train(); out[f"{name}/q_train"] = net(torch.from_numpy(obs)).numpy()
    np.savez_compressed(os.path.join(OUT, "qnet_golden.npz"), **out)

    # ---- QNetRNN: 12 carried steps for 48 envs (each env = a slice of the trajectory)
    T, B = 12, 48
    seq = np.stack([a1["obs"][i * 40:i * 40 + T, i % 2] for i in range(B)]).astype(np.float32)   # [B,T,7]

    def rnn_pack(net, name):
        o = dict(_sd_np(net.state_dict(), name))
        with torch.no_grad():
            for mode in ("eval", "train"):
                net.train(mode == "train")
                hc = net.init_hidden(B, "cpu")
                qs = []
                for t in range(T):
                    q, hc = net(torch.from_numpy(seq[:, t:t + 1]), hc)
                    qs.append(q.numpy())
                o[f"{name}/q_{mode}"] = np.stack(qs)                  # [T,B,3]
                o[f"{name}/h_{mode}"] = hc[0][0].numpy(); o[f"{name}/c_{mode}"] = hc[1][0].numpy()
        return o

    torch.manual_seed(0)
    np.savez_compressed(os.path.join(OUT, "qnetrnn_golden.npz"), seq=seq, **rnn_pack(QNetRNN(), "seed0"))
    ck = torch.load(os.path.join(ref_shim.REFERENCE_ROOT, "checkpoints_rnn", "rnn_agent_4.pth"),
                    map_location="cpu", weights_only=True)
    sd = ck.get("modelB_state", ck.get("modelA_state"))
    net = QNetRNN(); net.load_state_dict(sd)
    np.savez_compressed(os.path.join(OUT, "qnetrnn_ckpt_golden.npz"), seq=seq, **rnn_pack(net, "rnn_agent_4_B"))
    print(json.dumps(dict(config_yaml=s1, config_rnn_yaml=s2), indent=1))
    for fn in sorted(os.listdir(OUT)):
        print(f"{os.path.getsize(os.path.join(OUT, fn)):>9d}  {fn}")


if __name__ == "__main__":
    main()
