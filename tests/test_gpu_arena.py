"""GPU tests of the tournament front-end: every pairing of tests/arena.py's agent types
runs as one lock-step batch; game outcomes of the non-recurrent pairings are bit-exact against the oracle."""
import numpy as np
import pytest
import torch

import pp_testutil as gu

pytestmark = pytest.mark.gpu

pp = gu.pp
po = gu.po

from pingpong_selfplay_ai_b200 import arena, checkpoint as ck  # noqa: E402


@pytest.fixture(scope="module")
def H():
    return gu.hashes()


def _agents():
    torch.manual_seed(0); q0 = pp.QNet()
    torch.manual_seed(1); q1 = pp.QNet()
    torch.manual_seed(11); r0 = pp.QNetRNN()
    for net in (q0, q1, r0):
        net.eval()
    models = [{"id": "q0", "type": "QNet", "path": "mem"}, {"id": "q1", "type": "QNet", "path": "mem"},
              {"id": "rnn", "type": "QNetRNN", "path": "mem"},
              {"id": "BallFollowerBot", "type": "HardcodedBallFollower", "path": "N/A"}]
    nets = {"q0": q0, "q1": q1, "rnn": r0, "BallFollowerBot": None}
    return models, {m["id"]: ck.Agent(m, nets[m["id"]]) for m in models}


def _oracle_scores(cfg, agents, ida, idb, n, seed):
    """The same pairing on the oracle: env i plays the Philox serve (seed, i, episode 0); quota 1."""
    serves = np.array([po.philox_serve(seed, i, 0, cfg) for i in range(n)])
    pool = tuple(serves[:, k][None, :].copy() for k in range(3))
    b = po.EnvBatch(n, "f64")
    b.serve(pool[0][0], pool[1][0], pool[2][0])

    def pol(mid):
        if agents[mid].type == "HardcodedBallFollower":
            return po.make_policy(po.POLICY_FOLLOWER)
        return po.make_policy(po.POLICY_QNET, po.qnet_weights_from_state_dict(agents[mid].net.state_dict()))
    w = po.selfplay(po.make_params(cfg), b, pol(ida), pol(idb), 1 << 14, pool, quota=1, log_cap=n)
    log = w["ep_log"][np.argsort(w["ep_log"][:, 0], kind="stable")]
    assert len(log) == n
    return log[:, 2] >> 16, log[:, 2] & 0xffff, log[:, 3]


def test_round_robin_tournament_on_the_device(H, tmp_path):
    cfg = H["env_config_yaml"]
    models, agents = _agents()
    db_path = tmp_path / "arena_database.json"
    db = arena.load_database(db_path)
    assert arena.register_models(db, models)
    plan = arena.create_match_plan(db, 37)
    assert len(plan) == 6 and all(m["episodes_to_run"] == 37 for m in plan)
    res = arena.run_tournament(cfg, db, db_path, plan, agents=agents, seed=100, concurrent=3)
    assert len(db["match_history"]) == 6 * 37 and arena.load_database(db_path) == db
    assert arena.create_match_plan(db, 37) == []
    for k, m in enumerate(plan):
        ida, idb = m["p1_id"], m["p2_id"]
        sa, sb, ln = res[(ida, idb)]
        assert np.all(np.maximum(sa, sb) == cfg["max_score"]) and np.all(np.minimum(sa, sb) < cfg["max_score"]) and np.all(ln > 0)
        recs = [r for r in db["match_history"] if (r["p1"], r["p2"]) == (ida, idb)]
        assert [r["p1_score"] for r in recs] == sa.tolist() and [r["p2_score"] for r in recs] == sb.tolist()
        assert all(r["winner"] == (ida if r["p1_score"] > r["p2_score"] else idb) for r in recs)
        if "rnn" not in (ida, idb):                                            # fp32 fmaf chain: the oracle's very games
            wa, wb, wl = _oracle_scores(cfg, agents, ida, idb, 37, 100 + k)
            assert np.array_equal(sa, wa) and np.array_equal(sb, wb) and np.array_equal(ln, wl), (ida, idb)
    summary = {r["model_id"]: r for r in arena.generate_summary_report(db)}
    assert sum(r["win"] for r in summary.values()) == 6 * 37 and all(r["games_played"] == 3 * 37 for r in summary.values())
    # resume rule: a longer target only tops the pairings up (tests/arena.py:236-238)
    plan2 = arena.create_match_plan(db, 40)
    assert len(plan2) == 6 and all(m["episodes_to_run"] == 3 for m in plan2)
    arena.run_tournament(cfg, db, db_path, plan2, agents=agents, seed=900, concurrent=2)
    assert len(db["match_history"]) == 6 * 40 and arena.create_match_plan(db, 40) == []
    # ... and a top-up continues the pairing's serve sequence: its games are NEW games (Philox serve (seed + pair, g, 0)
    # with g counting from the games already recorded), not copies of the first ones
    sa, sb, ln = arena.play_match(cfg, agents["q0"], agents["q1"], 3, seed=100, first_game=37)
    wa, wb, wl = _oracle_scores(cfg, agents, "q0", "q1", 40, 100)
    assert np.array_equal(sa, wa[37:]) and np.array_equal(sb, wb[37:]) and np.array_equal(ln, wl[37:])
    assert not np.array_equal(ln, wl[:3])
    # the same plan and seed replay the same games, whatever the number of pairings in flight
    db_b = {"models": list(models), "match_history": []}
    res_b = arena.run_tournament(cfg, db_b, None, plan, agents=agents, seed=100, concurrent=1)
    for key in res:
        assert all(np.array_equal(x, y) for x, y in zip(res[key], res_b[key])), key


def test_tournament_on_the_tensor_core_paths(H):
    """precision f16: QNet x QNet on tcgen05, QNetRNN pairings on the tensor-core recurrent kernel with the QNet / bot
    on its CUDA cores.  Outcomes are valid games; the QNet x QNet games equal the fp32 ones except at near-ties."""
    cfg = H["env_config_yaml"]
    models, agents = _agents()
    db = {"models": list(models), "match_history": []}
    plan = arena.create_match_plan(db, 150)
    res16 = arena.run_tournament(cfg, db, None, plan, agents=agents, seed=31, precision="f16", concurrent=4)
    assert len(db["match_history"]) == 6 * 150
    for (ida, idb), (sa, sb, ln) in res16.items():
        assert np.all(np.maximum(sa, sb) == cfg["max_score"]) and np.all(np.minimum(sa, sb) < cfg["max_score"])
    k = next(i for i, m in enumerate(plan) if (m["p1_id"], m["p2_id"]) == ("q0", "q1"))
    wa, wb, _ = _oracle_scores(cfg, agents, "q0", "q1", 150, 31 + k)
    sa, sb, _ = res16[("q0", "q1")]
    assert (np.sign(sa - sb) == np.sign(wa - wb)).mean() > 0.95
    k = next(i for i, m in enumerate(plan) if (m["p1_id"], m["p2_id"]) == ("q1", "BallFollowerBot"))
    wa, wb, _ = _oracle_scores(cfg, agents, "q1", "BallFollowerBot", 150, 31 + k)
    sa, sb, _ = res16[("q1", "BallFollowerBot")]
    assert (np.sign(sa - sb) == np.sign(wa - wb)).mean() > 0.95


def test_play_match_with_a_reference_shaped_agent_record(H):
    """play_match with agents built from state_dicts in the reference's formats (legacy fc.* and dueling)."""
    cfg = H["env_config_yaml"]
    g = torch.Generator().manual_seed(5)
    legacy = {"fc.0.weight": torch.randn(64, 7, generator=g) * 0.5, "fc.0.bias": torch.zeros(64),
              "fc.2.weight": torch.randn(64, 64, generator=g) * 0.2, "fc.2.bias": torch.zeros(64),
              "fc.4.weight": torch.randn(3, 64, generator=g) * 0.3, "fc.4.bias": torch.zeros(3)}
    old = ck.Agent({"id": "old", "type": "QNet", "path": "mem"}, ck.qnet_from_state_dict(legacy))
    torch.manual_seed(2)
    new = ck.Agent({"id": "new", "type": "QNet", "path": "mem"}, ck.qnet_from_state_dict(pp.QNet().state_dict()))
    sa, sb, ln = arena.play_match(cfg, old, new, 64, seed=9)
    wa, wb, wl = _oracle_scores(cfg, {"old": old, "new": new}, "old", "new", 64, 9)
    assert np.array_equal(sa, wa) and np.array_equal(sb, wb) and np.array_equal(ln, wl)


def test_eval_vs_model_and_eval_vs_pool_on_the_batched_engine(H):
    """scripts/train_iterative.py:171-196 as lock-step batches: win rates equal the oracle's on the same serves."""
    import random
    cfg = H["env_config_yaml"]
    models, agents = _agents()
    q0, q1 = agents["q0"].net, agents["q1"].net
    wr = pp.eval_vs_model(cfg, q0, q1, 200, seed=12)
    wa, wb, _ = _oracle_scores(cfg, agents, "q0", "q1", 200, 12)
    assert wr == float((wb > wa).mean())
    assert pp.eval_vs_model(cfg, q0, q1, 300, max_envs=128, seed=3) == pytest.approx(wr, abs=0.12)   # 128 envs x 3 games
    assert pp.eval_vs_pool(cfg, q1, [], 50) == 1.0                                                    # :184-185
    torch.manual_seed(5); q2 = pp.QNet().eval()
    agents2 = dict(agents, q2=ck.Agent({"id": "q2", "type": "QNet", "path": "mem"}, q2))
    got = pp.eval_vs_pool(cfg, q1, [q0, q2], 150, seed=40, rng=random.Random(9))
    rng = random.Random(9)
    games = [0, 0]
    for _ in range(150):
        games[rng.choice(range(2))] += 1
    wins = 0
    for k, (opp, g) in enumerate(zip(("q0", "q2"), games)):
        wa, wb, _ = _oracle_scores(cfg, agents2, opp, "q1", g, 40 + k)
        wins += int((wb > wa).sum())
    assert games[0] > 0 and games[1] > 0 and got == wins / 150
