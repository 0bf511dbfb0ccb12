"""CPU tests of the section-8f host front-ends: checkpoint schema compatibility and the arena database logic, against
golden outputs of the reference's own loader / arena helpers (tests/golden/ckpt_golden.npz, arena_golden.json; made by
`python -m oracle.gen_golden ckpt|arena` in the build container)."""
import copy
import json
import os

import numpy as np
import pytest
import torch

from pingpong_selfplay_ai_b200 import arena, checkpoint as ck
from pingpong_selfplay_ai_b200.policy import QNet, QNetRNN

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF = os.environ.get("PP_REFERENCE_ROOT", "/root/reference")


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(GOLD, "ckpt_golden.npz"))


def _legacy_sd(G):
    return {k[len("legacy/sd/"):]: torch.from_numpy(G[k]) for k in G.files if k.startswith("legacy/sd/")}


def test_legacy_fc_state_dict_maps_onto_the_dueling_net_like_the_reference_loader(G):
    sd = _legacy_sd(G)
    assert ck.is_legacy_qnet(sd)
    net = ck.qnet_from_state_dict(sd)
    assert not net.training
    with torch.no_grad():
        q = net(torch.from_numpy(G["obs"])).numpy()
    assert np.abs(q - G["legacy/q"]).max() <= 1e-6
    # the mapping is Q-preserving: the dueling combine returns the old 3-way output layer
    x = torch.from_numpy(G["obs"])
    h = torch.relu(torch.relu(x @ sd["fc.0.weight"].t() + sd["fc.0.bias"]) @ sd["fc.2.weight"].t() + sd["fc.2.bias"])
    assert np.abs(q - (h @ sd["fc.4.weight"].t() + sd["fc.4.bias"]).numpy()).max() < 1e-5


def test_extract_state_dict_key_order_and_bare_dicts(G):
    sd = _legacy_sd(G)
    a, b = {"features.0.weight": torch.zeros(1)}, {"features.0.weight": torch.ones(1)}
    assert ck.extract_state_dict({"model": sd, "modelA": a, "modelB": b}) is b              # modelB before modelA before model
    assert ck.extract_state_dict({"modelA_state": a, "modelB_state": b, "modelB": sd}) is b
    assert ck.extract_state_dict({"model": sd, "modelB": b}, order=ck.TRAIN_KEYS) is b       # train_iterative.py:87
    assert ck.extract_state_dict({"model": sd, "epsilon": 0.1}, order=ck.TRAIN_KEYS) is sd
    assert ck.extract_state_dict(sd) is sd                                                   # bare state_dict
    with pytest.raises(KeyError):
        ck.extract_state_dict({"epsilon": 0.1, "episode": 3})
    with pytest.raises(KeyError):
        ck.remap_legacy_qnet({"fc.0.weight": sd["fc.0.weight"]})


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "checkpoints")), reason="reference checkpoints live in the build container only")
@pytest.mark.parametrize("tag", ["legacy_file", "dueling_file", "rnn_file"])
def test_reference_checkpoint_files_load_like_the_reference_loader(G, tag):
    info = {"id": tag, "type": str(G[tag + "/type"]), "path": str(G[tag + "/path"])}
    agent = ck.load_agent(info, {"feature_dim": 128, "lstm_hidden_dim": 128, "lstm_layers": 1, "head_hidden_dim": 128}, root=REF)
    obs = torch.from_numpy(G["obs"])
    with torch.no_grad():
        if agent.type == "QNet":
            q = agent.net(obs)
        else:
            h = agent.net.init_hidden(obs.shape[0], torch.device("cpu"))
            _, h = agent.net(torch.from_numpy(G["obs"][::-1].copy()).unsqueeze(1), h)
            q, _ = agent.net(obs.unsqueeze(1), h)
    assert np.abs(q.numpy() - G[tag + "/q"]).max() <= 1e-6


def test_load_agent_errors_and_the_hardcoded_bot(tmp_path):
    bot = ck.load_agent({"id": "BallFollowerBot", "type": "HardcodedBallFollower", "path": "N/A"})
    assert bot.net is None and bot.type == "HardcodedBallFollower"
    with pytest.raises(FileNotFoundError):
        ck.load_agent({"id": "x", "type": "QNet", "path": "nope.pth"}, root=str(tmp_path))
    with pytest.raises(ValueError):
        ck.load_agent({"id": "x", "type": "Transformer", "path": "nope.pth"})
    with pytest.raises(ValueError):
        ck.qnetrnn_from_state_dict({}, {"lstm_hidden_dim": 256})


def test_generation_files_use_the_reference_schemas_and_round_trip(tmp_path):
    torch.manual_seed(3)
    a, b = QNet(), QNet()
    opt = torch.optim.Adam(list(b.fc_V.parameters()) + list(b.fc_A.parameters()), lr=2.5e-4)
    fn = ck.save_qnet_generation(str(tmp_path), 5, 2, a, b, opt, 0.37, 1234, fault=True)
    assert os.path.basename(fn) == "model5-2_fault.pth"
    assert os.path.basename(ck.save_qnet_generation(str(tmp_path), 5, 3, a, b, opt, 0.3, 1300)) == "model5-3.pth"
    whole, sd = ck.load_checkpoint(fn, order=ck.TRAIN_KEYS)
    assert list(whole.keys()) == ["modelB", "optimizer", "epsilon", "episode", "modelA"]       # train_iterative.py:286-292
    assert whole["epsilon"] == 0.37 and whole["episode"] == 1234
    assert all(torch.equal(sd[k], v) for k, v in b.state_dict().items())
    net = ck.qnet_from_state_dict(sd)                                                        # strict load of the new format
    x = torch.rand(5, 7)
    b.eval()
    assert torch.equal(net(x), b(x))
    ra, rb = QNetRNN(), QNetRNN()
    fn = ck.save_qnetrnn_generation(str(tmp_path), "rnn_agent_", 4, ra, rb, opt, 0.1, 99, train_steps_count=17)
    whole, sd = ck.load_checkpoint(fn)
    assert os.path.basename(fn) == "rnn_agent_4.pth"
    assert list(whole.keys()) == ["modelA_state", "modelB_state", "optimizer_B_state", "epsilon", "episode", "generation",
                                  "train_steps_count", "old_state_for_reset"]                # train_rnn_iterative.py:841-850
    assert all(torch.equal(sd[k], v) for k, v in rb.state_dict().items())                     # modelB_state is tried first
    ck.qnetrnn_from_state_dict(sd)
    if os.path.isdir(os.path.join(REF, "checkpoints")):                                      # same keys as a real reference file
        real = torch.load(os.path.join(REF, "checkpoints", "model5-3_fault.pth"), map_location="cpu", weights_only=True)
        assert set(real.keys()) == {"modelB", "optimizer", "epsilon", "episode", "modelA"}
        assert set(real["modelB"].keys()) == set(b.state_dict().keys())


# ------------------------------------------------------------------------------------------ arena database logic
@pytest.fixture(scope="module")
def A():
    with open(os.path.join(GOLD, "arena_golden.json")) as f:
        return json.load(f)


def test_match_plan_and_summary_equal_the_reference_helpers(A):
    db = copy.deepcopy(A["database"])
    assert arena.create_match_plan(db, 5) == A["plan_5"]
    got = arena.generate_summary_report(db)
    want = A["summary"]
    assert sorted(got, key=lambda r: r["model_id"]) == sorted(want, key=lambda r: r["model_id"])
    assert [r["win_rate"] for r in got] == [r["win_rate"] for r in want]                    # best first (ties: any order)
    assert db == A["database"]                                                               # nothing mutated
    assert arena.create_match_plan(db, 0) == []
    db2 = copy.deepcopy(db)
    added = arena.register_models(db2, [{"id": "m1", "type": "QNet", "path": "x"}, {"id": "new", "type": "QNetRNN", "path": "y"}])
    assert added == A["register_added"] and [m["id"] for m in db2["models"]] == A["register_ids"]
    assert not arena.register_models(db2, [{"id": "new", "type": "QNetRNN", "path": "y"}])


def test_h2h_matrix_and_database_files(A, tmp_path):
    db = copy.deepcopy(A["database"])
    ids, wins = arena.h2h_wins(db)
    decided = [r for r in db["match_history"] if r["winner"] != "draw"]
    assert wins.sum() == len(decided) and np.all(np.diag(wins) == 0)
    summary = {r["model_id"]: r for r in arena.generate_summary_report(db)}
    for k, mid in enumerate(ids):
        assert wins[k].sum() == summary[mid]["win"] and wins[:, k].sum() == summary[mid]["lose"]
    path = tmp_path / "arena_database.json"
    assert arena.load_database(path) == {"models": [], "match_history": []}                  # missing file
    path.write_text("{ not json")
    assert arena.load_database(path) == {"models": [], "match_history": []}                  # corrupt file
    path.write_text(json.dumps({"models": db["models"]}))
    assert arena.load_database(path)["match_history"] == []                                  # missing section
    arena.save_database(path, db)
    assert arena.load_database(path) == db
    recs = arena._records("a", "b", np.array([3, 1, 2]), np.array([0, 3, 2]), "t")
    assert [r["winner"] for r in recs] == ["a", "b", "draw"]
    assert list(recs[0].keys()) == ["p1", "p2", "winner", "p1_score", "p2_score", "timestamp"]   # tests/arena.py:311-318


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "arena_database.json")), reason="the reference tree lives in the build container only")
def test_the_reference_s_own_database_gives_the_reference_s_own_report():
    """arena_database.json as shipped by the reference (10 models, 4 500 games): our plan / summary / head-to-head equal
    the outputs of the unmodified tests/arena.py functions on it."""
    from oracle import ref_shim
    ref_arena = ref_shim.load_reference_round_robin("arena.py")
    with open(os.path.join(REF, "arena_database.json"), encoding="utf-8") as f:
        db = json.load(f)
    assert arena.load_database(os.path.join(REF, "arena_database.json")) == db
    for target in (100, 120):
        assert arena.create_match_plan(copy.deepcopy(db), target) == ref_arena.create_match_plan(copy.deepcopy(db), target)
    want = ref_arena.generate_summary_report(copy.deepcopy(db)).reset_index().to_dict(orient="records")
    got = arena.generate_summary_report(db)
    assert sorted(got, key=lambda r: r["model_id"]) == sorted(want, key=lambda r: r["model_id"])
    assert [r["win_rate"] for r in got] == [r["win_rate"] for r in want]
    ids, wins = arena.h2h_wins(db)
    assert wins.sum() == sum(r["win"] for r in got) == 4500 - sum(r["draw"] for r in got) // 2
