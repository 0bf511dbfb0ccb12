"""Round-robin tournament on the batched engine: the front-end of the reference's
tests/arena.py with its `arena_database.json` schema, where every pairing is ONE lock-step batch on the device
instead of a Python loop over episodes.

    reference (tests/arena.py)                          here
    load_database / save_database / register_models     same names, same JSON  (:127-155)
    create_match_plan                         :222-245  create_match_plan (same pairing order and resume rule)
    run_tournament                            :247-320  run_tournament -> play_match: `episodes_to_run` envs, quota 1,
                                                        one launch of the fused self-play kernel per pairing
    generate_summary_report                   :323-352  generate_summary_report (rows sorted by win rate)
    plot_h2h_heatmap                          :354-377  h2h_wins (the matrix; plotting stays in the reference)

What differs, by construction: the reference plays the episodes of a pairing one after the other from the process-wide
`random` stream and saves the database after every episode; here the episodes of a pairing run side by side, serves come
from the device Philox stream (seed, episode index) and the database is saved once per pairing.  A pairing whose history
is incomplete is topped up exactly as in the reference (episodes_per_match - played).  Records carry the same fields.
"""
from __future__ import annotations

import itertools
import json
import os
from collections import Counter
from datetime import datetime, timezone

import numpy as np
import torch

from .checkpoint import Agent, load_agent
from .env import VecPongEnv2P
from .selfplay import SelfPlayEngine


# ------------------------------------------------------------------------------------------ database (host logic)
def load_database(db_path) -> dict:
    """tests/arena.py:127-139 — a missing, empty or corrupt file starts a new database."""
    if os.path.exists(db_path) and os.path.getsize(db_path) > 0:
        with open(db_path, "r", encoding="utf-8") as f:
            try:
                data = json.load(f)
            except json.JSONDecodeError:
                return {"models": [], "match_history": []}
        data.setdefault("models", [])
        data.setdefault("match_history", [])
        return data
    return {"models": [], "match_history": []}


def save_database(db_path, data: dict) -> None:
    with open(db_path, "w", encoding="utf-8") as f:
        json.dump(data, f, indent=2, ensure_ascii=False)


def register_models(database: dict, candidates: list) -> bool:
    """Append the candidates whose id is not registered yet; True if any was new (:147-155)."""
    known = {m["id"] for m in database["models"]}
    added = False
    for cand in candidates:
        if cand["id"] not in known:
            database["models"].append(cand)
            known.add(cand["id"])
            added = True
    return added


def create_match_plan(database: dict, episodes_per_match: int) -> list:
    """All unordered pairs of registered models, in registration order, that still miss episodes (:222-245)."""
    ids = [m["id"] for m in database["models"]]
    played = Counter(tuple(sorted((r["p1"], r["p2"]))) for r in database["match_history"])
    plan = []
    for p1, p2 in itertools.combinations(ids, 2):
        todo = episodes_per_match - played[tuple(sorted((p1, p2)))]
        if todo > 0:
            plan.append({"p1_id": p1, "p2_id": p2, "episodes_to_run": todo})
    return plan


def generate_summary_report(database: dict) -> list:
    """Per-model win / lose / draw / games_played / win_rate, best win rate first (:323-352).  A list of dicts; wrap it
    in pandas.DataFrame(...).set_index("model_id") for the reference's table."""
    stats = {m["id"]: {"win": 0, "lose": 0, "draw": 0} for m in database["models"]}
    for r in database["match_history"]:
        p1, p2, w = r["p1"], r["p2"], r["winner"]
        if w == "draw":
            stats[p1]["draw"] += 1; stats[p2]["draw"] += 1
        elif w == p1:
            stats[p1]["win"] += 1; stats[p2]["lose"] += 1
        elif w == p2:
            stats[p2]["win"] += 1; stats[p1]["lose"] += 1
    rows = []
    for mid, s in stats.items():
        games = s["win"] + s["lose"] + s["draw"]
        rows.append({"model_id": mid, **s, "games_played": games, "win_rate": s["win"] / games if games else 0})
    rows.sort(key=lambda r: -r["win_rate"])              # stable: ties keep registration order, like sort_values
    return rows


def h2h_wins(database: dict):
    """(ids, wins[winner, loser]) — the matrix behind the reference's heat map (:354-365)."""
    ids = [m["id"] for m in database["models"]]
    at = {mid: k for k, mid in enumerate(ids)}
    wins = np.zeros((len(ids), len(ids)), np.int64)
    for r in database["match_history"]:
        w = r.get("winner")
        if w != "draw":
            loser = r["p2"] if w == r["p1"] else r["p1"]
            wins[at[w], at[loser]] += 1
    return ids, wins


# ------------------------------------------------------------------------------------------ matches on the device
class _Match:
    """One pairing in flight: its own env slab, players, stream and episode log."""

    def __init__(self, env_cfg, agent_a: Agent, agent_b: Agent, episodes: int, seed: int, precision, mode, device, stream,
                 first_game: int = 0):
        """first_game: games of this pairing already in the database.  Game g of a pairing is served from Philox
        (seed, g, episode 0), so a top-up (first_game = played) continues the pairing's serve sequence instead of
        replaying games that are already recorded."""
        self.a, self.b, self.n, self.stream = agent_a, agent_b, int(episodes), stream
        with torch.cuda.stream(stream):
            self.env = VecPongEnv2P(self.n, device=device, mode=mode, serve="philox", seed=seed, env_id_base=int(first_game),
                                    **env_cfg)
            self.env.reset()
            self.log = torch.zeros(self.n, 4, dtype=torch.int32, device=device)
            self.engine = SelfPlayEngine(self.env, agent_a.policy(self.n, precision, device),
                                         agent_b.policy(self.n, precision, device), seed=seed)

    def launch(self, max_steps: int):
        with torch.cuda.stream(self.stream):              # quota 1: an env freezes when its game ends; the kernel
            self.engine.run(max_steps, quota=1, ep_log=self.log)   # leaves its step loop once all of them have

    def results(self):
        """-> (score_a[n], score_b[n], ep_len[n]) in env order (after the stream has been synchronised)."""
        self.stream.synchronize()
        done = self.env.ep_log_count()
        if done != self.n:
            raise RuntimeError(f"{self.a.id} vs {self.b.id}: {done} of {self.n} games finished within the step limit")
        log = self.log.cpu().numpy()
        log = log[np.argsort(log[:, 0], kind="stable")]
        return log[:, 2] >> 16, log[:, 2] & 0xffff, log[:, 3]


def play_match(env_cfg: dict, agent_a: Agent, agent_b: Agent, episodes: int, seed: int = 0, precision: str = "f32",
               mode: str = "f64", device="cuda", max_steps: int = 1 << 20, first_game: int = 0):
    """`episodes` games of A (top paddle) against B (bottom paddle), all in one launch.  -> (score_a, score_b, ep_len)."""
    m = _Match(env_cfg, agent_a, agent_b, episodes, seed, precision, mode, torch.device(device), torch.cuda.Stream(device),
               first_game=first_game)
    m.launch(max_steps)
    return m.results()


def _records(id_a, id_b, score_a, score_b, stamp):
    out = []
    for sa, sb in zip(score_a.tolist(), score_b.tolist()):
        winner = id_a if sa > sb else (id_b if sb > sa else "draw")          # tests/arena.py:306-308
        out.append({"p1": id_a, "p2": id_b, "winner": winner, "p1_score": int(sa), "p2_score": int(sb), "timestamp": stamp})
    return out


def run_tournament(env_cfg: dict, database: dict, db_path, match_plan: list, rnn_arch: dict | None = None,
                   device="cuda", precision: str = "f32", mode: str = "f64", seed: int = 0, concurrent: int = 8,
                   root: str = ".", agents: dict | None = None, max_steps: int = 1 << 20, on_error=print,
                   shard: tuple | None = None, match_factory=None) -> dict:
    """Play every pairing of `match_plan`, append one record per game to database['match_history'] and save the database
    after each pairing.  Up to `concurrent` pairings are in flight on their own CUDA streams (a pairing of 100 games
    fills one SM).  `agents` may carry already loaded Agent objects by id; the others are loaded from the database's
    model records, and a model that fails to load only cancels its own pairings, as in the reference (:268-289).
    Serves: the pairing with index j among ALL pairs of registered models (registration order, i.e. the order of a
    first full plan) plays with seed + j, and its game g takes the Philox serve (seed + j, g, 0) where g counts the
    pairing's games already in the database — a top-up after raising episodes_per_match plays NEW serves, like the
    reference's process-wide RNG does on resume, instead of replaying recorded games.
    Several ranks (torch.distributed initialised, or shard=(rank, world)): pairing k is played by rank k % world with
    the same seed as on one GPU, results are gathered, EVERY rank extends its database identically (plan order; timestamps aside) and
    rank 0 alone writes the file — pairings are independent, so there is no data-path collective.
    Returns {pair: (score_a, score_b, ep_len)}."""
    from . import dist as ppd
    import torch.distributed as tdist
    if shard is None and ppd.is_parallel():
        shard = (tdist.get_rank(), tdist.get_world_size())
    rank, world = shard if shard is not None else (0, 1)
    make = match_factory or _Match
    device = torch.device(device)
    info = {m["id"]: m for m in database["models"]}
    agents = dict(agents or {})
    for mid in sorted({m["p1_id"] for m in match_plan} | {m["p2_id"] for m in match_plan}):
        if mid in agents:
            continue
        try:
            agents[mid] = load_agent(info[mid], rnn_arch, root)
        except Exception as e:                                                # noqa: BLE001 — reference behaviour
            on_error(f"[arena] loading model {mid!r} failed: {e}")
    pair_index = {pair: j for j, pair in enumerate(itertools.combinations([m["id"] for m in database["models"]], 2))}
    played = Counter(tuple(sorted((r["p1"], r["p2"]))) for r in database["match_history"])
    results, flight = {}, []
    streams = [torch.cuda.Stream(device) for _ in range(max(1, int(concurrent)))] if match_factory is None else [object() for _ in range(max(1, int(concurrent)))]

    def record(ida, idb, sa, sb):
        stamp = datetime.now(timezone.utc).strftime("%Y-%m-%dT%H:%M:%S.%f") + "Z"
        database["match_history"].extend(_records(ida, idb, sa, sb, stamp))

    def land(m):
        sa, sb, ln = m.results()
        results[(m.a.id, m.b.id)] = (sa, sb, ln)
        if world == 1:                                                        # one process: record and save as we go
            record(m.a.id, m.b.id, sa, sb)
            if db_path is not None:
                save_database(db_path, database)

    for k, match in enumerate(match_plan):
        ida, idb = match["p1_id"], match["p2_id"]
        if ida not in agents or idb not in agents:
            if rank == 0:
                on_error(f"[arena] skipping {ida} vs {idb}: a model failed to load")
            continue
        if k % world != rank:
            continue
        if len(flight) == len(streams):
            land(flight.pop(0))
        used = {id(m.stream) for m in flight}
        stream = next(s for s in streams if id(s) not in used)
        m = make(env_cfg, agents[ida], agents[idb], match["episodes_to_run"], seed + pair_index.get((ida, idb), k), precision,
                 mode, device, stream, first_game=played[tuple(sorted((ida, idb)))])
        m.launch(max_steps)
        flight.append(m)
    while flight:
        land(flight.pop(0))
    if world > 1:
        mine = {k: tuple(np.asarray(v) for v in val) for k, val in results.items()}
        if tdist.is_available() and tdist.is_initialized():
            parts = [None] * world
            tdist.all_gather_object(parts, mine)
        else:                                                                 # shard=(rank, world) without a process group:
            parts = [mine]                                                    # the caller merges the ranks' return values
        results = {}
        for part in parts:
            results.update(part)
        for match in match_plan:                                              # plan order: identical on every rank
            key = (match["p1_id"], match["p2_id"])
            if key in results:
                record(key[0], key[1], results[key][0], results[key][1])
        if db_path is not None and rank == 0:
            save_database(db_path, database)
    return results
