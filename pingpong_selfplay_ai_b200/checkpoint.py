"""Checkpoint schema compatibility: every on-disk format the reference reads can drive
the engine, and what a generation produces can be written back in the reference's own schemas.

    reference                                                       here
    load_model_universal   tests/test_round_robin.py:116-185       load_agent / extract_state_dict / remap_legacy_qnet
                           tests/arena.py:158-197
    base checkpoint        scripts/train_iterative.py:86-87        extract_state_dict(order=TRAIN_KEYS)
    promote / fault files  scripts/train_iterative.py:272-278,     save_qnet_generation
                           286-292
    RNN promote files      scripts/train_rnn_iterative.py:841-850  save_qnetrnn_generation

On-disk formats (all present under the reference's checkpoints/ and checkpoints_rnn/):
  1. legacy QNet: a plain 3-layer MLP saved as `fc.0.* / fc.2.* / fc.4.*` under 'model' (and 'modelA'/'modelB');
     `fc.4` (3 x 64) is mapped onto the dueling head so that Q is unchanged: A = fc.4, V = mean over its rows, hence
     V + (A - mean A) = A (tests/test_round_robin.py:155-164).  Sigma / epsilon keep their construction values: they
     do not enter an eval-mode forward.
  2. dueling NoisyNet QNet: `features.* / fc_V.* / fc_A.*` under 'modelB' / 'modelA'.
  3. QNetRNN: `features_extractor.* / lstm.* / fc_shared_head.0.* / fc_V.* / fc_A.*` under 'modelB_state' / 'modelA_state'.
A bare state_dict (keys starting with `fc.` / `features`) is accepted as well (test_round_robin.py:145-147).

tests/arena.py loads a legacy file with `load_state_dict(strict=False)` and NO remap (arena.py:185-187), which silently
leaves a randomly initialised net; this module follows tests/test_round_robin.py, where the file's weights are used.
"""
from __future__ import annotations

import os

import torch

from .policy import Policy, QNet, QNetRNN

# first match wins — tests/test_round_robin.py:137, tests/arena.py:171
EVAL_KEYS = ("modelB_state", "modelA_state", "modelB", "modelA", "model", "state_dict")
# scripts/train_iterative.py:87: base_cp.get('modelB', base_cp.get('model'))
TRAIN_KEYS = ("modelB", "model")

AGENT_TYPES = ("QNet", "QNetRNN", "HardcodedBallFollower")


def _looks_like_state_dict(obj) -> bool:
    return (isinstance(obj, dict) and len(obj) > 0 and all(not isinstance(v, dict) for v in obj.values())
            and any(str(k).startswith(("fc.", "fc_", "features", "lstm.", "fc_shared_head.")) for k in obj))


def extract_state_dict(ckpt, order=EVAL_KEYS) -> dict:
    """The model state_dict inside a loaded checkpoint object."""
    if not isinstance(ckpt, dict):
        raise KeyError("checkpoint is not a dict")
    for key in order:
        if key in ckpt:
            return ckpt[key]
    if _looks_like_state_dict(ckpt):
        return ckpt
    raise KeyError(f"no model state_dict in the checkpoint (tried {list(order)}; keys: {list(ckpt.keys())[:8]})")


def is_legacy_qnet(state_dict) -> bool:
    return not any(k.startswith(("features.", "fc_V.", "fc_A.")) for k in state_dict)


def remap_legacy_qnet(state_dict) -> dict:
    """`fc.*` -> dueling keys (mu only), Q-preserving.                          tests/test_round_robin.py:155-164"""
    out = {}
    for k, v in state_dict.items():
        if k.startswith("fc.0."):
            out["features.0." + k[len("fc.0."):]] = v
        elif k.startswith("fc.2."):
            out["features.2." + k[len("fc.2."):]] = v
    if "fc.4.weight" not in state_dict or "fc.4.bias" not in state_dict:
        raise KeyError("legacy QNet state_dict without fc.4.weight / fc.4.bias")
    w4, b4 = state_dict["fc.4.weight"], state_dict["fc.4.bias"]
    out["fc_A.weight_mu"], out["fc_A.bias_mu"] = w4, b4
    out["fc_V.weight_mu"], out["fc_V.bias_mu"] = w4.mean(dim=0, keepdim=True), b4.mean().unsqueeze(0)
    return out


def qnet_from_state_dict(state_dict) -> QNet:
    net = QNet(input_dim=7, output_dim=3)
    if is_legacy_qnet(state_dict):
        net.load_state_dict(remap_legacy_qnet(state_dict), strict=False)
    else:
        net.load_state_dict(state_dict, strict=True)
    net.eval()
    return net


def qnetrnn_from_state_dict(state_dict, rnn_arch: dict | None = None) -> QNetRNN:
    a = rnn_arch or {}
    dims = (a.get("feature_dim", 128), a.get("lstm_hidden_dim", 128), a.get("lstm_layers", 1), a.get("head_hidden_dim", 128))
    if dims != (128, 128, 1, 128):
        raise ValueError(f"the device kernels are built for the reference's QNetRNN dims 128/128/1/128 (config_rnn.yaml:39-42), got {dims}")
    net = QNetRNN(input_dim=7, output_dim=3, feature_dim=dims[0], lstm_hidden_dim=dims[1], lstm_layers=dims[2],
                  head_hidden_dim=dims[3])
    net.load_state_dict(state_dict)
    net.eval()
    return net


def load_checkpoint(path, order=EVAL_KEYS) -> tuple[dict, dict]:
    """-> (whole checkpoint object, model state_dict).  `weights_only=True`: every reference file loads that way."""
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    ckpt = torch.load(path, map_location="cpu", weights_only=True)
    return ckpt, extract_state_dict(ckpt, order)


class Agent:
    """One tournament entrant: the reference's model-info record plus the loaded net (None for the hard-coded bot)."""

    def __init__(self, info: dict, net=None):
        if info["type"] not in AGENT_TYPES:
            raise ValueError(f"unsupported model type {info['type']!r} (model id {info.get('id')})")
        self.info, self.net = dict(info), net
        self.id, self.type = info["id"], info["type"]
        self._packed = {}

    def policy(self, num_envs: int, precision: str = "f32", device="cuda", eps: float = 0.0) -> Policy:
        """A player for `num_envs` lock-step games (eval mode: mu weights, arena.py:196).  The packed weights are built
        once per (precision, device) and shared between the matches of a tournament; (h, c) are per match."""
        if self.type == "HardcodedBallFollower":
            return Policy.follower(tol=0.02, eps=eps, device=device)
        key = (precision, str(device))
        if key not in self._packed:
            mk = Policy.qnet if self.type == "QNet" else Policy.qnetrnn
            kw = {"num_envs": 1} if self.type == "QNetRNN" else {}
            self._packed[key] = mk(self.net, precision=precision, device=device, **kw).weights
        kind_kw = {"num_envs": num_envs} if self.type == "QNetRNN" else {}
        from . import _lib
        pol = Policy(_lib.POLICY_QNET if self.type == "QNet" else _lib.POLICY_QNETRNN, None, eps=eps, precision=precision,
                     device=device, **kind_kw)
        pol.weights = self._packed[key]
        return pol


def load_agent(model_info: dict, rnn_arch: dict | None = None, root: str = ".") -> Agent:
    """model_info = {"id", "type", "path", ...} as in ARENA_CONFIG["candidate_models"] (tests/arena.py:60-118)."""
    t = model_info["type"]
    if t == "HardcodedBallFollower":
        return Agent(model_info)
    if t not in AGENT_TYPES:
        raise ValueError(f"unsupported model type {t!r} (model id {model_info.get('id')})")
    path = model_info["path"]
    path = path if os.path.isabs(path) else os.path.join(root, path)
    _, sd = load_checkpoint(path)
    return Agent(model_info, qnet_from_state_dict(sd) if t == "QNet" else qnetrnn_from_state_dict(sd, rnn_arch))


def _cpu_state(model) -> dict:
    return {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}


def save_qnet_generation(ckpt_dir: str, model_id, generation: int, model_a, model_b, optimizer_b, epsilon: float,
                         episode: int, fault: bool = False) -> str:
    """`model{id}-{gen}.pth` (promoted) or `model{id}-{gen}_fault.pth`         scripts/train_iterative.py:271-292"""
    os.makedirs(ckpt_dir, exist_ok=True)
    fn = os.path.join(ckpt_dir, f"model{model_id}-{generation}{'_fault' if fault else ''}.pth")
    torch.save({"modelB": _cpu_state(model_b), "optimizer": optimizer_b.state_dict(), "epsilon": float(epsilon),
                "episode": int(episode), "modelA": _cpu_state(model_a)}, fn)
    return fn


def save_qnetrnn_generation(ckpt_dir: str, prefix: str, generation: int, model_a, model_b, optimizer_b, epsilon: float,
                            episode: int, train_steps_count: int = 0, old_state_for_reset=None) -> str:
    """`{prefix}{generation}.pth`                                           scripts/train_rnn_iterative.py:839-850"""
    os.makedirs(ckpt_dir, exist_ok=True)
    fn = os.path.join(ckpt_dir, f"{prefix}{generation}.pth")
    a = _cpu_state(model_a)
    torch.save({"modelA_state": a, "modelB_state": _cpu_state(model_b), "optimizer_B_state": optimizer_b.state_dict(),
                "epsilon": float(epsilon), "episode": int(episode), "generation": int(generation),
                "train_steps_count": int(train_steps_count),
                "old_state_for_reset": a if old_state_for_reset is None else old_state_for_reset}, fn)
    return fn
