"""Import shim for the UNMODIFIED reference (test infrastructure, container-only).

TEST INFRASTRUCTURE — only `tests/`, `oracle/gen_golden.py`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` leg may import anything under
`oracle/`.  This particular module additionally needs `/root/reference`, which exists
only in the build container (never on the GPU box), so nothing marked `gpu` uses it.

The reference imports `gym` and `pygame` at module top
(`/root/reference/envs/my_pong_env_2p.py:1-4`); neither is installed here.  `gym.Env`
contributes nothing to the arithmetic (only `super().__init__()` / `super().reset(seed=)`
at `:40,84`), and pygame is used by `render()` only, so two stub modules are enough to
import `envs/` and `models/` unmodified.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PP_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "envs", "my_pong_env_2p.py"))


def _install_stubs() -> None:
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")

        class Env:  # what PongEnv2P needs from gym.Env
            def reset(self, seed=None, options=None):
                return None

        class MultiDiscrete:
            def __init__(self, nvec):
                self.nvec = list(nvec)

        class Box:
            def __init__(self, low, high, dtype=None):
                self.low, self.high, self.dtype = low, high, dtype
                self.shape = low.shape

        spaces = types.ModuleType("gym.spaces")
        spaces.MultiDiscrete = MultiDiscrete
        spaces.Box = Box
        gym.Env = Env
        gym.spaces = spaces
        sys.modules["gym"] = gym
        sys.modules["gym.spaces"] = spaces
    if "pygame" not in sys.modules:
        sys.modules["pygame"] = types.ModuleType("pygame")


def load_reference():
    """Returns (PongEnv2P, collide_sphere_with_moving_plane, QNet, QNetRNN) from the reference."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    # The reference's top-level packages are `envs` and `models` (namespace packages: no __init__.py).  This repository
    # ships import-compatible shims under the same names as REGULAR packages, which win over namespace packages
    # whatever the order of sys.path — so bind the two package names to the reference's directories explicitly.
    import importlib
    for name in [m for m in sys.modules if m.split(".")[0] in ("envs", "models")]:
        mod = sys.modules[name]
        where = getattr(mod, "__file__", None) or next(iter(getattr(mod, "__path__", [])), "")
        if not str(where).startswith(REFERENCE_ROOT):
            del sys.modules[name]
    for pkg in ("envs", "models"):
        if pkg not in sys.modules:
            mod = types.ModuleType(pkg)
            mod.__path__ = [os.path.join(REFERENCE_ROOT, pkg)]
            sys.modules[pkg] = mod
    PongEnv2P = importlib.import_module("envs.my_pong_env_2p").PongEnv2P
    collide_sphere_with_moving_plane = importlib.import_module("envs.physics").collide_sphere_with_moving_plane
    QNet = importlib.import_module("models.qnet").QNet
    QNetRNN = importlib.import_module("models.qnet_rnn").QNetRNN
    assert sys.modules["envs.my_pong_env_2p"].__file__.startswith(REFERENCE_ROOT)
    return PongEnv2P, collide_sphere_with_moving_plane, QNet, QNetRNN


def load_reference_config(name: str = "config.yaml") -> dict:
    import yaml
    with open(os.path.join(REFERENCE_ROOT, name), "r") as f:
        return yaml.safe_load(f)


def load_reference_round_robin(filename: str = "test_round_robin.py"):
    """tests/test_round_robin.py (or tests/arena.py) of the reference as a module, for its load_model_universal /
    select_action_universal / database helpers.  Both import matplotlib and seaborn at module top for their plots;
    neither is installed, neither touches the functions used here."""
    import importlib.util
    load_reference()                                   # gym / pygame stubs + the reference's envs / models in sys.modules
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    spec = importlib.util.spec_from_file_location("_ref_" + filename[:-3], os.path.join(REFERENCE_ROOT, "tests", filename))
    mod = importlib.util.module_from_spec(spec)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(REFERENCE_ROOT)
    return mod
