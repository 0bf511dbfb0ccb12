"""C-ABI checks that need no GPU: the library builds/loads, exports every symbol include/pong_b200.h declares,
the ctypes mirrors have the C struct layouts, and compute entries refuse to run without a device."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import pingpong_selfplay_ai_b200 as pp
from pingpong_selfplay_ai_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pong_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.EXPORTS) == names                      # the binding covers the whole header, nothing else
    assert lib.pp_version() == _lib.PP_ABI_VERSION
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.lib_path()], capture_output=True, text=True).stdout
    assert all(re.search(rf"\bT {n}\b", out) for n in names)


def test_ctypes_structs_match_c_layout(tmp_path):
    """Compile a C program against the header with gcc and compare sizeof / offsetof with the ctypes mirrors."""
    structs = {"PPParams": _lib.PPParams, "PPEnvState": _lib.PPEnvState, "PPServeSource": _lib.PPServeSource,
               "PPPolicy": _lib.PPPolicy, "PPRolloutOut": _lib.PPRolloutOut, "PPReplayRing": _lib.PPReplayRing,
               "PPNoisyLayer": _lib.PPNoisyLayer, "PPAdamParam": _lib.PPAdamParam,
               "PPQNetRNNParams": _lib.PPQNetRNNParams, "PPQNetRNNGrads": _lib.PPQNetRNNGrads,
               "PPPeerBlocks": _lib.PPPeerBlocks}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for name, st in structs.items():
        lines.append(f'printf("{name} %zu\\n", sizeof({name}));')
        for f in st._fields_:
            lines.append(f'printf("{name}.{f[0]} %zu\\n", offsetof({name}, {f[0]}));')
    lines.append('printf("QNET %d RNN %d ABI %d\\n", PP_QNET_BLOB_FLOATS, PP_RNN_BLOB_FLOATS, PP_ABI_VERSION);')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    got = dict(l.rsplit(" ", 1) for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()[:-1])
    for name, st in structs.items():
        assert int(got[name]) == C.sizeof(st), name
        for f in st._fields_:
            assert int(got[f"{name}.{f[0]}"]) == getattr(st, f[0]).offset, (name, f[0])
    last = subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines()[-1].split()
    assert (int(last[1]), int(last[3]), int(last[5])) == (_lib.QNET_BLOB_FLOATS, _lib.RNN_BLOB_FLOATS, _lib.PP_ABI_VERSION)


def test_argument_validation_needs_no_device():
    lib = _lib.load()
    assert lib.pp_env_step(0, 4, None, None, None, None, None, None, None, None, None, None) == -5      # PP_E_PARAM
    assert b"pp_env_step" in lib.pp_last_error()
    assert lib.pp_qnet_act(-1, None, None, 0, 0, 0, 1, None, None, None) == -2                          # PP_E_SIZE


def test_product_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(pp.PongB200Error):
        pp.VecPongEnv2P(8)
    with pytest.raises(pp.PongB200Error):
        pp.PongEnv2P()
    # the host-buffer entry reaches the CUDA runtime and reports its error code (> 0), never computes on the CPU
    cfg = pp.resolve_env_config({})
    pool = tuple(np.zeros((1, 4)) for _ in range(3))
    w = np.zeros(_lib.QNET_BLOB_FLOATS, np.float32)
    with pytest.raises(pp.PongB200Error, match="CUDA error"):
        pp.host_selfplay_eval({}, 4, 1, pool, w, w)
    del cfg


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: importing the product must not pull it in."""
    code = "import sys, pingpong_selfplay_ai_b200 as p; p._lib.load(); print(any(m.startswith('oracle') for m in sys.modules))"
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, check=True).stdout
    assert out.strip() == "False"
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pingpong_selfplay_ai_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "libpong_oracle" not in text, f


def test_import_compatible_shims_resolve_to_the_engine():
    """`envs/` and `models/` at the repository root carry the reference's import names (scripts/train_iterative.py:18-19,
    tests/arena.py:42-44) and re-export the engine's classes.  A fresh interpreter: other tests of this session may
    hold the UNMODIFIED reference under the same module names."""
    code = ("import envs.my_pong_env_2p as e, envs.physics as ph, models.qnet as q, models.qnet_rnn as r, "
            "pingpong_selfplay_ai_b200 as pp; "
            "assert e.PongEnv2P is pp.PongEnv2P and q.QNet is pp.QNet and r.QNetRNN is pp.QNetRNN; "
            "assert q.NoisyLinear is pp.NoisyLinear and callable(ph.collide_sphere_with_moving_plane); "
            "assert e.__file__.startswith(%r); print('ok')" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr
