"""Players of the self-play engine: weight packing for the device kernels and the host-side network mirrors.

`QNet` / `QNetRNN` here have the architecture, parameter names (state_dict keys) and initialisation of the
reference modules (models/qnet.py:6-75, models/qnet_rnn.py:53-152), so reference checkpoints load unchanged
and weights trained here load into the reference.  They exist for random initialisation, checkpoint I/O and
the batched DQN update; ACTION SELECTION never runs through them — it runs in libpong_b200.so from the packed
blobs built by `pack_qnet` / `pack_qnetrnn`.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


# ----------------------------------------------------------------------------------------- host mirrors
class NoisyLinear(nn.Module):
    """Factorised-Gaussian NoisyNet layer                                   models/qnet.py:6-50"""

    def __init__(self, in_features, out_features, sigma_init=0.017):
        super().__init__()
        self.in_features, self.out_features, self.sigma_init = in_features, out_features, sigma_init
        r = 1.0 / math.sqrt(in_features)
        self.weight_mu = nn.Parameter(torch.empty(out_features, in_features).uniform_(-r, r))
        self.bias_mu = nn.Parameter(torch.empty(out_features).uniform_(-r, r))
        self.weight_sigma = nn.Parameter(torch.full((out_features, in_features), float(sigma_init)))
        self.bias_sigma = nn.Parameter(torch.full((out_features,), float(sigma_init)))
        self.register_buffer("weight_epsilon", torch.zeros(out_features, in_features))
        self.register_buffer("bias_epsilon", torch.zeros(out_features))
        self.reset_noise()

    @staticmethod
    def _signed_sqrt(n, device):
        g = torch.randn(n, device=device)
        return g.sign() * g.abs().sqrt()

    def reset_noise(self):
        dev = self.weight_mu.device
        e_in, e_out = self._signed_sqrt(self.in_features, dev), self._signed_sqrt(self.out_features, dev)
        self.weight_epsilon.copy_(torch.outer(e_out, e_in))
        self.bias_epsilon.copy_(e_out)

    def effective(self, noisy: bool):
        if noisy:
            return (self.weight_mu + self.weight_sigma * self.weight_epsilon,
                    self.bias_mu + self.bias_sigma * self.bias_epsilon)
        return self.weight_mu, self.bias_mu

    def forward(self, x):
        w, b = self.effective(self.training)
        return F.linear(x, w, b)


class _NoisyNet(nn.Module):
    def reset_noise(self):
        for m in self.modules():
            if isinstance(m, NoisyLinear):
                m.reset_noise()


class QNet(_NoisyNet):
    """7 -> 64 -> 64 -> dueling noisy heads (V 1, A 3)                        models/qnet.py:52-75"""

    def __init__(self, input_dim=7, output_dim=3):
        super().__init__()
        self.features = nn.Sequential(nn.Linear(input_dim, 64), nn.ReLU(), nn.Linear(64, 64), nn.ReLU())
        self.fc_V = NoisyLinear(64, 1)
        self.fc_A = NoisyLinear(64, output_dim)

    def forward(self, x):
        z = self.features(x)
        adv = self.fc_A(z)
        return self.fc_V(z) + (adv - adv.mean(dim=1, keepdim=True))


class QNetRNN(_NoisyNet):
    """7 -> 64 -> 128 -> LSTM(128) -> noisy 128 -> dueling noisy heads        models/qnet_rnn.py:53-152"""

    def __init__(self, input_dim=7, output_dim=3, feature_dim=128, lstm_hidden_dim=128, lstm_layers=1,
                 head_hidden_dim=128):
        super().__init__()
        self.input_dim, self.feature_dim = input_dim, feature_dim
        self.lstm_hidden_dim, self.lstm_layers, self.head_hidden_dim = lstm_hidden_dim, lstm_layers, head_hidden_dim
        self.features_extractor = nn.Sequential(nn.Linear(input_dim, feature_dim // 2), nn.ReLU(),
                                                nn.Linear(feature_dim // 2, feature_dim), nn.ReLU())
        self.lstm = nn.LSTM(input_size=feature_dim, hidden_size=lstm_hidden_dim, num_layers=lstm_layers, batch_first=True)
        if head_hidden_dim > 0:
            self.fc_shared_head = nn.Sequential(NoisyLinear(lstm_hidden_dim, head_hidden_dim), nn.ReLU())
            d = head_hidden_dim
        else:
            self.fc_shared_head, d = None, lstm_hidden_dim
        self.fc_V = NoisyLinear(d, 1)
        self.fc_A = NoisyLinear(d, output_dim)

    def init_hidden(self, batch_size, device):
        z = torch.zeros(self.lstm_layers, batch_size, self.lstm_hidden_dim, device=device)
        return z, z.clone()

    def forward(self, x_sequence, hidden_state_tuple):
        b, t, _ = x_sequence.shape
        f = self.features_extractor(x_sequence.reshape(b * t, self.input_dim)).reshape(b, t, self.feature_dim)
        y, hc = self.lstm(f, hidden_state_tuple)
        z = y[:, -1, :]
        if self.fc_shared_head is not None:
            z = self.fc_shared_head(z)
        adv = self.fc_A(z)
        return self.fc_V(z) + (adv - adv.mean(dim=1, keepdim=True)), hc


# ----------------------------------------------------------------------------------------- packing
def _sd(obj):
    sd = obj.state_dict() if isinstance(obj, nn.Module) else obj
    # tensors stay on their device (packing a CUDA model is a handful of device kernels, no host round trip)
    return {k: (v.detach().to(torch.float32) if torch.is_tensor(v) else torch.as_tensor(np.asarray(v), dtype=torch.float32))
            for k, v in sd.items()}


def _noisy(sd, prefix, noisy):
    w, b = sd[prefix + ".weight_mu"], sd[prefix + ".bias_mu"]
    if noisy:   # train mode: mu + sigma * epsilon                              models/qnet.py:44-46
        w = w + sd[prefix + ".weight_sigma"] * sd[prefix + ".weight_epsilon"]
        b = b + sd[prefix + ".bias_sigma"] * sd[prefix + ".bias_epsilon"]
    return w, b


def pack_qnet(model_or_state_dict, noisy: bool = False) -> torch.Tensor:
    """Effective fp32 weights of a QNet -> the k-major blob of include/pong_b200.h (PP_QNET_*), on the model's device.
    noisy=False is eval mode (mu); noisy=True is the train-mode forward the reference's training script plays
    with (scripts/train_iterative.py never calls .eval() on modelA / modelB)."""
    sd = _sd(model_or_state_dict)
    wv, bv = _noisy(sd, "fc_V", noisy)
    wa, ba = _noisy(sd, "fc_A", noisy)
    w1, w2 = sd["features.0.weight"], sd["features.2.weight"]
    if tuple(w1.shape) != (64, 7) or tuple(w2.shape) != (64, 64) or tuple(wa.shape) != (3, 64):
        raise ValueError("the device QNet is the reference architecture 7-64-64-(1,3)")
    blob = torch.cat([w1.t().reshape(-1), sd["features.0.bias"], w2.t().reshape(-1), sd["features.2.bias"],
                      torch.cat([wv, wa], 0).t().reshape(-1), torch.cat([bv, ba])]).contiguous()
    assert blob.numel() == _lib.QNET_BLOB_FLOATS
    return blob


def pack_qnetrnn(model_or_state_dict, noisy: bool = False) -> torch.Tensor:
    """QNetRNN (default dims 7-64-128 / LSTM 128 / head 128) -> the PP_RNN_* blob, on the model's device.  The gate matrix
    is [W_ih^T ; W_hh^T] (k-major, 256 rows) with columns ordered unit*4 + gate (gates i, f, g, o)."""
    sd = _sd(model_or_state_dict)
    wf1, wf2 = sd["features_extractor.0.weight"], sd["features_extractor.2.weight"]
    wih, whh = sd["lstm.weight_ih_l0"], sd["lstm.weight_hh_l0"]
    if tuple(wf1.shape) != (64, 7) or tuple(wf2.shape) != (128, 64) or tuple(wih.shape) != (512, 128) or \
            tuple(whh.shape) != (512, 128) or "lstm.weight_ih_l1" in sd:
        raise ValueError("the device QNetRNN is the reference default 7-64-128 / 1-layer LSTM 128 / head 128")
    ws, bs = _noisy(sd, "fc_shared_head.0", noisy)
    wv, bv = _noisy(sd, "fc_V", noisy)
    wa, ba = _noisy(sd, "fc_A", noisy)
    gates = torch.cat([wih, whh], 1)                                   # [512 = gate*128+unit, 256]
    gates = gates.reshape(4, 128, 256).permute(2, 1, 0).reshape(256, 512)      # [k][unit*4+gate]
    bg = (sd["lstm.bias_ih_l0"] + sd["lstm.bias_hh_l0"]).reshape(4, 128).t().reshape(-1)
    blob = torch.cat([wf1.t().reshape(-1), sd["features_extractor.0.bias"], wf2.t().reshape(-1),
                      sd["features_extractor.2.bias"], gates.reshape(-1), bg, ws.t().reshape(-1), bs,
                      torch.cat([wv, wa], 0).t().reshape(-1), torch.cat([bv, ba])]).contiguous()
    assert blob.numel() == _lib.RNN_BLOB_FLOATS
    return blob


def _tile_b(w_kn: torch.Tensor, part: str) -> torch.Tensor:
    """[K, N] fp32 (k-major) -> fp16 B-operand tile [K/8][N][8], hi = fp16(w) or lo = fp16(w - hi)."""
    hi = w_kn.to(torch.float16)
    t = hi if part == "hi" else (w_kn - hi.to(torch.float32)).to(torch.float16)
    k, n = t.shape
    return t.reshape(k // 8, 8, n).permute(0, 2, 1).contiguous().reshape(-1)


def _bias_tile(b: torch.Tensor, n: int) -> torch.Tensor:
    """[2][n][8] tile: bias hi in row k = 7, lo in row k = 15 (they meet the ones columns of the obs tile)."""
    t = torch.zeros(16, n, dtype=torch.float32, device=b.device)
    hi = b.to(torch.float16).to(torch.float32)
    t[7, :b.numel()] = hi
    t[15, :b.numel()] = b - hi
    return _tile_b(t, "hi")          # the two rows are already exact fp16 values


def pack_qnetrnn_tc(model_or_state_dict, noisy: bool = False) -> torch.Tensor:
    """QNetRNN -> the fp16 stage image PP_RNNTC_* of the tensor-core path (a uint8 tensor of PP_RNNTC_BLOB_BYTES)."""
    sd = _sd(model_or_state_dict)
    wf1, bf1 = sd["features_extractor.0.weight"], sd["features_extractor.0.bias"]          # [64, 7]
    wf2, bf2 = sd["features_extractor.2.weight"], sd["features_extractor.2.bias"]          # [128, 64]
    wih, whh = sd["lstm.weight_ih_l0"], sd["lstm.weight_hh_l0"]                            # [512, 128] each
    if tuple(wf1.shape) != (64, 7) or tuple(wf2.shape) != (128, 64) or tuple(wih.shape) != (512, 128) or \
            tuple(whh.shape) != (512, 128) or "lstm.weight_ih_l1" in sd:
        raise ValueError("the device QNetRNN is the reference default 7-64-128 / 1-layer LSTM 128 / head 128")
    ws, bs = _noisy(sd, "fc_shared_head.0", noisy)
    wv, bv = _noisy(sd, "fc_V", noisy)
    wa, ba = _noisy(sd, "fc_A", noisy)
    dev = wf1.device
    z = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=dev)
    # S0: layer 1 with K = 16: rows 0..6 W1^T, 7 b_hi | rows 8..14 W1^T (meets obs_lo), 15 b_lo ; lo tile: W1_lo in rows 0..6
    w1 = z(16, 64)
    w1[0:7] = wf1.t(); w1[8:15] = wf1.t()
    bh = bf1.to(torch.float16).to(torch.float32)
    w1h = _tile_b(w1, "hi").clone().reshape(2, 64, 8)
    w1h[0, :, 7] = bh.to(torch.float16); w1h[1, :, 7] = (bf1 - bh).to(torch.float16)
    w1l_src = z(16, 64); w1l_src[0:7] = wf1.t()
    w1l = _tile_b(w1l_src, "lo")
    parts = [w1h.reshape(-1), w1l]
    parts += [_tile_b(wf2.t().contiguous(), "hi"), _bias_tile(bf2, 128), _tile_b(wf2.t().contiguous(), "lo")]          # S1, S2
    wcat = torch.cat([wih, whh], 1).t().contiguous()                                        # [256 k, 512 = gate*128 + unit]
    bg = sd["lstm.bias_ih_l0"] + sd["lstm.bias_hh_l0"]
    for q in range(4):
        cols = torch.cat([torch.arange(g * 128 + 32 * q, g * 128 + 32 * q + 32) for g in range(4)]).to(dev)
        wq = wcat[:, cols]                                                                   # [256, 128], column = gate*32 + unit%32
        for c in range(4):
            parts.append(_tile_b(wq[64 * c:64 * c + 64].contiguous(), "hi"))
            if c == 0:
                parts.append(_bias_tile(bg[cols], 128))
        for c in range(4):
            parts.append(_tile_b(wq[64 * c:64 * c + 64].contiguous(), "lo"))
    wst = ws.t().contiguous()                                                                # [128 k, 128]
    parts += [_tile_b(wst[:64].contiguous(), "hi"), _bias_tile(bs, 128), _tile_b(wst[64:].contiguous(), "hi"),
              _tile_b(wst[:64].contiguous(), "lo"), _tile_b(wst[64:].contiguous(), "lo")]
    wh = z(128, 16); wh[:, 0:1] = wv.t(); wh[:, 1:4] = wa.t()
    parts += [_tile_b(wh, "hi"), _tile_b(wh, "lo"), _bias_tile(torch.cat([bv, ba]), 16)]
    blob = torch.cat(parts).contiguous().view(torch.uint8)
    assert blob.numel() == _lib.RNNTC_BLOB_BYTES, blob.numel()
    return blob


def eps_threshold(eps: float) -> int:
    """explore iff (uint64) philox.x < floor(eps * 2^32)"""
    return int(min(max(float(eps), 0.0), 1.0) * 4294967296.0)


class Policy:
    """One player: kind + packed weights on the device (+ per-env (h, c) for QNetRNN, unit-major [128, n])."""

    def __init__(self, kind, weights=None, eps=0.0, precision="f32", tol=0.02, device="cuda", num_envs=None):
        self.kind, self.eps, self.tol = kind, float(eps), float(tol)
        self.precision = {"f32": _lib.PREC_F32, "f16": _lib.PREC_F16}[precision]
        self.device = torch.device(device)
        self.weights = None if weights is None else weights.to(self.device, torch.float32).contiguous()
        self.h = self.c = None
        if kind == _lib.POLICY_QNETRNN:
            if num_envs is None:
                raise ValueError("QNetRNN players carry per-env (h, c): pass num_envs")
            # fp32 path: unit-major [128, n] (coalesced along the env index); tensor-core path: blocked by warp,
            # [ceil(n / 32)][32 unit groups][32 envs][4] (a warp's accesses are contiguous 512-byte runs).  `hidden()`
            # returns either as [n, 128].
            self.num_envs = int(num_envs)
            shape = ((num_envs + 31) // 32, 32, 32, 4) if self.precision == _lib.PREC_F16 else (128, num_envs)
            self.h = torch.zeros(*shape, dtype=torch.float32, device=self.device)
            self.c = torch.zeros(*shape, dtype=torch.float32, device=self.device)

    @classmethod
    def qnet(cls, model_or_state_dict, noisy=False, **kw):
        return cls(_lib.POLICY_QNET, pack_qnet(model_or_state_dict, noisy), **kw)

    @classmethod
    def qnetrnn(cls, model_or_state_dict, num_envs, noisy=False, **kw):
        if kw.get("precision", "f32") == "f16":          # tensor-core path: the fp16 stage image instead of the fp32 blob
            pol = cls(_lib.POLICY_QNETRNN, None, num_envs=num_envs, **kw)
            pol.weights = pack_qnetrnn_tc(model_or_state_dict, noisy).to(pol.device).contiguous()
            return pol
        return cls(_lib.POLICY_QNETRNN, pack_qnetrnn(model_or_state_dict, noisy), num_envs=num_envs, **kw)

    @classmethod
    def follower(cls, tol=0.02, **kw):
        """HardcodedBallFollower                                                  tests/arena.py:211-217"""
        return cls(_lib.POLICY_FOLLOWER, tol=tol, **kw)

    @classmethod
    def random(cls, **kw):
        return cls(_lib.POLICY_RANDOM, **kw)

    def hidden(self):
        """(h, c) as [num_envs, 128] views, whatever the storage layout of the precision path."""
        if self.h is None:
            return None, None
        if self.precision == _lib.PREC_F16:
            unblock = lambda t: t.permute(0, 2, 1, 3).reshape(-1, 128)[:self.num_envs]
            return unblock(self.h), unblock(self.c)
        return self.h.t(), self.c.t()

    def set_weights(self, blob):
        self.weights.copy_(blob.to(self.weights.device, torch.float32), non_blocking=True)

    def struct(self) -> _lib.PPPolicy:
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        return _lib.PPPolicy(self.kind, self.precision, eps_threshold(self.eps), self.tol,
                             p(self.weights), p(self.h), p(self.c))
