"""torch restatement of the reference policy modules, for the CPU baseline loop only.

TEST INFRASTRUCTURE — used by `bench.py --impl reference` / `cpu_baseline` (the reference's own
per-step `model(torch.tensor(obs).unsqueeze(0)).argmax(1).item()` loop, scripts/train_iterative.py:
171-181) and by tests.  /root/reference does not exist on the GPU box, so the baseline loop needs
modules with the same architecture, parameter names and initialisation:
  NoisyLinear   models/qnet.py:6-50     (factorised Gaussian noise, sigma_init 0.017)
  QNet          models/qnet.py:52-75    features 7-64-64, dueling noisy heads
  QNetRNN       models/qnet_rnn.py:53-152   7-64-128 -> LSTM(128) -> noisy 128 -> dueling
`tests/test_oracle_vs_reference.py` checks state_dict interchangeability and identical outputs.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _f(n):
    g = torch.randn(n)
    return g.sign() * g.abs().sqrt()


class NoisyDense(nn.Module):
    def __init__(self, n_in, n_out, sigma0=0.017):
        super().__init__()
        bound = 1.0 / math.sqrt(n_in)
        self.weight_mu = nn.Parameter(torch.empty(n_out, n_in).uniform_(-bound, bound))
        self.bias_mu = nn.Parameter(torch.empty(n_out).uniform_(-bound, bound))
        self.weight_sigma = nn.Parameter(torch.full((n_out, n_in), sigma0))
        self.bias_sigma = nn.Parameter(torch.full((n_out,), sigma0))
        self.register_buffer("weight_epsilon", torch.zeros(n_out, n_in))
        self.register_buffer("bias_epsilon", torch.zeros(n_out))
        self.reset_noise()

    def reset_noise(self):
        e_in, e_out = _f(self.weight_mu.shape[1]), _f(self.weight_mu.shape[0])
        self.weight_epsilon.copy_(torch.outer(e_out, e_in))
        self.bias_epsilon.copy_(e_out)

    def forward(self, x):
        if self.training:
            return F.linear(x, self.weight_mu + self.weight_sigma * self.weight_epsilon,
                            self.bias_mu + self.bias_sigma * self.bias_epsilon)
        return F.linear(x, self.weight_mu, self.bias_mu)


class _Resettable(nn.Module):
    def reset_noise(self):
        for m in self.modules():
            if isinstance(m, NoisyDense):
                m.reset_noise()


class QNetPort(_Resettable):
    def __init__(self, input_dim=7, output_dim=3):
        super().__init__()
        self.features = nn.Sequential(nn.Linear(input_dim, 64), nn.ReLU(), nn.Linear(64, 64), nn.ReLU())
        self.fc_V = NoisyDense(64, 1)
        self.fc_A = NoisyDense(64, output_dim)

    def forward(self, x):
        z = self.features(x)
        adv = self.fc_A(z)
        return self.fc_V(z) + (adv - adv.mean(dim=1, keepdim=True))


class QNetRNNPort(_Resettable):
    def __init__(self, input_dim=7, output_dim=3, feature_dim=128, lstm_hidden_dim=128, lstm_layers=1,
                 head_hidden_dim=128):
        super().__init__()
        self.input_dim, self.feature_dim = input_dim, feature_dim
        self.lstm_hidden_dim, self.lstm_layers = lstm_hidden_dim, lstm_layers
        self.features_extractor = nn.Sequential(nn.Linear(input_dim, feature_dim // 2), nn.ReLU(),
                                                nn.Linear(feature_dim // 2, feature_dim), nn.ReLU())
        self.lstm = nn.LSTM(feature_dim, lstm_hidden_dim, lstm_layers, batch_first=True)
        self.fc_shared_head = nn.Sequential(NoisyDense(lstm_hidden_dim, head_hidden_dim), nn.ReLU())
        self.fc_V = NoisyDense(head_hidden_dim, 1)
        self.fc_A = NoisyDense(head_hidden_dim, output_dim)

    def init_hidden(self, batch, device):
        z = torch.zeros(self.lstm_layers, batch, self.lstm_hidden_dim, device=device)
        return z, z.clone()

    def forward(self, x_seq, hc):
        b, t, _ = x_seq.shape
        feats = self.features_extractor(x_seq.reshape(b * t, self.input_dim)).reshape(b, t, self.feature_dim)
        y, hc = self.lstm(feats, hc)
        z = self.fc_shared_head(y[:, -1, :])
        adv = self.fc_A(z)
        return self.fc_V(z) + (adv - adv.mean(dim=1, keepdim=True)), hc
