"""Feed-forward Gaussian actor and value critic with the reference's interface
(rl/policies/actor.py:122-189 Gaussian_FF_Actor, rl/policies/critic.py:15-49 FF_V, rl/policies/base.py:5-22):
2 x 256 ReLU MLP on (state - obs_mean) / obs_std, fixed or learned per-action std, "normc" initialisation
(rows of N(0,1) weights scaled to unit norm, zero bias, output layer x 0.01).  The GEMMs go to cuBLAS — the
north-star leaves the small MLP to the library."""
from __future__ import annotations

import torch
import torch.nn as nn


def _normc_(linear: nn.Linear, gain: float = 1.0):
    with torch.no_grad():
        w = torch.randn_like(linear.weight)
        w *= gain / w.pow(2).sum(1, keepdim=True).sqrt()
        linear.weight.copy_(w)
        if linear.bias is not None:
            linear.bias.zero_()


class _MLP(nn.Module):
    def __init__(self, in_dim, layers, out_dim, out_gain):
        super().__init__()
        dims = [in_dim] + list(layers)
        self.hidden = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))
        self.out = nn.Linear(dims[-1], out_dim)
        for lin in self.hidden:
            _normc_(lin)
        _normc_(self.out, out_gain)

    def forward(self, x):
        for lin in self.hidden:
            x = torch.relu(lin(x))
        return self.out(x)


class _NormAttrs(nn.Module):
    """obs_mean / obs_std / stds are plain tensor attributes in the reference (moved by hand in rl/algos/ppo.py:136-147);
    here they follow .to()/.cuda()/.cpu() automatically."""

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        for name in ("stds", "obs_mean", "obs_std"):
            v = getattr(self, name, None)
            if torch.is_tensor(v) and not isinstance(v, nn.Parameter):
                setattr(self, name, fn(v))
        return self


class Gaussian_FF_Actor(_NormAttrs):
    def __init__(self, state_dim, action_dim, layers=(256, 256), init_std=0.2, learn_std=False, bounded=False):
        super().__init__()
        self.net = _MLP(state_dim, layers, action_dim, 0.01)
        self.learn_std = learn_std
        if learn_std:
            self.stds = nn.Parameter(init_std * torch.ones(action_dim))
        else:
            self.stds = init_std * torch.ones(action_dim)
        self.state_dim, self.action_dim, self.bounded = state_dim, action_dim, bounded
        self.obs_mean, self.obs_std = 0.0, 1.0

    def _get_dist_params(self, state):
        mean = self.net((state - self.obs_mean) / self.obs_std)
        if self.bounded:
            mean = torch.tanh(mean)
        return mean, self.stds

    def forward(self, state, deterministic=True):
        mu, sd = self._get_dist_params(state)
        return mu if deterministic else torch.distributions.Normal(mu, sd).sample()

    def distribution(self, inputs):
        mu, sd = self._get_dist_params(inputs)
        # validate_args would test the parameters with a host-synchronising `.all()` (not capturable in a CUDA graph);
        # mu is finite by construction of the update (checked by the trainer's loss statistics) and sd is a constant
        return torch.distributions.Normal(mu, sd, validate_args=False)


class FF_V(_NormAttrs):
    def __init__(self, state_dim, layers=(256, 256)):
        super().__init__()
        self.net = _MLP(state_dim, layers, 1, 1.0)
        self.obs_mean, self.obs_std = 0.0, 1.0

    stds = None

    def forward(self, state):
        return self.net((state - self.obs_mean) / self.obs_std)
