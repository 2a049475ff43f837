"""Feed-forward value function with the reference's interface and attribute names (rl/policies/critic.py:15-49):
`critic_layers`, `network_out`, `obs_mean`, `obs_std`, `nonlinearity`."""
from __future__ import annotations

import torch
import torch.nn as nn

from .base import Net, linear


class Critic(Net):
    def forward(self, state):
        raise NotImplementedError


class FF_V(Critic):
    stds = None

    def __init__(self, state_dim, layers=(256, 256), nonlinearity=torch.nn.functional.relu, normc_init=True, obs_std=None,
                 obs_mean=None):
        super().__init__()
        dims = [state_dim] + list(layers)
        self.critic_layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))
        self.network_out = nn.Linear(dims[-1], 1)
        self.nonlinearity = nonlinearity
        self.obs_std = 1.0 if obs_std is None else obs_std
        self.obs_mean = 0.0 if obs_mean is None else obs_mean
        self.normc_init = normc_init
        self.init_parameters()

    def forward(self, state):
        x = (state - self.obs_mean) / self.obs_std
        for layer in self.critic_layers:
            x = self.nonlinearity(linear(layer, x))
        return linear(self.network_out, x)
