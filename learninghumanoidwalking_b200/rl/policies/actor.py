"""Gaussian feed-forward actor with the reference's interface AND attribute names (rl/policies/actor.py:122-189):
`actor_layers` (ModuleList of Linear), `means`, `stds`, `obs_mean`, `obs_std`, `nonlinearity`, `bounded`, `learn_std` — a
module pickled here therefore has the state-dict keys and the instance dictionary the reference's class expects.
2 x 256 ReLU MLP on (state - obs_mean) / obs_std, fixed or learned per-action std.  The GEMMs go to cuBLAS — the north-star
leaves the small MLP to the library."""
from __future__ import annotations

import torch
import torch.nn as nn

from .base import Net, linear


class Actor(Net):
    def forward(self, state, deterministic=True):
        raise NotImplementedError


class Gaussian_FF_Actor(Actor):
    def __init__(self, state_dim, action_dim, layers=(256, 256), nonlinearity=torch.nn.functional.relu, init_std=0.2,
                 learn_std=False, bounded=False, normc_init=True):
        super().__init__()
        dims = [state_dim] + list(layers)
        self.actor_layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))
        self.means = nn.Linear(dims[-1], action_dim)
        self.learn_std = learn_std
        if learn_std:
            self.stds = nn.Parameter(init_std * torch.ones(action_dim))
        else:
            self.stds = init_std * torch.ones(action_dim)
        self.action_dim, self.state_dim = action_dim, state_dim
        self.nonlinearity = nonlinearity
        self.obs_std, self.obs_mean = 1.0, 0.0
        self.bounded = bounded
        self.normc_init = normc_init
        self.init_parameters(self.means)

    def _get_dist_params(self, state):
        x = (state - self.obs_mean) / self.obs_std
        for layer in self.actor_layers:
            x = self.nonlinearity(linear(layer, x))
        mean = linear(self.means, x)
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
