"""Net base of the actor / critic modules (rl/policies/base.py:5-22): "normc" initialisation — every Linear gets N(0,1)
weights with rows scaled to unit norm and a zero bias, the output layer is then multiplied by 0.01 where asked."""
from __future__ import annotations

import torch
import torch.nn as nn


def normc_fn(m):
    if isinstance(m, nn.Linear):
        with torch.no_grad():
            w = torch.randn_like(m.weight)
            w *= 1 / w.pow(2).sum(1, keepdim=True).sqrt()
            m.weight.copy_(w)
            if m.bias is not None:
                m.bias.zero_()


class Net(nn.Module):
    """obs_mean / obs_std / stds are plain tensor attributes in the reference (moved by hand in rl/algos/ppo.py:136-147);
    here they follow .to() / .cuda() / .cpu() automatically."""

    def init_parameters(self, output_layer=None):
        if getattr(self, "normc_init", True):
            self.apply(normc_fn)
            if output_layer is not None:
                with torch.no_grad():
                    output_layer.weight.mul_(0.01)

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        for name in ("stds", "obs_mean", "obs_std"):
            v = self.__dict__.get(name, None)
            if torch.is_tensor(v):
                self.__dict__[name] = fn(v)
        return self
