"""FusedClipAdam — torch.nn.utils.clip_grad_norm_(params, max_norm) + torch.optim.Adam.step
(rl/algos/ppo.py:393-396) as two CUDA launches over ONE flat parameter buffer (lhw_grad_sumsq + lhw_clip_adam)
instead of ~6 kernels per parameter tensor.  A torch.optim.Optimizer subclass, because the reference's tests
create the optimisers by hand and assign them to ppo.actor_optimizer / critic_optimizer
(tests/test_training.py:140-141).
"""
from __future__ import annotations

import torch

from .. import _lib


def flatten_module_(module: torch.nn.Module):
    """Re-home all parameters (and their grads) of `module` as views into two flat float32 buffers."""
    params = [p for p in module.parameters()]
    n = sum(p.numel() for p in params)
    dev = params[0].device
    flat = torch.zeros(n, dtype=torch.float32, device=dev)
    grad = torch.zeros(n, dtype=torch.float32, device=dev)
    off = 0
    for p in params:
        k = p.numel()
        flat[off:off + k].copy_(p.data.reshape(-1))
        p.data = flat[off:off + k].view_as(p)
        p.grad = grad[off:off + k].view_as(p)
        off += k
    return flat, grad


def flatten_modules_(modules, grad_buffer=None):
    """Several modules in ONE pair of flat buffers (one all-reduce covers all of them).  `grad_buffer`: use this
    (e.g. peer-visible IPC) tensor for the gradients.  Returns flat, grad, [(start, end) per module]."""
    params = [[p for p in m.parameters()] for m in modules]
    n = sum(p.numel() for ps in params for p in ps)
    dev = params[0][0].device
    flat = torch.zeros(n, dtype=torch.float32, device=dev)
    grad = torch.zeros(n, dtype=torch.float32, device=dev) if grad_buffer is None else grad_buffer
    assert grad.numel() == n and grad.dtype == torch.float32
    grad.zero_()
    off, segs = 0, []
    for ps in params:
        start = off
        for p in ps:
            k = p.numel()
            flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = flat[off:off + k].view_as(p)
            p.grad = grad[off:off + k].view_as(p)
            off += k
        segs.append((start, off))
    return flat, grad, segs


class FusedClipAdam(torch.optim.Optimizer):
    def __init__(self, module: torch.nn.Module, lr=1e-4, eps=1e-5, betas=(0.9, 0.999), max_norm=0.05,
                 process_group=None, views=None):
        params = list(module.parameters())
        if not params or not params[0].is_cuda:
            raise _lib.LhwError("FusedClipAdam needs CUDA parameters (no CPU fallback)")
        super().__init__(params, dict(lr=lr, eps=eps, betas=betas, max_norm=max_norm))
        self.flat, self.grad = views[:2] if views is not None else flatten_module_(module)
        if views is not None and len(views) == 4:
            self.exp_avg, self.exp_avg_sq = views[2], views[3]
        else:
            self.exp_avg = torch.zeros_like(self.flat)
            self.exp_avg_sq = torch.zeros_like(self.flat)
        self.norm = torch.zeros(1, dtype=torch.float32, device=self.flat.device)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=self.flat.device)   # completed Adam steps, device side
        self.step_count = 0
        self.process_group = process_group
        self.world = 1

    def zero_grad(self, set_to_none: bool = False):
        self.grad.zero_()

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        self.step_count += 1
        O = _lib.ops()
        scale = 1.0 / self.world
        O.grad_sumsq(self.grad, self.norm, scale)
        # the step number lives on the device (bias corrections computed in the kernel), so that an eager step and a step
        # replayed from a CUDA graph (PPO._update_graphed) are the same launches with the same arguments
        O.clip_adam_dev(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self.norm, self.step_dev, g["lr"], g["betas"][0],
                        g["betas"][1], g["eps"], g["max_norm"], scale)

    def total_norm(self) -> torch.Tensor:
        return self.norm.sqrt()
