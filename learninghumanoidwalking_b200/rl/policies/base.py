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


class _LinearWgrad(torch.autograd.Function):
    """y = x W^T + b with the library's forward and input gradient, and the parameter gradients of a long minibatch from
    csrc/wgrad_kernels.cu: gW = gy^T x and gb = sum(gy, 0) in one streaming pass + an ordered reduction (deterministic)
    instead of a split-K library GEMM, its epilogue launch and a separate column-sum (lhw_linear_wgrad)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = x.contiguous()
        ctx.save_for_backward(x, weight)
        return torch.nn.functional.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, gy):
        from ... import _lib
        x, weight = ctx.saved_tensors
        gy = gy.contiguous()
        gx = gy.mm(weight) if ctx.needs_input_grad[0] else None
        (M, N), K = gy.shape, x.shape[1]
        gw = torch.empty_like(weight, memory_format=torch.contiguous_format)
        gb = torch.empty(N, dtype=gy.dtype, device=gy.device)
        ws = torch.empty(max(1, _lib.lib().lhw_linear_wgrad_workspace_floats(M, N, K)), dtype=torch.float32, device=gy.device)
        _lib.ops().linear_wgrad(gy, x, gw, gb, ws)
        return gx, gw, gb


def _use_wgrad_kernel(x, layer):
    if not (x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and torch.is_grad_enabled()):
        return False
    if layer.bias is None or not (layer.weight.requires_grad and layer.bias.requires_grad) or layer.weight.dtype != torch.float32:
        return False
    import os
    if os.environ.get("LHW_WGRAD_KERNEL", "1") == "0" or torch.backends.cuda.matmul.allow_tf32:
        return False      # --tf32 hands every GEMM of the update to the tensor cores; this kernel is the fp32 FFMA path
    from ... import _lib
    return _lib.use_torch_ops()


def linear(layer, x):
    """`layer(x)` for an nn.Linear; on the training path (CUDA, grad enabled) the parameter gradients come from lhw_linear_wgrad."""
    if _use_wgrad_kernel(x, layer):
        return _LinearWgrad.apply(x, layer.weight, layer.bias)
    return layer(x)


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
