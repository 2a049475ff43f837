"""lhw_linear_wgrad (csrc/wgrad_kernels.cu): the parameter gradients of a Linear layer, gW = gy^T x and gb = sum(gy, 0), against
float64 matrix products of the same operands; every tile configuration, ragged sizes, unaligned operands, empty batch, run-to-run
determinism, and the autograd function the actor / critic use against torch's own Linear backward (rl/algos/ppo.py:389-392)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _L():
    from learninghumanoidwalking_b200 import _lib
    return _lib


def _run(gy, x, bias=True):
    L = _L()
    M, N = gy.shape
    K = x.shape[1]
    gw = torch.full((N, K), float("nan"), device="cuda")
    gb = torch.full((N,), float("nan"), device="cuda") if bias else None
    ws = torch.empty(max(1, L.lib().lhw_linear_wgrad_workspace_floats(M, N, K)), device="cuda")
    L.ops().linear_wgrad(gy, x, gw, gb, ws)
    return gw, gb


@pytest.mark.parametrize("M,N,K", [(43690, 256, 256), (21845, 256, 37), (43690, 12, 256), (21845, 1, 256), (64, 256, 39), (1000, 64, 64),
                                   (777, 17, 130), (5, 3, 7), (1, 1, 1), (9, 130, 65), (4097, 16, 64), (300, 65, 300)])
def test_wgrad_matches_float64_products(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + 7 * N + 13 * K)
    gy, x = torch.randn(M, N, device="cuda", generator=g), torch.randn(M, K, device="cuda", generator=g)
    gw, gb = _run(gy, x)
    ref = gy.double().t().mm(x.double())
    scale = gy.abs().double().t().mm(x.abs().double()).max().item()
    assert torch.isfinite(gw).all() and (gw.double() - ref).abs().max().item() < 2e-6 * scale
    if M * N * K <= 1 << 24:      # the numpy oracle (pinned to torch autograd on the CPU) where it is quick
        from oracle.ppo_oracle import linear_param_grads
        ow, ob = linear_param_grads(gy.cpu().numpy(), x.cpu().numpy())
        assert np.abs(gw.double().cpu().numpy() - ow).max() < 2e-6 * scale
    refb = gy.double().sum(0)
    assert torch.isfinite(gb).all() and (gb.double() - refb).abs().max().item() < 2e-6 * gy.abs().double().sum(0).max().item()
    gw2, gb2 = _run(gy, x)
    assert torch.equal(gw, gw2) and torch.equal(gb, gb2)            # fixed summation order
    gw3, none = _run(gy, x, bias=False)
    assert none is None and torch.equal(gw, gw3)


def test_wgrad_unaligned_operands_and_empty_batch():
    g = torch.Generator(device="cuda").manual_seed(3)
    M, N, K = 999, 256, 256
    gbuf, xbuf = torch.randn(M * N + 1, device="cuda", generator=g), torch.randn(M * K + 3, device="cuda", generator=g)
    gy, x = gbuf[1:].view(M, N), xbuf[3:].view(M, K)        # 4 / 12 bytes off a 16-byte boundary: the scalar loader
    assert gy.data_ptr() % 16 != 0 and x.data_ptr() % 16 != 0
    gw, gb = _run(gy, x)
    ref = gy.double().t().mm(x.double())
    assert (gw.double() - ref).abs().max().item() < 2e-6 * gy.abs().double().t().mm(x.abs().double()).max().item()
    gw0, gb0 = _run(torch.empty(0, 12, device="cuda"), torch.empty(0, 256, device="cuda"))
    assert float(gw0.abs().sum()) == 0.0 and float(gb0.abs().sum()) == 0.0
    with pytest.raises(RuntimeError):
        _run(torch.randn(8, 4, device="cuda"), torch.randn(9, 4, device="cuda"))
    with pytest.raises(RuntimeError):
        _L().ops().linear_wgrad(torch.randn(8, 4), torch.randn(8, 4), torch.empty(4, 4), None, torch.empty(64))      # CPU tensors


def test_actor_and_critic_gradients_match_torchs_linear_backward(monkeypatch):
    from learninghumanoidwalking_b200.rl import FF_V, Gaussian_FF_Actor
    torch.manual_seed(0)
    a, c = Gaussian_FF_Actor(37, 12).cuda(), FF_V(37).cuda()
    x = torch.randn(2048, 37, device="cuda")
    grads = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("LHW_WGRAD_KERNEL", flag)
        for p in list(a.parameters()) + list(c.parameters()):
            p.grad = None
        (a(x).pow(2).mean() * 50 + (c(x) - 1).pow(2).mean()).backward()
        grads[flag] = [p.grad.clone() for p in list(a.parameters()) + list(c.parameters())]
    for k, t in zip(grads["1"], grads["0"]):
        assert (k - t).abs().max().item() < 1e-5 * max(1e-3, t.abs().max().item())
    # the training path really goes through the kernel: with grad enabled the forward is the custom autograd function
    assert type(a(x).grad_fn).__name__ != type(torch.nn.functional.linear(x, a.actor_layers[0].weight).grad_fn).__name__
    with torch.no_grad():
        assert a(x).grad_fn is None
