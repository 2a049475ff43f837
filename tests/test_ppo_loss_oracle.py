"""The closed-form loss tail (oracle/ppo_oracle.py: ppo_loss_and_grads, what csrc/ppo_kernels.cu: ppo_loss_kernel computes) against
torch autograd on the reference's own formulation of the losses (rl/algos/ppo.py:302-386, restated with torch.distributions) —
CPU only; the kernel itself is compared with both in tests/test_gpu_ppo.py."""
import numpy as np
import torch

from oracle.ppo_oracle import ppo_loss_and_grads


def _torch_reference(mu, old_mu, act, adv, ret, val, mirr, stds, clip, mc, ec):
    mu, val, mirr = (torch.tensor(x, dtype=torch.float64, requires_grad=True) for x in (mu, val, mirr))
    old_mu, act, adv, ret, stds = (torch.tensor(x, dtype=torch.float64) for x in (old_mu, act, adv, ret, stds))
    pdf, old_pdf = torch.distributions.Normal(mu, stds), torch.distributions.Normal(old_mu, stds)
    log_probs = pdf.log_prob(act).sum(-1, keepdim=True)
    old_log_probs = old_pdf.log_prob(act).sum(-1, keepdim=True)
    ratio = (log_probs - old_log_probs).exp()
    cpi_loss = ratio * adv
    clip_loss = ratio.clamp(1.0 - clip, 1.0 + clip) * adv
    actor_loss = -torch.min(cpi_loss, clip_loss).mean()
    clip_fraction = torch.mean((torch.abs(ratio - 1) > clip).double())
    critic_loss = torch.nn.functional.mse_loss(ret, val)
    entropy_penalty = -pdf.entropy().mean()
    mirror_loss = (mu - mirr).pow(2).mean()
    approx_kl = torch.mean((ratio - 1) - (log_probs - old_log_probs))
    total = actor_loss + mc * mirror_loss + ec * entropy_penalty + critic_loss
    total.backward()
    out = [actor_loss, entropy_penalty, critic_loss, approx_kl, mirror_loss, torch.zeros(()), clip_fraction, total]
    return np.array([float(x) for x in out]), mu.grad.numpy(), mirr.grad.numpy(), val.grad.numpy()


def test_closed_form_loss_and_gradients_equal_autograd_on_the_reference_formulation():
    rng = np.random.RandomState(0)
    B, A = 257, 12
    stds = np.full(A, 0.223)
    for scale in (0.02, 0.3):          # 0.3: many ratios leave the clip range on both sides, with advantages of both signs
        mu = rng.normal(size=(B, A)) * 0.2
        old_mu = mu + rng.normal(size=(B, A)) * scale * 0.223
        act = old_mu + rng.normal(size=(B, A)) * 0.223
        adv, ret, val = rng.normal(size=(B, 1)), rng.normal(size=(B, 1)), rng.normal(size=(B, 1))
        mirr = mu + rng.normal(size=(B, A)) * 0.05
        exp = _torch_reference(mu, old_mu, act, adv, ret, val, mirr, stds, 0.2, 0.4, 0.01)
        got = ppo_loss_and_grads(mu, old_mu, act, adv, ret, val, mirr, stds, 0.2, 0.4, 0.01)
        assert np.abs(got[0] - exp[0]).max() < 1e-12
        assert np.abs(got[1] - exp[1]).max() < 1e-14 and np.abs(got[2] - exp[2]).max() < 1e-14
        assert np.abs(got[3].reshape(-1) - exp[3].reshape(-1)).max() < 1e-14
        if scale == 0.3:
            assert 0.1 < got[0][6] < 0.9        # the clip fraction says the case exercises the clipped branches


def test_linear_param_grads_equal_autograd_of_nn_linear():
    """oracle.ppo_oracle.linear_param_grads (the checker of lhw_linear_wgrad) against torch's own backward of the reference's layer
    type, float64, for the four layer shapes of the actor / critic and a ragged one."""
    import torch
    from oracle.ppo_oracle import linear_param_grads
    torch.manual_seed(0)
    for M, N, K in ((64, 256, 37), (64, 256, 256), (64, 12, 256), (64, 1, 256), (7, 3, 5)):
        layer = torch.nn.Linear(K, N).double()
        x, gy = torch.randn(M, K, dtype=torch.float64), torch.randn(M, N, dtype=torch.float64)
        layer(x).backward(gy)
        gw, gb = linear_param_grads(gy.numpy(), x.numpy())
        assert np.abs(gw - layer.weight.grad.numpy()).max() < 1e-12 and np.abs(gb - layer.bias.grad.numpy()).max() < 1e-12
