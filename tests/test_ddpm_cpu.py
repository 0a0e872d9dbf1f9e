"""unlearn_saliency_b200/diffusion (own U-Net restatement, eps-loss, beta schedule) pinned against outputs of the
UNMODIFIED reference DDPM code (tests/golden/ddpm_tiny.npz from tests/golden/make_golden_ddpm.py).  CPU only."""
import os

import numpy as np
import torch

from tests.golden.make_golden_ddpm import inputs, sample_idx, synth_weights, tiny_config
from unlearn_saliency_b200.diffusion.runner import eps_loss, get_beta_schedule, q_sample
from oracle.unet import ConditionalUNet, cifar10_config

G = os.path.join(os.path.dirname(__file__), "golden", "ddpm_tiny.npz")


def test_parameter_names_and_order_equal_reference():
    z = np.load(G)
    full = ConditionalUNet(cifar10_config())
    assert [n for n, _ in full.named_parameters()] == list(z["keys_full"])  # 334 names, null_classes_emb first
    assert sum(p.numel() for p in full.parameters()) == int(z["numel_full"]) == 38632323
    tiny = ConditionalUNet(tiny_config())
    assert [n for n, _ in tiny.named_parameters()] == list(z["keys_tiny"])


def test_beta_schedule():
    z = np.load(G)
    b = get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)
    np.testing.assert_allclose(b.astype(np.float32), z["betas"], rtol=1e-6)


def test_unet_forward_loss_backward_equal_reference():
    z = np.load(G)
    model = ConditionalUNet(tiny_config())
    model.load_state_dict(synth_weights(model))  # same formula, same key order -> same weights as the reference model
    model.eval()
    x0, e, t, c = inputs()
    betas = torch.from_numpy(z["betas"])
    xt = q_sample(x0, t, e, betas)
    np.testing.assert_allclose(model(xt, t.float(), c, cond_scale=2.0, mode="test").detach().numpy(), z["eps_test"],
                               rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(model(xt, t.float(), c, mode="train", cond_drop_prob=0.0).detach().numpy(),
                               z["eps_train"], rtol=1e-4, atol=1e-4)
    model.zero_grad()
    loss = eps_loss(model, x0, t, c, e, betas, cond_drop_prob=0.0)
    loss.backward()
    np.testing.assert_allclose(loss.item(), z["loss"], rtol=1e-5)
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in model.parameters()]
    np.testing.assert_allclose(np.array([g.norm().item() for g in grads]), z["gnorm"], rtol=1e-3, atol=1e-6)
    gs = np.concatenate([g.flatten()[sample_idx(g.numel())].numpy() for g in grads])
    np.testing.assert_allclose(gs, z["gsample"], rtol=1e-2, atol=1e-4)


def test_unet_channel_changing_config_equals_reference():
    """tests/golden/ddpm_small.npz (made from the UNMODIFIED reference): nin_shortcut, the 384-wide skip concatenation whose
    GroupNorm groups straddle the concat boundary, strided down / nearest up sampling, attention with 64 tokens"""
    from tests.golden.make_golden_ddpm import default_init_weights, small_config
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ddpm_small.npz"))
    model = ConditionalUNet(small_config())
    assert [n for n, _ in model.named_parameters()] == list(z["keys"])
    model.load_state_dict(default_init_weights(model))
    model.eval()
    x0, e, t, c = inputs(seed=2, n=4, size=16)
    betas = torch.from_numpy(get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)).float()
    xt = q_sample(x0, t, e, betas)
    np.testing.assert_allclose(model(xt, t.float(), c, mode="train", cond_drop_prob=0.0).detach().numpy(), z["eps_cond"],
                               rtol=1e-4, atol=1e-5)
    drop = torch.ones(4, dtype=torch.bool)
    np.testing.assert_allclose(model(xt, t.float(), c, mode="train", drop_mask=drop).detach().numpy(), z["eps_null"],
                               rtol=1e-4, atol=1e-5)
    model.zero_grad()
    loss = eps_loss(model, x0, t, c, e, betas, cond_drop_prob=0.0)
    loss.backward()
    np.testing.assert_allclose(loss.item(), z["loss"], rtol=1e-5)
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in model.parameters()]
    np.testing.assert_allclose(np.array([g.norm().item() for g in grads]), z["gnorm"], rtol=1e-3, atol=1e-6)
    gs = np.concatenate([g.flatten()[sample_idx(g.numel())].numpy() for g in grads])
    np.testing.assert_allclose(gs, z["gsample"], rtol=1e-2, atol=1e-5)
