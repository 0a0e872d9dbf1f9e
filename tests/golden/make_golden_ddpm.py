"""Golden fixture pinning unlearn_saliency_b200/diffusion (U-Net, eps-loss, q-sample) against the UNMODIFIED reference
DDPM code.  Run in the build container only:  python tests/golden/make_golden_ddpm.py   (separate process from
make_golden.py: both reference trees define top-level `models`, `datasets`, ...).

ddpm_tiny.npz: for a small config (ch 128 -- the only width the reference's hard-coded cemb_channels=512 admits --, mult [1,1],
1 res block, attention at 4x4, 8x8 images, dropout 0) with weights
from a seeded formula: eps-prediction of Conditional_Model in "test" mode (cond_scale 2) and "train" mode
(cond_drop_prob 0), noise_estimation_loss_conditional, per-parameter gradient norms, sampled gradient entries; plus the
named_parameters() key lists of the tiny AND the full cifar10 config (334 keys), and get_beta_schedule('linear').
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference/DDPM")
sys.path.insert(1, ROOT)


def tiny_config():
    return SimpleNamespace(
        model=SimpleNamespace(type="conditional", in_channels=3, out_ch=3, ch=128, ch_mult=[1, 1], num_res_blocks=1,
                              attn_resolutions=[4], dropout=0.0, resamp_with_conv=True, cond_drop_prob=0.1),
        data=SimpleNamespace(image_size=8, channels=3, n_classes=10),
        diffusion=SimpleNamespace(beta_schedule="linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000),
    )


def small_config():
    """two levels with a channel change: nin_shortcut (128 -> 256), a 384-wide skip concatenation whose GroupNorm groups
    straddle the concat boundary, strided-conv down / nearest up sampling, attention with 64 tokens"""
    return SimpleNamespace(
        model=SimpleNamespace(type="conditional", in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2], num_res_blocks=1,
                              attn_resolutions=[8], dropout=0.0, resamp_with_conv=True, cond_drop_prob=0.1),
        data=SimpleNamespace(image_size=16, channels=3, n_classes=10),
        diffusion=SimpleNamespace(beta_schedule="linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000),
    )


def default_init_weights(model, seed=0):
    """PyTorch-default-scale weights (uniform +-1/sqrt(fan_in)) from a seeded formula, norms 1 +- 0.1, biases +-0.05:
    keeps activations O(1) through the deeper config (synth_weights' 0.1 * randn grows them by ~3x per conv)"""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in model.state_dict().items():
        if "norm" in k and k.endswith("weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith("bias") or v.dim() == 1:
            sd[k] = 0.05 * torch.randn(v.shape, generator=g)
        else:
            fan_in = v[0].numel()
            sd[k] = (torch.rand(v.shape, generator=g) * 2 - 1) / fan_in ** 0.5
    return sd


def synth_weights(model, seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in model.state_dict().items():
        if "norm" in k and k.endswith("weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        else:
            sd[k] = 0.1 * torch.randn(v.shape, generator=g)
    return sd


def inputs(seed=1, n=6, size=8):
    g = torch.Generator().manual_seed(seed)
    x0 = torch.rand(n, 3, size, size, generator=g) * 2 - 1
    e = torch.randn(n, 3, size, size, generator=g)
    t = torch.randint(0, 1000, (n,), generator=g)
    c = torch.randint(0, 10, (n,), generator=g)
    return x0, e, t, c


def sample_idx(n, k=32):
    return np.sort(np.random.default_rng(n).choice(n, size=min(k, n), replace=False))


def main():
    from models.diffusion import Conditional_Model
    from functions.losses import noise_estimation_loss_conditional
    from runners.diffusion import get_beta_schedule
    from unlearn_saliency_b200.diffusion.config import cifar10_config

    cfg = tiny_config()
    model = Conditional_Model(cfg)
    sd = synth_weights(model)
    model.load_state_dict(sd)
    model.eval()
    x0, e, t, c = inputs()
    betas = torch.from_numpy(get_beta_schedule(beta_schedule="linear", beta_start=1e-4, beta_end=0.02,
                                               num_diffusion_timesteps=1000)).float()
    a = (1 - betas).cumprod(dim=0).index_select(0, t).view(-1, 1, 1, 1)
    xt = x0 * a.sqrt() + e * (1.0 - a).sqrt()
    out = {}
    out["eps_test"] = model(xt, t.float(), c, cond_scale=2.0, mode="test").detach().numpy()
    out["eps_train"] = model(xt, t.float(), c, mode="train", cond_drop_prob=0.0).detach().numpy()
    model.zero_grad()
    loss = noise_estimation_loss_conditional(model, x0, t, c, e, betas, cond_drop_prob=0.0)
    loss.backward()
    out["loss"] = np.float32(loss.item())
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in model.parameters()]  # null emb unused at p=0
    out["gnorm"] = np.array([g.norm().item() for g in grads])
    out["gsample"] = np.concatenate([g.flatten()[sample_idx(g.numel())].numpy() for g in grads])
    out["keys_tiny"] = np.array([n for n, _ in model.named_parameters()])
    out["betas"] = betas.numpy()
    full = Conditional_Model(cifar10_config())
    out["keys_full"] = np.array([n for n, _ in full.named_parameters()])
    out["numel_full"] = np.int64(sum(p.numel() for p in full.parameters()))
    np.savez_compressed(os.path.join(HERE, "ddpm_tiny.npz"), **out)
    print("ddpm_tiny.npz loss", out["loss"], "keys", len(out["keys_tiny"]), len(out["keys_full"]), out["numel_full"])
    main_small(Conditional_Model, noise_estimation_loss_conditional, betas)


def main_small(Conditional_Model, noise_estimation_loss_conditional, betas):
    """ddpm_small.npz: the channel-changing config (small_config) -- eps in "train" mode with externally chosen
    class-dropout decisions reproduced through cond_drop_prob in {0, 1} per half, the eps loss, every parameter's gradient
    norm and sampled gradient entries, from the UNMODIFIED reference model."""
    cfg = small_config()
    model = Conditional_Model(cfg)
    model.load_state_dict(default_init_weights(model))
    model.eval()
    x0, e, t, c = inputs(seed=2, n=4, size=16)
    a = (1 - betas).cumprod(dim=0).index_select(0, t).view(-1, 1, 1, 1)
    xt = x0 * a.sqrt() + e * (1.0 - a).sqrt()
    out = {}
    out["eps_cond"] = model(xt, t.float(), c, mode="train", cond_drop_prob=0.0).detach().numpy()
    out["eps_null"] = model(xt, t.float(), c, mode="train", cond_drop_prob=1.0).detach().numpy()
    model.zero_grad()
    loss = noise_estimation_loss_conditional(model, x0, t, c, e, betas, cond_drop_prob=0.0)
    loss.backward()
    out["loss"] = np.float32(loss.item())
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in model.parameters()]
    out["gnorm"] = np.array([g.norm().item() for g in grads])
    out["gsample"] = np.concatenate([g.flatten()[sample_idx(g.numel())].numpy() for g in grads])
    out["keys"] = np.array([n for n, _ in model.named_parameters()])
    np.savez_compressed(os.path.join(HERE, "ddpm_small.npz"), **out)
    print("ddpm_small.npz loss", out["loss"], "keys", len(out["keys"]), "max |eps|", np.abs(out["eps_cond"]).max())


if __name__ == "__main__":
    main()
