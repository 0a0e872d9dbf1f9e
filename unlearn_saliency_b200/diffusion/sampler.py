"""Class-conditional sampling on the U-Net engine's forward kernels: mirror of

  compute_alpha / generalized_steps_conditional     DDPM/functions/denoising.py:4-7, 72-95
  Diffusion.sample_image / sample_visualization     DDPM/runners/diffusion.py:828-931

Every step runs the conditional and the null pass of classifier-free guidance as ONE engine batch of 2n (GroupNorm does
not couple samples) and one fused update kernel (salun_ddim_step); nothing is copied to the host between steps (the
reference moves every x_t and x0 prediction to the CPU, denoising.py:86,95).  In the reference this sampling -- 100 images
x 1000 steps at every snapshot -- costs far more than the unlearning itself.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import numpy as np
import torch

from .._lib import check
from ..tail import _ptr, _stream


def compute_alpha(beta: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """denoising.py:4-7 as a flat [n] tensor: alpha-bar of t, with alpha-bar(-1) = 1"""
    beta = torch.cat([torch.zeros(1, device=beta.device), beta], dim=0)
    return (1 - beta).cumprod(dim=0).index_select(0, t + 1)


def timestep_sequence(num_timesteps: int, timesteps: int, skip_type: str = "uniform") -> Sequence[int]:
    """runners/diffusion.py:834-847"""
    if skip_type == "uniform":
        return list(range(0, num_timesteps, num_timesteps // timesteps))
    if skip_type == "quad":
        return [int(s) for s in list(np.linspace(0, np.sqrt(num_timesteps * 0.8), timesteps) ** 2)]
    raise NotImplementedError(skip_type)


class EngineSampler:
    def __init__(self, engine, betas):
        self.engine, self.device = engine, engine.device
        self.betas = torch.as_tensor(betas).float().to(self.device)
        self.num_timesteps = int(self.betas.shape[0])

    @torch.no_grad()
    def generalized_steps_conditional(self, x, c, seq, cond_scale: float = 3.0, eta: float = 0.0, noise_fn=None,
                                      keep: bool = False):
        """denoising.py:72-95.  Returns (xs, x0_preds): lists with every step when keep=True, else only the last state
        (what sample_image consumes).  noise_fn(step_index, like) supplies the eta noise (default torch.randn_like)."""
        eng, dev = self.engine, self.device
        x = x.to(dev).float().contiguous()
        c = c.to(dev).long().contiguous()
        n, chw = x.shape[0], x[0].numel()
        if 2 * n > eng.max_batch and cond_scale != 0:
            raise ValueError(f"batch {n}: the guided sampler runs 2n = {2 * n} images per step, engine max_batch is {eng.max_batch}")
        seq = list(seq)
        seq_next = [-1] + seq[:-1]
        drop = torch.cat([torch.zeros(n, dtype=torch.uint8, device=dev), torch.ones(n, dtype=torch.uint8, device=dev)])
        c2, xs, x0s = torch.cat([c, c]), [x], []
        lib, ctx = eng._lib, eng.ctx
        for k, (i, j) in enumerate(zip(reversed(seq), reversed(seq_next))):
            t = torch.full((n,), float(i), device=dev)
            at = compute_alpha(self.betas, t.long()).contiguous()
            at_next = compute_alpha(self.betas, torch.full((n,), j, device=dev, dtype=torch.long)).contiguous()
            xt = xs[-1]
            if cond_scale == 0:
                eps_c, eps_n = eng.forward(xt, t, c, drop=drop[:n].contiguous(), save=False, train=False), None
            else:
                eps2 = eng.forward(torch.cat([xt, xt]), torch.cat([t, t]), c2, drop=drop, save=False, train=False)
                eps_c, eps_n = eps2[:n], eps2[n:]
            noise = None
            if eta != 0:
                noise = (noise_fn(k, xt) if noise_fn is not None else torch.randn_like(xt)).contiguous()
            x_next = torch.empty_like(xt)
            x0 = torch.empty_like(xt) if keep else None
            check(lib.salun_ddim_step(ctx.handle, _ptr(eps_c), _ptr(eps_n), _ptr(xt), _ptr(noise), _ptr(at), _ptr(at_next),
                                      float(cond_scale), float(eta), n, chw, _ptr(x_next), _ptr(x0), _stream(dev)),
                  "salun_ddim_step")
            if keep:
                xs.append(x_next)
                x0s.append(x0)
            else:
                xs = [x_next]
        return xs, x0s

    def sample_image(self, x, c, cond_scale, sample_type: str = "generalized", skip_type: str = "uniform",
                     timesteps: int = 1000, eta: float = 1.0, last: bool = True):
        """runners/diffusion.py:828-875"""
        if sample_type != "generalized":
            # the reference's "ddpm_noisy" branch imports a function that does not exist (ddpm_steps_conditional, :868)
            raise NotImplementedError(f"sample_type {sample_type!r}: only 'generalized' runs in the reference")
        seq = timestep_sequence(self.num_timesteps, timesteps, skip_type)
        xs, x0s = self.generalized_steps_conditional(x, c, seq, cond_scale, eta=eta, keep=not last)
        return xs[-1] if last else (xs, x0s)

    def sample_visualization(self, n_classes: int, total_n_samples: int, batch_size: int, cond_scale: float, image_size: int,
                             channels: int = 3, path: Optional[str] = None, **kw):
        """runners/diffusion.py:877-931: total_n_samples images, the same number per class, as one image grid in [0, 1]"""
        assert total_n_samples % n_classes == 0
        n_rounds = total_n_samples // batch_size if batch_size < total_n_samples else 1
        c = torch.repeat_interleave(torch.arange(n_classes), total_n_samples // n_classes).to(self.device)
        imgs = []
        for cc in torch.chunk(c, n_rounds, dim=0):
            x = torch.randn(cc.size(0), channels, image_size, image_size, device=self.device)
            x = self.sample_image(x, cc, cond_scale, **kw)
            imgs.append(torch.clamp((x + 1.0) / 2.0, 0.0, 1.0))       # inverse_data_transform (rescaled data)
        imgs = torch.cat(imgs)
        if path is not None:
            import torchvision.utils as tvu
            os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
            grid = tvu.make_grid(imgs, nrow=total_n_samples // n_classes, normalize=True, padding=0)
            tvu.save_image(grid, path)
        return imgs
