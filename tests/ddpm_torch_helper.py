"""TEST HELPER: the DDPM loop bodies around a torch ``nn.Module`` with the fused flat tail (FlatParams / FlatMaskedAdam /
FlatSaliency) -- the mechanism the SD mirrors use around the reference's LatentDiffusion object (unlearn_saliency_b200/sd),
exercised here on the DDPM network where a reference-pinned torch model exists.  Not part of the product package: the
DDPM path itself runs on the sm_100a U-Net engine (diffusion/runner.py: DDPMEngineUnlearner)."""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from unlearn_saliency_b200.diffusion.runner import antithetic_t, eps_loss, q_sample
from unlearn_saliency_b200.flat import FlatMaskedAdam, FlatParams, FlatSaliency


class DDPMUnlearner:
    """model + fused tail.  `mask`: the dict torch.load(mask_path) gives (CPU int64, keys as saved); None = no mask."""

    def __init__(self, model: torch.nn.Module, betas, lr=1e-4, beta1=0.9, eps=1e-8, weight_decay=0.0, grad_clip=1.0,
                 mask: Optional[Dict[str, torch.Tensor]] = None, ctx=None):
        self.model = model
        self.device = next(model.parameters()).device
        self.betas = torch.as_tensor(betas).float().to(self.device)
        self.num_timesteps = self.betas.shape[0]
        self.flat = FlatParams(model, ctx)
        if mask is not None:  # DataParallel checkpoints / masks carry a "module." prefix
            mask = {(k[7:] if k.startswith("module.") else k): v for k, v in mask.items()}
        self.opt = FlatMaskedAdam(self.flat, lr=lr, betas=(beta1, 0.999), eps=eps, weight_decay=weight_decay, mask=mask,
                                  max_norm=grad_clip)  # functions/__init__.py:9-18, optim.grad_clip
        self.saliency = FlatSaliency(self.flat, max_norm=grad_clip)

    # ---- runners/diffusion.py:959-996 -------------------------------------------------------------------------
    def generate_mask_batch(self, x, c, cond_scale: float = 2.0, t=None, e=None):
        """x in [0,1] (data_transform 2x-1 applied here, datasets/__init__.py:241-255), eval-mode model."""
        self.model.eval()
        x = 2 * x.to(self.device) - 1.0
        c = c.to(self.device)
        n = x.shape[0]
        e = torch.randn_like(x) if e is None else e.to(self.device)
        t = antithetic_t(n, self.num_timesteps, self.device) if t is None else t.to(self.device)
        xt = q_sample(x, t, e, self.betas)
        out = self.model(xt, t.float(), c, cond_scale=cond_scale, mode="test")  # two U-Net passes, CFG-combined
        loss = (e - out).square().sum(dim=(1, 2, 3)).mean(dim=0)
        self.flat.zero_grad()
        loss.backward()
        self.saliency.accumulate()  # clip to norm 1 (per batch, :985-990) then gradients += grad (:992-996)
        return loss.detach()

    def finish_mask(self, path: Optional[str] = None, ratio: float = 0.5, key_prefix: str = "module."):
        """abs, global top-k, int64 dict saved like results/cifar10/mask/<label>/with_0.5.pt (:998-1039)."""
        self.saliency.all_reduce()
        if path is None:
            return self.saliency.mask(ratio, key_prefix=key_prefix)
        return self.saliency.save(path, ratio, key_prefix=key_prefix)

    # ---- runners/diffusion.py:519-593 -------------------------------------------------------------------------
    def saliency_unlearn_step(self, remain_x, remain_c, forget_x, forget_c, alpha: float = 1e-3, method: str = "rl",
                              n_classes: int = 10, rng: Optional[dict] = None):
        """One iteration.  `rng` may carry externally drawn (t_r, e_r, t_f, e_f, drop masks) for parity runs."""
        rng = rng or {}
        m, dev = self.model, self.device
        m.train()
        xr, cr = 2 * remain_x.to(dev) - 1.0, remain_c.to(dev)
        n = xr.shape[0]
        e = rng.get("e_r", None)
        e = torch.randn_like(xr) if e is None else e.to(dev)
        t = rng.get("t_r", None)
        t = antithetic_t(n, self.num_timesteps, dev) if t is None else t.to(dev)
        kw = {"drop_mask": rng["drop_r"].to(dev)} if "drop_r" in rng else {}
        remain_loss = eps_loss(m, xr, t, cr, e, self.betas, cond_drop_prob=0.1, **kw)             # :523-536
        xf, cf = 2 * forget_x.to(dev) - 1.0, forget_c.to(dev)
        n = xf.shape[0]
        e = rng.get("e_f", None)
        e = torch.randn_like(xf) if e is None else e.to(dev)
        t = rng.get("t_f", None)
        t = antithetic_t(n, self.num_timesteps, dev) if t is None else t.to(dev)
        if method == "ga":
            forget_loss = -eps_loss(m, xf, t, cf, e, self.betas, cond_drop_prob=0.1)              # :552-555
        elif method == "rl":
            xt = q_sample(xf, t, e, self.betas)                                                   # :558-559
            kwf = {"drop_mask": rng["drop_f"].to(dev)} if "drop_f" in rng else {}
            kwp = {"drop_mask": rng["drop_p"].to(dev)} if "drop_p" in rng else {}
            out = m(xt, t.float(), cf, mode="train", **kwf)
            with torch.no_grad():
                pseudo = m(xt, t.float(), (cf + 1) % n_classes, mode="train", **kwp)              # :561-570
            forget_loss = torch.nn.functional.mse_loss(out, pseudo)
        else:
            raise NotImplementedError(method)
        loss = forget_loss + alpha * remain_loss                                                   # :572
        self.opt.zero_grad()
        loss.backward()                                                                            # :579-580
        self.opt.step()   # clip_grad_norm_(1.0) BEFORE the mask, grad *= mask, Adam -- one fused pass (:582-593)
        return loss.detach()

    def save_checkpoint(self, path: str, step: int):
        """states = [model_sd, optim_sd, step] like :598-610 (optimizer state exported in torch.optim.Adam layout)."""
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        sd = {"module." + k: v for k, v in self.model.state_dict().items()}
        optim = {"exp_avg": self.flat.dict_from_flat(self.opt.exp_avg), "exp_avg_sq": self.flat.dict_from_flat(self.opt.exp_avg_sq),
                 "step": self.opt.step_count}
        torch.save([sd, optim, step], path)
