"""Mirrors of the SD SalUn entry points with the reference's signatures:

  train_esd(prompt, train_method, start_guidance, negative_guidance, iterations, lr, config_path, ckpt_path, mask_path,
            diffusers_config_path, devices, seperator, image_size, ddim_steps)        SD/train-scripts/train-esd.py:129-343
  certain_label(class_to_forget, train_method, alpha, batch_size, epochs, lr, ...)    SD/train-scripts/random_label.py:13-160
  generate_mask(classes, c_guidance, batch_size, epochs, lr, ...)                     SD/train-scripts/generate_mask.py:8-108

`model` is the reference's ``LatentDiffusion`` (ldm/models/diffusion/ddpm.py): its ``model.diffusion_model`` parameters and
gradients are re-pointed at flat fp32 arenas (flat.FlatParams), so

  * ``p.grad *= mask[name].to(device)`` + ``opt.step()``  (train-esd.py:318-323, random_label.py:132-139) become ONE
    salun_masked_adam_step over the arena with a resident 1-bit mask (107 MB instead of a 6.9 GB int64 upload per step);
    parameters left out by ``train_method`` are masked out the same way (zero gradient => zero Adam update, moments stay 0);
  * ``gradients[name] += param.grad.data.cpu()``  (generate_mask.py:66-69) becomes salun_saliency_accumulate_flat on the
    device and the CPU double argsort over 859.5 M keys becomes salun_topk_mask.

The SD stack (pytorch_lightning, omegaconf, CLIP, the VAE) is not importable in the build container, so model
construction, data loaders and DDIM sampling are the reference's own objects, passed in (`models=`, `loaders=`,
`sample_fn=`); without them the entry points try the reference's loaders (``get_models`` / ``setup_model`` /
``setup_forget_data``) and raise a clear error if the SD stack is missing.
"""
from __future__ import annotations

import os
import random
from typing import Callable, Dict, List, Optional

import torch

from ..flat import FlatMaskedAdam, FlatParams, FlatSaliency


def select_parameters(names: List[str], train_method: str) -> List[str]:
    """names of model.model.diffusion_model parameters trained by `train_method` (train-esd.py:192-224,
    random_label.py:45-54)"""
    out = []
    for name in names:
        if train_method == "noxattn":
            if not (name.startswith("out.") or "attn2" in name or "time_embed" in name):
                out.append(name)
        elif train_method == "selfattn":
            if "attn1" in name:
                out.append(name)
        elif train_method == "xattn":
            if "attn2" in name:
                out.append(name)
        elif train_method == "full":
            out.append(name)
        elif train_method == "notime":
            if not (name.startswith("out.") or "time_embed" in name):
                out.append(name)
        elif train_method == "xlayer":
            if "attn2" in name and ("output_blocks.6." in name or "output_blocks.8." in name):
                out.append(name)
        elif train_method == "selflayer":
            if "attn1" in name and ("input_blocks.4." in name or "input_blocks.7." in name):
                out.append(name)
    return out


class SDTail:
    """flat arenas over model.model.diffusion_model + the fused mask (.) grad + Adam step and saliency tail"""

    def __init__(self, model, lr: float = 1e-5, train_method: str = "full",
                 mask: Optional[Dict[str, torch.Tensor]] = None, ctx=None):
        self.unet = model.model.diffusion_model
        self.flat = FlatParams(self.unet, ctx)
        names = self.flat.names
        selected = set(select_parameters(names, train_method))
        if not selected:
            raise ValueError(f"train_method {train_method!r} selects no parameter")
        self.trained = selected
        combined = None
        if mask is not None or len(selected) != len(names):
            combined = {}
            for n, shp in self.flat.shapes.items():
                if n not in selected:
                    combined[n] = torch.zeros(shp, dtype=torch.int64)
                elif mask is not None:
                    combined[n] = mask[n.split("model.diffusion_model.")[-1]].to(torch.int64)   # train-esd.py:321
                else:
                    combined[n] = torch.ones(shp, dtype=torch.int64)
        self.opt = FlatMaskedAdam(self.flat, lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, mask=combined,
                                  max_norm=None)                      # torch.optim.Adam(parameters, lr=lr), no clip
        self.saliency = FlatSaliency(self.flat, max_norm=None)


def esd_iteration(model, model_orig, sample_fn: Callable, tail: SDTail, word: str, start_guidance: float,
                  negative_guidance: float, image_size: int = 512, ddim_steps: int = 50, devices=None, rng=None):
    """One iteration of train_esd (train-esd.py:268-323).  sample_fn(emb, guidance, start_code, t_enc) is the reference's
    quick_sample_till_t (DDIM partial sampling with CFG, no grad).  `rng` may carry t_enc, t_enc_ddpm, start_code."""
    rng = rng or {}
    d0 = devices[0] if devices else next(tail.unet.parameters()).device
    d1 = devices[1] if devices else d0
    emb_0 = model.get_learned_conditioning([""])
    emb_p = model.get_learned_conditioning([word])
    emb_n = model.get_learned_conditioning([f"{word}"])
    tail.opt.zero_grad()
    t_enc = rng["t_enc"] if "t_enc" in rng else torch.randint(ddim_steps, (1,), device=d0)
    og_num = round((int(t_enc) / ddim_steps) * 1000)
    og_num_lim = round((int(t_enc + 1) / ddim_steps) * 1000)
    t_enc_ddpm = rng["t_enc_ddpm"] if "t_enc_ddpm" in rng else torch.randint(og_num, og_num_lim, (1,), device=d0)
    start_code = rng["start_code"] if "start_code" in rng else torch.randn((1, 4, image_size // 8, image_size // 8)).to(d0)
    with torch.no_grad():
        z = sample_fn(emb_p.to(d0), start_guidance, start_code, int(t_enc))
        if hasattr(model_orig, "apply_model_pair"):     # engine-backed frozen model: both conditionings in one pass
            e_0, e_p = model_orig.apply_model_pair(z.to(d1), t_enc_ddpm.to(d1), emb_0.to(d1), emb_p.to(d1))
        else:
            e_0 = model_orig.apply_model(z.to(d1), t_enc_ddpm.to(d1), emb_0.to(d1))
            e_p = model_orig.apply_model(z.to(d1), t_enc_ddpm.to(d1), emb_p.to(d1))
    e_n = model.apply_model(z.to(d0), t_enc_ddpm.to(d0), emb_n.to(d0))
    target = e_0.to(d0) - (negative_guidance * (e_p.to(d0) - e_0.to(d0)))
    loss = torch.nn.functional.mse_loss(e_n.to(d0), target)                     # :301-311
    loss.backward()
    tail.opt.step()          # grad *= mask (+ train_method selection) and Adam in one pass (:318-323)
    return loss.detach()


def _reference_sd_stack(what: str):
    raise RuntimeError(
        f"{what}: the Stable Diffusion stack of the reference (ldm, pytorch_lightning, omegaconf, CLIP, the VAE and its "
        "checkpoints) is not importable here; pass the reference's own objects (models=..., loaders=..., sample_fn=...)")


def engine_passes(model, model_orig, tail: "SDTail", precision: str, image_size: int = 512, ddim_steps: int = 50, ctx=None):
    """The no-grad passes of ESD on the U-Net engine (sd/engine.py + sd/sampler.py): returns (frozen, sample_fn) where
    `frozen.apply_model` serves e_0 / e_p of model_orig (train-esd.py:295-297) and sample_fn is quick_sample_till_t
    (:281-283) sampling from `model`'s CURRENT weights -- the trained tensors are pushed into the engine before every call."""
    from .engine import SDUNetEngine, config_from_module
    from .sampler import EngineApplyModel, EngineDDIMSampler, make_quick_sample_till_t
    unet, unet_orig = model.model.diffusion_model, model_orig.model.diffusion_model
    uncond = model.get_learned_conditioning([""])
    kw = dict(latent_size=image_size // 8, max_batch=2, context_len=int(uncond.shape[1]), ctx=ctx, precision=precision)
    sampling = SDUNetEngine(config_from_module(unet), **kw).load_state_dict(dict(unet.named_parameters()))
    frozen = SDUNetEngine(config_from_module(unet_orig), **kw).load_state_dict(dict(unet_orig.named_parameters()))
    sampler = EngineDDIMSampler(sampling, alphas_cumprod=getattr(model, "alphas_cumprod", None))
    sample_fn = make_quick_sample_till_t(sampler, uncond, image_size=image_size, ddim_steps=ddim_steps, module=unet,
                                         train_keys=sorted(tail.trained))
    return EngineApplyModel(frozen), sample_fn


def train_esd(prompt, train_method, start_guidance, negative_guidance, iterations, lr, config_path, ckpt_path, mask_path,
              diffusers_config_path, devices, seperator=None, image_size=512, ddim_steps=50, *, models=None,
              sample_fn=None, on_save=None, ctx=None, engine: Optional[str] = None):
    """train-esd.py:129-343.  models = (model_orig, sampler_orig, model, sampler) from the reference's get_models().
    engine = "bf16" | "split": the DDIM partial sampling and the frozen model's two passes run on the U-Net engine
    (engine_passes); otherwise sample_fn is the caller's quick_sample_till_t."""
    if prompt == "allartist":
        prompt = "Kelly Mckernan, Thomas Kinkade, Ajin Demi Human, Alena Aenami, Tyler Edlin, Kilian Eng"
    if prompt == "i2p":
        prompt = "hate, harassment, violence, suffering, humiliation, harm, suicide, sexual, nudity, bodily fluids, blood"
    words = [w.strip() for w in prompt.split(seperator)] if seperator is not None else [prompt]
    if models is None or (sample_fn is None and engine is None):
        _reference_sd_stack("train_esd")
    model_orig, _sampler_orig, model, _sampler = models
    mask = torch.load(mask_path) if mask_path else None
    tail = SDTail(model, lr=lr, train_method=train_method, mask=mask, ctx=ctx)
    frozen = model_orig
    if engine is not None:
        frozen, sample_fn = engine_passes(model, model_orig, tail, engine, image_size, ddim_steps, ctx=ctx)
    model.train()
    losses = []
    for i in range(iterations):
        word = random.sample(words, 1)[0]
        loss = esd_iteration(model, frozen, sample_fn, tail, word, start_guidance, negative_guidance, image_size,
                             ddim_steps, devices)
        losses.append(float(loss))
        if on_save is not None and (i + 1) % 500 == 0 and i + 1 != iterations:
            on_save(model, i - 1)
    model.eval()
    if on_save is not None:
        on_save(model, None)
    return losses


def mask_batch(model, tail: SDTail, images, prompts, c_guidance: float, rng=None):
    """One batch of SD generate_mask (generate_mask.py:33-69): -MSE(noise, (1+s) eps(c) - s eps(null)), eval mode."""
    rng = rng or {}
    device = images.device
    null_prompts = ["" for _ in prompts]
    forget_batch = {"jpg": images.permute(0, 2, 3, 1), "txt": list(prompts)}
    null_batch = {"jpg": images.permute(0, 2, 3, 1), "txt": null_prompts}
    forget_input, forget_emb = model.get_input(forget_batch, model.first_stage_key)
    _null_input, null_emb = model.get_input(null_batch, model.first_stage_key)
    t = rng["t"] if "t" in rng else torch.randint(0, model.num_timesteps, (forget_input.shape[0],), device=device).long()
    noise = rng["noise"] if "noise" in rng else torch.randn_like(forget_input, device=device)
    forget_noisy = model.q_sample(x_start=forget_input, t=t, noise=noise)
    forget_out = model.apply_model(forget_noisy, t, forget_emb)
    null_out = model.apply_model(forget_noisy, t, null_emb)
    preds = (1 + c_guidance) * forget_out - c_guidance * null_out
    loss = -torch.nn.functional.mse_loss(noise, preds)
    tail.flat.zero_grad()
    loss.backward()
    tail.saliency.accumulate()       # gradients[name] += grad, on the device (no .cpu() round trip)
    return loss.detach()


def generate_mask(classes, c_guidance, batch_size, epochs, lr, config_path, ckpt_path, diffusers_config_path, device,
                  image_size=512, num_timesteps=1000, *, model=None, loader=None, descriptions=None, ratio=0.5, ctx=None):
    """generate_mask.py:8-108: writes mask/<classes>/with_0.5.pt (CPU int64 dict keyed by diffusion_model parameter names)."""
    if model is None or loader is None or descriptions is None:
        _reference_sd_stack("generate_mask")
    model.eval()
    tail = SDTail(model, lr=lr, train_method="full", mask=None, ctx=ctx)
    for images, labels in loader:
        mask_batch(model, tail, images.to(device), [descriptions[int(l)] for l in labels], c_guidance)
    path = os.path.join("mask", str(classes), f"with_{ratio}.pt")
    return tail.saliency.save(path, ratio)


def certain_label_step(model, tail: SDTail, remain_images, remain_prompts, forget_images, forget_prompts, pseudo_prompts,
                       alpha: float, rng=None):
    """One step of random_label.certain_label (random_label.py:77-139)."""
    rng = rng or {}
    tail.opt.zero_grad()
    remain_loss = model.shared_step({"jpg": remain_images.permute(0, 2, 3, 1), "txt": list(remain_prompts)})[0]
    forget_input, forget_emb = model.get_input({"jpg": forget_images.permute(0, 2, 3, 1), "txt": list(forget_prompts)},
                                               model.first_stage_key)
    pseudo_input, pseudo_emb = model.get_input({"jpg": forget_images.permute(0, 2, 3, 1), "txt": list(pseudo_prompts)},
                                               model.first_stage_key)
    t = rng["t"] if "t" in rng else torch.randint(0, model.num_timesteps, (forget_input.shape[0],),
                                                   device=forget_input.device).long()
    noise = rng["noise"] if "noise" in rng else torch.randn_like(forget_input)
    forget_noisy = model.q_sample(x_start=forget_input, t=t, noise=noise)
    pseudo_noisy = model.q_sample(x_start=pseudo_input, t=t, noise=noise)
    forget_out = model.apply_model(forget_noisy, t, forget_emb)
    pseudo_out = model.apply_model(pseudo_noisy, t, pseudo_emb).detach()
    loss = torch.nn.functional.mse_loss(forget_out, pseudo_out) + alpha * remain_loss
    loss.backward()
    tail.opt.step()
    return loss.detach()


def certain_label(class_to_forget, train_method, alpha, batch_size, epochs, lr, config_path, ckpt_path, mask_path,
                  diffusers_config_path, device, image_size=512, ddim_steps=50, *, model=None, loaders=None,
                  descriptions=None, on_save=None, ctx=None):
    """random_label.py:13-160.  loaders = (remain_dl, forget_dl) from the reference's setup_remain_data / setup_forget_data."""
    if model is None or loaders is None or descriptions is None:
        _reference_sd_stack("certain_label")
    remain_dl, forget_dl = loaders
    mask = torch.load(mask_path) if mask_path else None
    tail = SDTail(model, lr=lr, train_method=train_method, mask=mask, ctx=ctx)
    model.train()
    losses = []
    for _epoch in range(epochs):
        remain_iter = iter(remain_dl)
        for forget_images, forget_labels in forget_dl:
            try:
                remain_images, remain_labels = next(remain_iter)
            except StopIteration:
                remain_iter = iter(remain_dl)
                remain_images, remain_labels = next(remain_iter)
            loss = certain_label_step(
                model, tail, remain_images.to(device), [descriptions[int(l)] for l in remain_labels],
                forget_images.to(device), [descriptions[int(l)] for l in forget_labels],
                [descriptions[(int(class_to_forget) + 1) % 10] for _ in forget_labels], alpha)
            losses.append(float(loss) / batch_size)
    model.eval()
    if on_save is not None:
        on_save(model, None)
    return losses
