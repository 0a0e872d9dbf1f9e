"""Fused SalUn tail for ANY torch.nn.Module whose forward/backward still runs in PyTorch (DDPM / SD U-Nets this round).

The reference's DDPM and SD loops (DDPM/runners/diffusion.py:519-593, 959-1039; SD/train-scripts/train-esd.py:267-323,
random_label.py:66-143, generate_mask.py:33-108) do, per step and per parameter tensor:
    clip_grad_norm_  ->  grad *= mask[name].to(device)  (a 309 MB / 6.9 GB int64 H2D copy EVERY step)  ->  Adam.step()
and, for the mask:  gradients[name] += grad.cpu()  (D2H every batch)  ->  CPU double argsort.
Here the module's parameters and gradients are re-pointed at two flat fp32 arenas, so that the whole tail is four
kernel launches over the arena (include/salun.h):
    salun_grad_sumsq -> salun_clip_coef -> salun_masked_adam_step        (FlatMaskedAdam.step)
    salun_saliency_accumulate_flat [* clip coefficient] -> salun_topk_mask  (FlatSaliency)
with the mask resident on the GPU as 1 bit per parameter.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict
from typing import Dict, Optional

import torch

from .tail import SalunContext, mask_words, topk_count


class FlatParams:
    """Re-points module.parameters() (and their .grad) at flat fp32 CUDA arenas, in named_parameters() order."""

    def __init__(self, module: torch.nn.Module, ctx: Optional[SalunContext] = None):
        named = [(n, p) for n, p in module.named_parameters()]
        if not named:
            raise ValueError("module has no parameters")
        dev = named[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("FlatParams needs the module on a CUDA device (no CPU fallback)")
        for n, p in named:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError(f"parameter {n} must be fp32 on {dev}")
        self.module, self.ctx = module, ctx if ctx is not None else SalunContext(dev)
        self.names = [n for n, _ in named]
        self.shapes = OrderedDict((n, tuple(p.shape)) for n, p in named)
        self.offsets, off = {}, 0
        for n, p in named:
            self.offsets[n] = off
            off += (p.numel() + 3) // 4 * 4  # keep every tensor 16-byte aligned inside the arena
        self.n = off
        self.numel = sum(p.numel() for _, p in named)
        self.params = torch.zeros(self.n, device=dev)
        self.grads = torch.zeros(self.n, device=dev)
        with torch.no_grad():
            for n, p in named:
                o, c = self.offsets[n], p.numel()
                self.params[o: o + c] = p.detach().reshape(-1)
                p.data = self.params[o: o + c].view_as(p)
                p.grad = self.grads[o: o + c].view_as(p)

    def zero_grad(self):
        self.grads.zero_()  # p.grad views stay attached: autograd accumulates straight into the arena

    def flat_from_dict(self, d: Dict[str, torch.Tensor], dtype=None) -> torch.Tensor:
        out = torch.zeros(self.n, dtype=dtype or next(iter(d.values())).dtype, device=self.params.device)
        for n in self.names:
            key = n if n in d else n.split("model.diffusion_model.")[-1]  # SD consumer strips the prefix, train-esd.py:321
            t = d[key]
            out[self.offsets[n]: self.offsets[n] + t.numel()] = t.reshape(-1).to(out.device)
        return out

    def dict_from_flat(self, flat: torch.Tensor, cpu: bool = False) -> "OrderedDict[str, torch.Tensor]":
        res = OrderedDict()
        for n, shp in self.shapes.items():
            t = flat[self.offsets[n]: self.offsets[n] + math.prod(shp)].reshape(shp).clone()
            res[n] = t.cpu() if cpu else t
        return res

    def pad_mask(self) -> torch.Tensor:
        """1 for real parameter slots, 0 for the alignment padding between tensors (never selected, never updated)"""
        m = torch.zeros(self.n, dtype=torch.int64, device=self.params.device)
        for n, shp in self.shapes.items():
            m[self.offsets[n]: self.offsets[n] + math.prod(shp)] = 1
        return m


class FlatMaskedAdam:
    """clip_grad_norm_ (before masking) + grad*=mask + torch.optim.Adam.step(), fused.
    DDPM/runners/diffusion.py:582-593 with get_optimizer (DDPM/functions/__init__.py:9-18); SD: max_norm=None."""

    def __init__(self, flat: FlatParams, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                 mask: Optional[Dict[str, torch.Tensor]] = None, max_norm: Optional[float] = None):
        self.flat, self.ctx = flat, flat.ctx
        self.lr, self.betas, self.eps, self.wd, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.exp_avg = torch.zeros_like(flat.params)
        self.exp_avg_sq = torch.zeros_like(flat.params)
        self.step_count = 0
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=flat.params.device)
        self._coef = torch.ones(1, device=flat.params.device)
        self.mask_bits = None
        if mask is not None:
            m = flat.flat_from_dict(mask, dtype=torch.int64)
            self.mask_bits = self.ctx.pack_mask(m.contiguous())

    def zero_grad(self):
        self.flat.zero_grad()

    def grad_norm(self) -> torch.Tensor:
        """pre-clip total norm (device scalar), as returned by clip_grad_norm_"""
        return self._sumsq.sqrt()

    def step(self):
        self.step_count += 1
        coef = None
        if self.max_norm is not None:
            self.ctx.grad_sumsq(self.flat.grads, self._sumsq)
            coef = self.ctx.clip_coef(self._sumsq, self.max_norm, self._coef)
        self.ctx.masked_adam_step(self.flat.params, self.flat.grads, self.exp_avg, self.exp_avg_sq, self.mask_bits,
                                  self.lr, self.betas[0], self.betas[1], self.eps, self.wd, self.step_count, coef)


class FlatSaliency:
    """gradients[name] += clip(grad) ; abs ; top-k ; save  -- DDPM/runners/diffusion.py:985-1039,
    SD/train-scripts/generate_mask.py:66-108 -- without the per-batch .cpu() copies or the CPU double argsort."""

    def __init__(self, flat: FlatParams, max_norm: Optional[float] = None):
        self.flat, self.ctx, self.max_norm = flat, flat.ctx, max_norm
        self.acc = torch.zeros_like(flat.params)
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=flat.params.device)
        self._coef = torch.ones(1, device=flat.params.device)

    def accumulate(self):
        """call after loss.backward() (and instead of clip_grad_norm_ + the per-tensor += loop)"""
        coef = None
        if self.max_norm is not None:
            self.ctx.grad_sumsq(self.flat.grads, self._sumsq)
            coef = self.ctx.clip_coef(self._sumsq, self.max_norm, self._coef)
        self.ctx.saliency_accumulate_flat(self.flat.grads, self.acc, coef)

    def all_reduce(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.acc)

    def mask(self, ratio: float = 0.5, cpu: bool = True, key_prefix: str = ""):
        """{name: int64 0/1 tensor} in the reference format (CPU tensors for DDPM/SD, runners/diffusion.py:995,1039)."""
        if getattr(self.flat, "native_layout", False):
            # engine arena (conv weights OHWI, no padding) -> the reference's flat order, so that ties at the threshold
            # resolve in the same element order as torch.cat([g.flatten() ...]) (runners/diffusion.py:1006-1008)
            dense = self.flat.from_native_flat(self.acc).contiguous()
        else:
            # compact the arena (drop alignment padding) so that k = int(N * ratio) counts real parameters only
            pieces = [self.acc[self.flat.offsets[n]: self.flat.offsets[n] + math.prod(s)] for n, s in self.flat.shapes.items()]
            dense = torch.cat(pieces).contiguous()
        k = topk_count(dense.numel(), ratio)
        m64, _, info = self.ctx.topk_mask(dense, k, want_bits=False, want_info=True)
        out, off = OrderedDict(), 0
        for n, s in self.flat.shapes.items():
            c = math.prod(s)
            t = m64[off: off + c].reshape(s)
            out[key_prefix + n] = t.cpu() if cpu else t.clone()
            off += c
        return out, info

    def save(self, path: str, ratio: float = 0.5, wait: bool = True, **kw):
        """the reference's mask file (CPU int64 dict) written through the streaming saver (io.py); wait=False returns while
        the file is still being written (call io.default_saver().wait() before reading it back)"""
        from .io import default_saver
        m, info = self.mask(ratio, cpu=False, **{k: v for k, v in kw.items() if k != "cpu"})
        saver = default_saver()
        saver.save(m, path)          # staged to pinned host memory on a side stream, pickled on a worker thread
        if wait:
            saver.wait()
        return info
