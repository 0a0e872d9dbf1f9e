"""Device-resident input pipeline: the training set lives in HBM as uint8, and one kernel per
batch (salun_augment_batch) does what the reference's DataLoader + transforms do on the host -- gather by index,
RandomCrop(32, padding=4), RandomHorizontalFlip, ToTensor (Classification/dataset.py:549-555, main_forget.py:42-48).

``DeviceLoader`` iterates like the reference's loaders -- ``for image, target in loader`` yields a fp32 NCHW batch in
[0, 1] and int64 labels, both already on the GPU (the loops' ``.cuda()`` / ``.to(device)`` are no-ops on them) -- and has
``len()`` and ``.dataset`` like a torch DataLoader.  Shuffling uses torch.randperm on the CPU generator exactly like
RandomSampler, so a run seeded like the reference visits the samples in the reference's order; the crop offsets and flip
decisions come from a device generator (torchvision draws them from Python-side RNG streams that cannot be reproduced
bit for bit; pass ``draw=`` for externally drawn decisions, as the parity tests do).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import torch

from .. import _lib
from .._lib import check
from ..tail import SalunContext, _ptr, _stream


class DeviceDataset:
    """uint8 images [N][H][W][3] (the layout of torchvision's ``CIFAR10.data``) + int64 labels, resident on the GPU."""

    def __init__(self, images_hwc_u8, labels, device=None, ctx: Optional[SalunContext] = None):
        images = torch.as_tensor(images_hwc_u8)
        if images.dtype != torch.uint8 or images.dim() != 4 or images.shape[3] != 3:
            raise ValueError("images must be uint8 [N][H][W][3]")
        self.ctx = ctx if ctx is not None else SalunContext(device)
        self.device = self.ctx.device
        self.images = images.to(self.device).contiguous()
        self.targets = torch.as_tensor(labels).long().to(self.device).contiguous()
        if self.targets.numel() != self.images.shape[0]:
            raise ValueError("labels / images length mismatch")
        self.H, self.W = int(images.shape[1]), int(images.shape[2])
        self._lib = _lib.lib()

    def __len__(self):
        return int(self.images.shape[0])

    @classmethod
    def from_float_nchw(cls, x01: torch.Tensor, labels, **kw):
        """from [0,1] float NCHW tensors (synthetic data): quantised to uint8 like a stored image"""
        return cls((x01.clamp(0, 1) * 255.0).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous(), labels, **kw)

    def batch(self, index: torch.Tensor, crop_xy: Optional[torch.Tensor] = None, flip: Optional[torch.Tensor] = None,
              pad: int = 4, out: Optional[torch.Tensor] = None):
        """(image fp32 [n,3,H,W], target int64 [n]) for dataset positions `index` (int64, device)."""
        index = index.to(self.device, torch.int64).contiguous()
        n = int(index.numel())
        if out is None:
            out = torch.empty(n, 3, self.H, self.W, device=self.device)
        if n == 0:
            return out, self.targets[:0]
        if crop_xy is not None:
            crop_xy = crop_xy.to(self.device, torch.int32).contiguous()
        if flip is not None:
            flip = flip.to(self.device, torch.uint8).contiguous()
        check(self._lib.salun_augment_batch(self.ctx.handle, _ptr(self.images), len(self), _ptr(index), _ptr(crop_xy),
                                            _ptr(flip), n, self.H, self.W, int(pad), _ptr(out), _stream(self.device)),
              "salun_augment_batch")
        return out, self.targets.index_select(0, index)


class DeviceLoader:
    """DataLoader stand-in over a DeviceDataset subset.  augment=True: RandomCrop(H, padding=pad) + RandomHorizontalFlip
    (the reference's train transform); augment=False: ToTensor only (its test transform / no_aug)."""

    def __init__(self, dataset: DeviceDataset, indices=None, batch_size: int = 256, shuffle: bool = False,
                 augment: bool = False, pad: int = 4, drop_last: bool = False,
                 draw: Optional[Callable[[int], tuple]] = None, generator: Optional[torch.Generator] = None):
        self.dataset, self.batch_size, self.shuffle, self.augment, self.pad = dataset, int(batch_size), shuffle, augment, pad
        self.indices = (torch.arange(len(dataset)) if indices is None else torch.as_tensor(indices).long()).cpu()
        self.drop_last, self.draw, self.generator = drop_last, draw, generator
        self._dev_gen = torch.Generator(device=dataset.device)
        self._dev_gen.manual_seed(int(torch.initial_seed()) & 0x7FFFFFFF)

    def __len__(self):
        n = len(self.indices)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.indices)
        order = torch.randperm(n, generator=self.generator) if self.shuffle else torch.arange(n)   # RandomSampler
        order = self.indices[order].to(self.dataset.device)
        dev = self.dataset.device
        for b in range(len(self)):
            idx = order[b * self.batch_size: (b + 1) * self.batch_size]
            k = int(idx.numel())
            crop = flip = None
            if self.augment:
                if self.draw is not None:
                    crop, flip = self.draw(k)
                else:   # RandomCrop.get_params: i, j uniform in [0, 2*pad]; RandomHorizontalFlip: p = 0.5
                    crop = torch.randint(0, 2 * self.pad + 1, (k, 2), device=dev, generator=self._dev_gen, dtype=torch.int32)
                    flip = (torch.rand(k, device=dev, generator=self._dev_gen) < 0.5).to(torch.uint8)
            yield self.dataset.batch(idx, crop, flip, pad=self.pad)
