"""The shared loop body of RL / GA / FT: one masked unlearning step on the engine, optionally data parallel."""
from __future__ import annotations

import torch

from ..common import dist_info


class Meter:
    """utils.AverageMeter (utils.py:64-80) kept ON THE DEVICE: no .item() sync per step (RL.py:166-167 syncs each)."""

    def __init__(self, device):
        self.sum = torch.zeros(1, device=device, dtype=torch.float64)
        self.count = 0
        self.val = torch.zeros(1, device=device)

    def update(self, val, n):
        self.val = val.detach().reshape(1).clone()
        self.sum += self.val.double() * n
        self.count += n

    @property
    def avg(self):
        return float(self.sum.item() / max(1, self.count))


def masked_step(engine, optimizer, image, target, loss_sign=1.0, want_logits=False):
    """output = model(image); loss = sign*criterion(output, target); zero_grad; backward; mask; step; restore
    (RL.py:128-140).  With torch.distributed initialised the mini-batch is sharded across ranks and the gradient
    is averaged with ONE all-reduce before the fused masked update."""
    rank, world = dist_info()
    if world > 1:
        n = image.shape[0]
        per = (n + world - 1) // world
        lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
        image, target = image[lo:hi], target[lo:hi]
    image = image.to(engine.device, non_blocking=True).float().contiguous()
    target = target.to(engine.device, non_blocking=True).long().contiguous()
    if image.shape[0] > 0:
        loss, logits = engine.forward_backward(image, target, loss_sign=loss_sign, want_logits=want_logits, train=True)
    else:  # a rank without samples in the last partial batch contributes zeros
        engine.grads.zero_()
        loss, logits = torch.zeros(1, device=engine.device), None
    if world > 1:
        # per-rank mean CE over its shard -> weight by shard size to recover the global batch mean
        w = image.shape[0] * world / float(n)
        if w != 1.0:
            engine.grads.mul_(w)
        if not hasattr(optimizer, "momentum_shard"):  # plain MaskedSGD: NCCL all-reduce, then the local fused step
            torch.distributed.all_reduce(engine.grads)
            engine.grads.div_(world)
        # DistMaskedSGD averages the peers' gradients inside its kernel
    optimizer.step()
    return loss, logits, target


def accuracy_top1(logits, target):
    """utils.accuracy(output, target)[0] (utils.py:321-334), as a device scalar in percent."""
    return (logits.argmax(1) == target).float().mean() * 100.0
