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
        if n <= 0:      # a rank whose shard of a trailing partial batch is empty has nothing to record
            return
        self.val = val.detach().reshape(1).clone()
        self.sum += self.val.double() * n
        self.count += n

    @property
    def avg(self):
        """Data parallel: the (sum, count) pairs of all ranks are combined, so every rank reports the global average."""
        s, c = self.sum.clone(), torch.tensor([float(self.count)], device=self.sum.device, dtype=torch.float64)
        _, world = dist_info()
        if world > 1:
            torch.distributed.all_reduce(s)
            torch.distributed.all_reduce(c)
        return float(s.item() / max(1.0, float(c.item())))


def unpack_batch(data, args=None):
    """(image, target) of one loader item: tuples for the CIFAR-style loaders, {"image", "label"} dicts for the
    imagenet_arch loaders (imagenet.get_x_y_from_data_dict, GA.py:66 / FT.py:67)."""
    if isinstance(data, dict):
        return data["image"], data["label"]
    return data[0], data[1]


def warmup_lr(epoch, step, optimizer, one_epoch_step, args):
    """utils.warmup_lr (utils.py:33-41): linear ramp of args.lr over the first args.warmup epochs."""
    overall_steps = args.warmup * one_epoch_step
    current_steps = epoch * one_epoch_step + step
    lr = min(args.lr * current_steps / overall_steps, args.lr)
    for p in optimizer.param_groups:
        p["lr"] = lr


def masked_step(engine, optimizer, image, target, loss_sign=1.0, want_logits=False, l1_alpha=0.0):
    """output = model(image); loss = sign*criterion(output, target) [+ l1_alpha*||theta||_1]; zero_grad; backward; mask;
    step; restore (RL.py:128-140, FT.py:131-144).  With torch.distributed initialised the mini-batch is sharded across
    ranks and the gradient is averaged before the fused masked update.
    Returns (loss, logits, target, n_local): loss / logits of THIS rank's shard (empty logits for an empty shard)."""
    rank, world = dist_info()
    if world > 1:
        n = image.shape[0]
        per = (n + world - 1) // world
        lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
        image, target = image[lo:hi], target[lo:hi]
    image = image.to(engine.device, non_blocking=True).float().contiguous()
    target = target.to(engine.device, non_blocking=True).long().contiguous()
    n_local = image.shape[0]
    if n_local > 0:
        loss, logits = engine.forward_backward(image, target, loss_sign=loss_sign, want_logits=want_logits, train=True)
    else:  # a rank without samples in the last partial batch contributes zeros
        engine.grads.zero_()
        loss = torch.zeros(1, device=engine.device)
        logits = torch.zeros(0, engine.num_classes, device=engine.device) if want_logits else None
    if world > 1:
        # per-rank mean CE over its shard -> weight by shard size to recover the global batch mean
        w = n_local * world / float(n)
        if w != 1.0:
            engine.grads.mul_(w)
        if not hasattr(optimizer, "momentum_shard"):  # plain MaskedSGD: NCCL all-reduce, then the local fused step
            torch.distributed.all_reduce(engine.grads)
            engine.grads.div_(world)
        # DistMaskedSGD averages the peers' gradients inside its kernel
    if l1_alpha:
        # the penalty depends on the (replicated) weights only: added after the rank average.  The fused DP optimizer
        # averages inside its kernel, so there every rank adds the same term before it (mean of identical terms)
        l1 = engine.ctx.l1_penalty_grad(engine.params, engine.grads, l1_alpha)
        loss = loss + l1_alpha * l1.float()
    optimizer.step()
    return loss, logits, target, n_local


def accuracy_top1(logits, target):
    """utils.accuracy(output, target)[0] (utils.py:321-334), as a device scalar in percent."""
    if logits.shape[0] == 0:
        return torch.zeros((), device=logits.device)
    return (logits.argmax(1) == target).float().mean() * 100.0


def sync_bn_buffers(engine):
    """Data parallel: BatchNorm running statistics are per shard; average them across ranks (what a checkpoint of a
    DataParallel / SyncBN run would hold up to the shard-variance term) before they are exported."""
    _, world = dist_info()
    if world > 1:
        for t in (engine.running_mean, engine.running_var):
            torch.distributed.all_reduce(t)
            t.div_(world)
