"""mirror of Classification/unlearn/RL.py:37-178 (cifar10 / svhn branch, :109-176) on the engine."""
from __future__ import annotations

import time

import torch

from .impl import iterative_unlearn
from .steps import Meter, accuracy_top1, masked_step


@iterative_unlearn
def RL(data_loaders, model, criterion, optimizer, epoch, args, mask=None):
    forget_loader = data_loaders["forget"]
    retain_loader = data_loaders["retain"]
    if args.dataset not in ("cifar10", "svhn"):
        raise NotImplementedError("only the cifar10/svhn branch of RL (RL.py:109-176) is on the sm_100a path")
    if getattr(args, "warmup", 0) > 0:
        raise NotImplementedError("RL.py:119-121 references an undefined loop variable when warmup > 0")
    losses, top1 = Meter(model.device), Meter(model.device)
    model.train()  # RL.py:114
    start = time.time()
    loader_len = len(forget_loader) + len(retain_loader)
    for i, (image, target) in enumerate(forget_loader):
        # random labels drawn from the CPU generator exactly like RL.py:125
        target = torch.randint(0, args.num_classes, target.shape)
        masked_step(model, optimizer, image, target)
    for i, (image, target) in enumerate(retain_loader):
        loss, logits, tgt = masked_step(model, optimizer, image, target, want_logits=True)
        losses.update(loss, image.size(0))
        top1.update(accuracy_top1(logits, tgt), image.size(0))
        if (i + 1) % args.print_freq == 0:
            end = time.time()
            print("Epoch: [{0}][{1}/{2}]\t" "Loss {3:.4f} ({4:.4f})\t" "Accuracy {5:.3f} ({6:.3f})\t" "Time {7:.2f}".format(
                epoch, i, loader_len, float(losses.val.item()), losses.avg, float(top1.val.item()), top1.avg,
                end - start))
            start = time.time()
    return top1.avg
