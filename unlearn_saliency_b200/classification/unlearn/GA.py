"""mirror of Classification/unlearn/GA.py:44-152: gradient ascent on the forget set (both loader branches: tuples
:107-150, imagenet_arch dict batches :62-105)."""
from __future__ import annotations

import time

from .impl import iterative_unlearn
from .steps import Meter, accuracy_top1, masked_step, unpack_batch, warmup_lr


@iterative_unlearn
def GA(data_loaders, model, criterion, optimizer, epoch, args, mask=None):
    train_loader = data_loaders["forget"]
    losses, top1 = Meter(model.device), Meter(model.device)
    model.train()
    start = time.time()
    for i, data in enumerate(train_loader):
        image, target = unpack_batch(data, args)
        if epoch < getattr(args, "warmup", 0):
            warmup_lr(epoch, i + 1, optimizer, one_epoch_step=len(train_loader), args=args)       # GA.py:109-112
        loss, logits, tgt, n = masked_step(model, optimizer, image, target, loss_sign=-1.0, want_logits=True)  # GA.py:115
        losses.update(loss, n)
        top1.update(accuracy_top1(logits, tgt), n)
        if (i + 1) % args.print_freq == 0:
            end = time.time()
            print("Epoch: [{0}][{1}/{2}]\t" "Loss {3:.4f} ({4:.4f})\t" "Accuracy {5:.3f} ({6:.3f})\t" "Time {7:.2f}".format(
                epoch, i, len(train_loader), float(losses.val.item()), losses.avg, float(top1.val.item()), top1.avg,
                end - start))
            start = time.time()
    avg = top1.avg
    print("train_accuracy {top1:.3f}".format(top1=avg))
    return avg
