"""mirror of Classification/unlearn/FT.py:44-180: fine-tune on the retain set (FT), optionally with the decaying
alpha * ||theta||_1 penalty (FT_l1, FT.py:13-17,122-134); tuple and imagenet_arch dict batches."""
from __future__ import annotations

import time

from .impl import iterative_unlearn
from .steps import Meter, accuracy_top1, masked_step, unpack_batch, warmup_lr


def FT_iter(data_loaders, model, criterion, optimizer, epoch, args, mask=None, with_l1=False):
    train_loader = data_loaders["retain"]
    losses, top1 = Meter(model.device), Meter(model.device)
    model.train()
    start = time.time()
    for i, data in enumerate(train_loader):
        image, target = unpack_batch(data, args)
        if epoch < getattr(args, "warmup", 0):
            warmup_lr(epoch, i + 1, optimizer, one_epoch_step=len(train_loader), args=args)       # FT.py:117-120
        current_alpha = 0.0
        if with_l1:                                                                                # FT.py:124-129
            span = args.unlearn_epochs - args.no_l1_epochs
            current_alpha = args.alpha * (1 - epoch / span) if epoch < span else 0.0
        loss, logits, tgt, n = masked_step(model, optimizer, image, target, want_logits=True, l1_alpha=current_alpha)
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


@iterative_unlearn
def FT(data_loaders, model, criterion, optimizer, epoch, args, mask=None):
    return FT_iter(data_loaders, model, criterion, optimizer, epoch, args, mask)


@iterative_unlearn
def FT_l1(data_loaders, model, criterion, optimizer, epoch, args, mask=None):
    return FT_iter(data_loaders, model, criterion, optimizer, epoch, args, mask, with_l1=True)
