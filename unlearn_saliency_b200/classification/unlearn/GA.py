"""mirror of Classification/unlearn/GA.py:44-152 (non-imagenet branch :107-150): gradient ascent on the forget set."""
from __future__ import annotations

import time

from .impl import iterative_unlearn
from .steps import Meter, accuracy_top1, masked_step


@iterative_unlearn
def GA(data_loaders, model, criterion, optimizer, epoch, args, mask=None):
    train_loader = data_loaders["forget"]
    losses, top1 = Meter(model.device), Meter(model.device)
    model.train()
    start = time.time()
    for i, (image, target) in enumerate(train_loader):
        loss, logits, tgt = masked_step(model, optimizer, image, target, loss_sign=-1.0, want_logits=True)  # GA.py:115
        losses.update(loss, image.size(0))
        top1.update(accuracy_top1(logits, tgt), image.size(0))
        if (i + 1) % args.print_freq == 0:
            end = time.time()
            print("Epoch: [{0}][{1}/{2}]\t" "Loss {3:.4f} ({4:.4f})\t" "Accuracy {5:.3f} ({6:.3f})\t" "Time {7:.2f}".format(
                epoch, i, len(train_loader), float(losses.val.item()), losses.avg, float(top1.val.item()), top1.avg,
                end - start))
            start = time.time()
    print("train_accuracy {top1:.3f}".format(top1=top1.avg))
    return top1.avg
