"""mirror of Classification/unlearn/RL.py:37-178 on the engine: the cifar10 / svhn branch (:109-176, forget batches with
fresh random labels, then the retain batches) and the cifar100 / TinyImagenet branch (:51-107, the forget set relabelled
once per epoch and shuffled together with the retain set)."""
from __future__ import annotations

import time

import numpy as np
import torch

from .impl import iterative_unlearn
from .steps import Meter, accuracy_top1, masked_step


def _relabelled_concat_loader(forget_loader, retain_loader, args):
    """RL.py:51-59: np.random.randint labels for the whole forget set, ConcatDataset([forget, retain]), shuffle."""
    fd, rd = forget_loader.dataset, retain_loader.dataset
    n_f = len(fd)
    new_y = torch.from_numpy(np.random.randint(0, args.num_classes, n_f))

    class _Relabelled(torch.utils.data.Dataset):
        def __len__(self):
            return n_f

        def __getitem__(self, i):
            return fd[i][0], new_y[i]

    cat = torch.utils.data.ConcatDataset([_Relabelled(), rd])
    return torch.utils.data.DataLoader(cat, batch_size=args.batch_size, shuffle=True)


@iterative_unlearn
def RL(data_loaders, model, criterion, optimizer, epoch, args, mask=None):
    forget_loader = data_loaders["forget"]
    retain_loader = data_loaders["retain"]
    if getattr(args, "warmup", 0) > 0:
        raise NotImplementedError("RL.py:69-71,119-121 reference an undefined loop variable when warmup > 0")
    losses, top1 = Meter(model.device), Meter(model.device)
    model.train()  # RL.py:66,114
    start = time.time()
    loader_len = len(forget_loader) + len(retain_loader)

    def report(i):
        nonlocal start
        if (i + 1) % args.print_freq == 0:
            end = time.time()
            print("Epoch: [{0}][{1}/{2}]\t" "Loss {3:.4f} ({4:.4f})\t" "Accuracy {5:.3f} ({6:.3f})\t" "Time {7:.2f}".format(
                epoch, i, loader_len, float(losses.val.item()), losses.avg, float(top1.val.item()), top1.avg,
                end - start))
            start = time.time()

    if args.dataset in ("cifar100", "TinyImagenet"):
        train_loader = _relabelled_concat_loader(forget_loader, retain_loader, args)
        for it, (image, target) in enumerate(train_loader):
            i = it + len(forget_loader)                                                           # RL.py:74
            loss, logits, tgt, n = masked_step(model, optimizer, image, target, want_logits=True)
            losses.update(loss, n)
            top1.update(accuracy_top1(logits, tgt), n)
            report(i)
    elif args.dataset in ("cifar10", "svhn"):
        for i, (image, target) in enumerate(forget_loader):
            # random labels drawn from the CPU generator exactly like RL.py:125
            target = torch.randint(0, args.num_classes, target.shape)
            masked_step(model, optimizer, image, target)
        for i, (image, target) in enumerate(retain_loader):
            loss, logits, tgt, n = masked_step(model, optimizer, image, target, want_logits=True)
            losses.update(loss, n)
            top1.update(accuracy_top1(logits, tgt), n)
            report(i)
    # any other dataset: the reference falls through both branches and trains nothing (RL.py:51,109)
    return top1.avg
