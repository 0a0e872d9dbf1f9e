"""mirror of Classification/unlearn/impl.py:54-127 (``iterative_unlearn``) on the fused optimizer."""
from __future__ import annotations

import time

from ...engine import DistMaskedSGD, MaskedSGD
from ..common import as_engine, check_criterion, dist_info, sync_to_module
from .steps import sync_bn_buffers


def _lr_at(base_lr, epoch, milestones, gamma=0.1):
    """MultiStepLR(milestones, gamma=0.1) (impl.py:95-97): lr of `epoch` after `epoch` scheduler.step() calls."""
    return base_lr * (gamma ** sum(1 for m in milestones if epoch >= m))


def _iterative_unlearn_impl(unlearn_iter_func):
    def _wrapped(data_loaders, model, criterion, args, mask=None, **kwargs):
        check_criterion(criterion)
        if getattr(args, "rewind_epoch", 0) != 0:
            raise NotImplementedError("weight rewinding (impl.py:58-67, 98-101) is outside the SalUn hot path")
        decreasing_lr = list(map(int, args.decreasing_lr.split(",")))
        _, world = dist_info()
        engine = as_engine(model, args, symmetric=world > 1)
        # `if mask:` RL.py:134.  A path (str) instead of the loaded dict lets the packed side-car be used (io.py)
        bits = (engine.mask_bits_from_file(mask) if isinstance(mask, str) else engine.mask_bits_from_dict(mask)) if mask else None
        if world > 1 and engine.symmetric:
            # data parallel: gradient reduce-scatter + masked SGD + weight all-gather in ONE kernel over NVLink peer memory
            optimizer = DistMaskedSGD(engine, args.unlearn_lr, momentum=args.momentum, weight_decay=args.weight_decay,
                                      mask_bits=bits)
        else:
            optimizer = MaskedSGD(engine, args.unlearn_lr, momentum=args.momentum, weight_decay=args.weight_decay,
                                  mask_bits=bits)  # impl.py:68-73 + RL.py:11-34
        train_acc = None
        for epoch in range(0, args.unlearn_epochs):
            start_time = time.time()
            optimizer.param_groups[0]["lr"] = _lr_at(args.unlearn_lr, epoch, decreasing_lr)
            print("Epoch #{}, Learning rate: {}".format(epoch, optimizer.param_groups[0]["lr"]))
            train_acc = unlearn_iter_func(data_loaders, engine, criterion, optimizer, epoch, args, mask, **kwargs)
            print("one epoch duration:{}".format(time.time() - start_time))
        sync_bn_buffers(engine)
        sync_to_module(engine)
        return train_acc

    return _wrapped


def iterative_unlearn(func):
    """usage: @iterative_unlearn  def func(data_loaders, model, criterion, optimizer, epoch, args, mask)"""
    return _iterative_unlearn_impl(func)
