"""Forget / retain split with the reference's marking convention, restated.

The reference marks forget samples by negating labels (dataset.py:648-705: label := -label-1) inside one
"marked" loader and then splits it again in main (generate_mask.py:120-181, main_forget.py:36-116).  The net effect,
restated here directly:  pick `num_indexes_to_replace` training indices (of class `class_to_replace`, or of all
classes for -1) with np.random.RandomState(seed).choice(..., replace=False), forget = those, retain = the rest.
Data loading / augmentation itself is host-side plumbing outside the hot path."""
from __future__ import annotations

import numpy as np
import torch
from torch.utils.data import DataLoader, TensorDataset


def load_train_test(args):
    """(x_train uint8/float NCHW in [0,1], y_train, x_test, y_test) as tensors."""
    if args.synthetic:
        g = torch.Generator().manual_seed(args.seed)
        n, nt = args.synthetic, max(256, args.synthetic // 10)
        s = args.input_size
        return (torch.rand(n, 3, s, s, generator=g), torch.randint(0, args.num_classes, (n,), generator=g),
                torch.rand(nt, 3, s, s, generator=g), torch.randint(0, args.num_classes, (nt,), generator=g))
    if args.dataset != "cifar10":
        raise NotImplementedError("this mirror loads cifar10 (or --synthetic N); other datasets: pass your own loaders to "
                                  "save_gradient_ratio / get_unlearn_method(...) as the reference's main() does")
    import torchvision
    tr = torchvision.datasets.CIFAR10(args.data, train=True, download=False)
    te = torchvision.datasets.CIFAR10(args.data, train=False, download=False)
    to_t = lambda d: torch.from_numpy(d.data).permute(0, 3, 1, 2).float().div_(255.0)
    return to_t(tr), torch.tensor(tr.targets), to_t(te), torch.tensor(te.targets)


def forget_retain_split(y_train: torch.Tensor, args):
    rng = np.random.RandomState(args.seed)  # dataset.py:663-671 uses a seeded RandomState for the choice
    y = y_train.numpy()
    pool = np.arange(len(y)) if args.class_to_replace == -1 else np.flatnonzero(y == args.class_to_replace)
    k = len(pool) if args.num_indexes_to_replace is None else min(args.num_indexes_to_replace, len(pool))
    forget = np.sort(rng.choice(pool, size=k, replace=False))
    retain = np.setdiff1d(np.arange(len(y)), forget)
    return torch.from_numpy(forget), torch.from_numpy(retain)


def make_loaders(args):
    """OrderedDict(retain, forget, val, test) of DataLoaders like main_forget.py:110-116 (val == test here)."""
    from collections import OrderedDict
    xtr, ytr, xte, yte = load_train_test(args)
    fi, ri = forget_retain_split(ytr, args)
    torch.manual_seed(args.seed)  # utils.setup_seed just before the loaders are built (main_forget.py:38-48)
    mk = lambda x, y, shuffle: DataLoader(TensorDataset(x, y), batch_size=args.batch_size, shuffle=shuffle,
                                          pin_memory=True, num_workers=0)
    return OrderedDict(retain=mk(xtr[ri], ytr[ri], True), forget=mk(xtr[fi], ytr[fi], True),
                       val=mk(xte, yte, False), test=mk(xte, yte, False))
