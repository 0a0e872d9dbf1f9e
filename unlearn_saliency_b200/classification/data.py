"""Forget / retain / val / test split of the reference's CIFAR-10 pipeline, restated index for index.

Reference: ``cifar10_dataloaders`` (dataset.py:529-650) followed by the re-split of the "marked" loader in main
(generate_mask.py:120-181, main_forget.py:36-116):

  1. validation set: for every class, ``RandomState(seed).choice(class_idx, int(0.1 * len(class_idx)), replace=False)``
     drawn from ONE generator in class order (dataset.py:581-589) -> 45 000 train / 5 000 val on CIFAR-10;
  2. the training subset is ``train_set.data[train_idx]`` with ``train_idx = list(set(range(n)) - set(valid_idx))``
     (ascending, dataset.py:595-598);
  3. forget indices are drawn INSIDE that subset by ``replace_class`` with ``RandomState(seed - 1)``
     (dataset.py:604-611, 686-705); ``only_mark`` negates their labels and main() splits on the sign, which is the same as
     forget = those positions, retain = the others;
  4. the test set drops ``class_to_replace`` when the whole class is forgotten (dataset.py:612-614).

Data loading / augmentation itself is host-side plumbing outside the hot path; the resident
uint8 dataset + fused crop / flip kernel that replaces the loader at >100 steps/s is in ``device_data.py``."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
from torch.utils.data import DataLoader, TensorDataset


def load_train_test(args):
    """(x_train float NCHW in [0,1], y_train, x_test, y_test) as tensors."""
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


def train_val_split(y_all: np.ndarray, seed: int):
    """dataset.py:581-598: per-class 10 % validation indices from one RandomState(seed); the rest, ascending, is train."""
    rng = np.random.RandomState(seed)
    valid = []
    for c in range(int(y_all.max()) + 1):
        class_idx = np.where(y_all == c)[0]
        valid.append(rng.choice(class_idx, int(0.1 * len(class_idx)), replace=False))
    valid_idx = np.hstack(valid)
    train_idx = np.setdiff1d(np.arange(len(y_all)), valid_idx)  # list(set(range(n)) - set(valid_idx)) is ascending
    return train_idx, valid_idx


def forget_positions(y_train: np.ndarray, class_to_replace, num_indexes_to_replace, seed: int):
    """replace_class (dataset.py:686-705): positions INSIDE the training subset, RandomState(seed - 1) -- the caller
    passes ``seed`` exactly like cifar10_dataloaders does (``seed=seed - 1`` at dataset.py:609)."""
    if class_to_replace is None:
        return np.zeros(0, dtype=np.int64)
    pool = np.arange(len(y_train)) if class_to_replace == -1 else np.flatnonzero(y_train == class_to_replace)
    if num_indexes_to_replace is not None:
        if num_indexes_to_replace > len(pool):
            raise AssertionError(f"Want to replace {num_indexes_to_replace} indexes but only {len(pool)} samples in dataset")
        pool = np.random.RandomState(seed - 1).choice(pool, size=num_indexes_to_replace, replace=False)
    return np.asarray(pool, dtype=np.int64)


def split_indices(y_all: torch.Tensor, args):
    """-> dict(train, val, forget, retain) of index arrays into the ORIGINAL training set (forget / retain in the order the
    reference's Subset-by-sign split yields them: ascending position inside the training subset)."""
    y = y_all.numpy() if isinstance(y_all, torch.Tensor) else np.asarray(y_all)
    train_idx, valid_idx = train_val_split(y, args.seed)
    pos = forget_positions(y[train_idx], getattr(args, "class_to_replace", None),
                           getattr(args, "num_indexes_to_replace", None), args.seed)
    is_forget = np.zeros(len(train_idx), dtype=bool)
    is_forget[pos] = True
    return dict(train=train_idx, val=valid_idx, forget=train_idx[is_forget], retain=train_idx[~is_forget])


def forget_retain_split(y_train: torch.Tensor, args):
    """(forget, retain) index tensors into the original training set (kept for callers of the round-1 API)."""
    s = split_indices(y_train, args)
    return torch.from_numpy(s["forget"]), torch.from_numpy(s["retain"])


def test_filter(y_test: torch.Tensor, args):
    """dataset.py:612-614: a fully forgotten class is removed from the test set."""
    c, k = getattr(args, "class_to_replace", None), getattr(args, "num_indexes_to_replace", None)
    if c is not None and (k is None or k == 4500):
        return torch.nonzero(y_test != c).flatten()
    return torch.arange(len(y_test))


def make_loaders(args, test_mode: bool = False):
    """OrderedDict(retain, forget, val, test) of loaders like main_forget.py:110-116.  args.device_data: DeviceLoaders over
    uint8 datasets resident on the GPU (device_data.py) with the reference's train transform as a kernel; test_mode=True
    turns shuffling and augmentation off (utils.dataset_convert_to_test, main_forget.py:143)."""
    xtr, ytr, xte, yte = load_train_test(args)
    s = split_indices(ytr, args)
    ti = test_filter(yte, args)
    torch.manual_seed(args.seed)  # utils.setup_seed just before the loaders are built (main_forget.py:38-48)
    if getattr(args, "device_data", False):
        from .device_data import DeviceDataset, DeviceLoader
        dev = f"cuda:{int(getattr(args, 'gpu', 0))}"
        tr = DeviceDataset.from_float_nchw(xtr, ytr, device=dev)
        te = DeviceDataset.from_float_nchw(xte, yte, device=dev, ctx=tr.ctx)
        aug = (not test_mode) and not getattr(args, "no_aug", False)
        mk = lambda ds, idx, train: DeviceLoader(ds, idx, batch_size=args.batch_size, shuffle=train and not test_mode,
                                                 augment=train and aug)
        return OrderedDict(retain=mk(tr, s["retain"], True), forget=mk(tr, s["forget"], True), val=mk(tr, s["val"], False),
                           test=mk(te, ti, False))
    mk = lambda x, y, shuffle: DataLoader(TensorDataset(x, y), batch_size=args.batch_size, shuffle=shuffle,
                                          pin_memory=True, num_workers=0)
    fi, ri, vi = (torch.from_numpy(s[k]) for k in ("forget", "retain", "val"))
    return OrderedDict(retain=mk(xtr[ri], ytr[ri], True), forget=mk(xtr[fi], ytr[fi], True),
                       val=mk(xtr[vi], ytr[vi], False), test=mk(xte[ti], yte[ti], False))
