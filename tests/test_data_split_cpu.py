"""The host mirror's forget / retain / val / test split against the UNMODIFIED reference pipeline
(Classification/dataset.py:529-650 run on a fake CIFAR10 class, tests/golden/make_golden_split.py)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from unlearn_saliency_b200.classification import data as D

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "cifar_split.npz"))


@pytest.mark.parametrize("tag,cls,num,seed", [("rand450_s2", -1, 450, 2), ("class3_all_s1", 3, None, 1),
                                              ("class0_100_s5", 0, 100, 5)])
def test_split_matches_reference(tag, cls, num, seed):
    args = SimpleNamespace(seed=seed, class_to_replace=cls, num_indexes_to_replace=num)
    y = torch.from_numpy(G["labels_train"])
    s = D.split_indices(y, args)
    # the reference's loaders hold the subsets in ascending position order; the sets AND the order must agree
    assert np.array_equal(s["forget"], G[tag + "_forget"])
    assert np.array_equal(s["retain"], G[tag + "_retain"])
    assert np.array_equal(s["val"], G[tag + "_val"])
    ti = D.test_filter(torch.from_numpy(G["labels_test"]), args)
    assert np.array_equal(ti.numpy(), G[tag + "_test"])
    assert len(np.intersect1d(s["val"], np.concatenate([s["forget"], s["retain"]]))) == 0


def test_full_class_forget_removes_class_from_test_only_at_4500():
    y = torch.from_numpy(G["labels_test"])
    keep = D.test_filter(y, SimpleNamespace(class_to_replace=3, num_indexes_to_replace=100))
    assert len(keep) == len(y)
    keep = D.test_filter(y, SimpleNamespace(class_to_replace=3, num_indexes_to_replace=4500))
    assert (y[keep] != 3).all() and len(keep) < len(y)
