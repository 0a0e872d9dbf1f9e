"""Golden vectors for the forget / retain / val split: runs the UNMODIFIED reference ``cifar10_dataloaders``
(Classification/dataset.py:529-650) on a fake CIFAR10 class (synthetic labels; the sample index is encoded in the
pixels so that the reference's reshuffled subsets can be traced back).  Build container only (/root/reference).

    python tests/golden/make_golden_split.py  ->  tests/golden/cifar_split.npz
"""
import os
import sys

import numpy as np

REF = "/root/reference/Classification"
HERE = os.path.dirname(os.path.abspath(__file__))


class FakeCIFAR10:
    N = 5000

    def __init__(self, root, train=True, transform=None, download=False):
        n = self.N if train else self.N // 5
        rng = np.random.RandomState(1234 if train else 4321)
        self.targets = list(rng.randint(0, 10, n))
        self.data = np.zeros((n, 32, 32, 3), dtype=np.uint8)
        idx = np.arange(n)
        self.data[:, 0, 0, 0] = idx & 255
        self.data[:, 0, 0, 1] = (idx >> 8) & 255
        self.transform = transform

    def __len__(self):
        return len(self.data)


def ids(ds):
    return ds.data[:, 0, 0, 0].astype(np.int64) | (ds.data[:, 0, 0, 1].astype(np.int64) << 8)


def main():
    sys.path.insert(0, REF)
    import dataset as D
    D.CIFAR10 = FakeCIFAR10
    out = {"labels_train": np.array(FakeCIFAR10("", True).targets), "labels_test": np.array(FakeCIFAR10("", False).targets)}
    for tag, kw in (("rand450_s2", dict(class_to_replace=-1, num_indexes_to_replace=450, seed=2)),
                    ("class3_all_s1", dict(class_to_replace=3, num_indexes_to_replace=None, seed=1)),
                    ("class0_100_s5", dict(class_to_replace=0, num_indexes_to_replace=100, seed=5))):
        tr, va, te = D.cifar10_dataloaders(batch_size=64, data_dir="", only_mark=True, shuffle=True, no_aug=True, **kw)
        t = np.asarray(tr.dataset.targets)
        i = ids(tr.dataset)
        out[tag + "_forget"] = i[t < 0]
        out[tag + "_retain"] = i[t >= 0]
        out[tag + "_val"] = ids(va.dataset)
        out[tag + "_test"] = ids(te.dataset)
    np.savez_compressed(os.path.join(HERE, "cifar_split.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
