"""oracle/data.py against the torchvision functional ops behind the reference's train transform
(Classification/dataset.py:549-555: RandomCrop(32, padding=4), RandomHorizontalFlip, ToTensor) and torch's cross-entropy."""
import numpy as np
import torch

from oracle import data as OD


def test_augment_equals_torchvision_ops():
    import torchvision.transforms.functional as TF
    from PIL import Image
    rng = np.random.RandomState(0)
    imgs = rng.randint(0, 256, (6, 32, 32, 3), dtype=np.uint8)
    index = [3, 0, 5, 5, 1]
    crop = [(0, 0), (8, 8), (4, 4), (1, 7), (6, 2)]
    flip = [0, 1, 0, 1, 1]
    out = OD.augment(imgs, index, crop, flip, pad=4)
    for i, src in enumerate(index):
        im = Image.fromarray(imgs[src])
        im = TF.pad(im, 4)                                           # RandomCrop: F.pad(img, self.padding, fill=0, "constant")
        im = TF.crop(im, crop[i][1], crop[i][0], 32, 32)             # F.crop(img, i=top, j=left, h, w)
        if flip[i]:
            im = TF.hflip(im)
        assert np.array_equal(out[i], TF.to_tensor(im).numpy())
    plain = OD.augment(imgs, [2], None, None)
    assert np.array_equal(plain[0], TF.to_tensor(Image.fromarray(imgs[2])).numpy())


def test_eval_logits_equals_torch():
    g = torch.Generator().manual_seed(0)
    z = torch.randn(37, 10, generator=g) * 3
    y = torch.randint(0, 10, (37,), generator=g)
    ce, hits, probs = OD.eval_logits(z.numpy(), y.numpy())
    assert abs(ce - float(torch.nn.functional.cross_entropy(z.double(), y, reduction="sum"))) < 1e-9
    assert hits == int((z.argmax(1) == y).sum())
    np.testing.assert_allclose(probs, torch.softmax(z, -1).numpy(), rtol=1e-6, atol=1e-8)
