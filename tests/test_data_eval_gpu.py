"""salun_augment_batch / salun_eval_logits and their host mirrors (DeviceLoader, validate, collect_prob) on the GPU against
oracle/data.py (pinned to torchvision / torch by tests/test_data_eval_cpu.py)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import classification as OC
from oracle import data as OD

pytestmark = pytest.mark.gpu


def test_augment_batch_bit_exact(salun_ctx):
    from unlearn_saliency_b200.classification.device_data import DeviceDataset
    rng = np.random.RandomState(1)
    imgs = rng.randint(0, 256, (50, 32, 32, 3), dtype=np.uint8)
    labels = rng.randint(0, 10, 50)
    ds = DeviceDataset(imgs, labels, ctx=salun_ctx)
    index = rng.randint(0, 50, 33)
    crop = rng.randint(0, 9, (33, 2)).astype(np.int32)
    flip = rng.randint(0, 2, 33).astype(np.uint8)
    x, y = ds.batch(torch.from_numpy(index), torch.from_numpy(crop), torch.from_numpy(flip))
    assert np.array_equal(x.cpu().numpy(), OD.augment(imgs, index, crop, flip))       # bit-exact (integer / 255)
    assert np.array_equal(y.cpu().numpy(), labels[index])
    x2, _ = ds.batch(torch.from_numpy(index))                                        # test transform: ToTensor only
    assert np.array_equal(x2.cpu().numpy(), OD.augment(imgs, index))
    x3, _ = ds.batch(torch.zeros(0, dtype=torch.int64))                              # empty batch
    assert x3.shape == (0, 3, 32, 32)


def test_device_loader_epoch_covers_every_sample_once(salun_ctx):
    from unlearn_saliency_b200.classification.device_data import DeviceDataset, DeviceLoader
    imgs = np.zeros((103, 32, 32, 3), dtype=np.uint8)
    imgs[:, 16, 16, 0] = np.arange(103)                     # the sample id sits in the centre pixel (survives any crop / flip?)
    ds = DeviceDataset(imgs, np.arange(103) % 10, ctx=salun_ctx)
    torch.manual_seed(7)
    loader = DeviceLoader(ds, indices=np.arange(3, 103), batch_size=32, shuffle=True, augment=False)
    assert len(loader) == 4
    seen, order = [], []
    for x, y in loader:
        assert x.is_cuda and x.dtype == torch.float32 and y.dtype == torch.int64
        ids = (x[:, 0, 16, 16] * 255).round().long().cpu()
        assert torch.equal(ids % 10, y.cpu())
        seen += ids.tolist()
    assert sorted(seen) == list(range(3, 103))
    torch.manual_seed(7)
    assert (np.arange(3, 103)[torch.randperm(100).numpy()]).tolist() == seen      # RandomSampler order of a seeded run
    aug = DeviceLoader(ds, batch_size=64, shuffle=False, augment=True)
    n = sum(int(x.shape[0]) for x, _ in aug)
    assert n == 103


def test_validate_and_collect_prob_vs_oracle(salun_ctx):
    from unlearn_saliency_b200.classification.evaluation import collect_prob, validate
    from unlearn_saliency_b200.engine import ResNetEngine
    params, buffers = OC.synth_state(10, seed=0)
    eng = ResNetEngine("resnet18", 10, 32, max_batch=64, ctx=salun_ctx, precision="split")
    eng.load_state_dict(OC.state_dict_of(params, buffers))
    g = torch.Generator().manual_seed(9)
    x = torch.rand(150, 3, 32, 32, generator=g)
    y = torch.randint(0, 10, (150,), generator=g)
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=64, shuffle=False)
    args = SimpleNamespace(print_freq=100, imagenet_arch=False)
    top1 = validate(loader, eng, torch.nn.CrossEntropyLoss(), args)
    b = {k: v.clone() for k, v in buffers.items()}
    logits = OC.resnet_forward(params, b, x, train=False)
    ce, hits, probs = OD.eval_logits(logits.detach().numpy(), y.numpy())
    assert abs(top1 - 100.0 * hits / 150) <= 100.0 * 1 / 150 + 1e-9        # at most one near-tie decided differently
    assert abs(validate.last_loss - ce / 150) < 2e-3 * max(1.0, ce / 150)
    p, t = collect_prob(loader, eng)
    assert p.shape == (150, 10) and torch.equal(t.cpu(), y)
    np.testing.assert_allclose(p.cpu().numpy(), probs, rtol=5e-3, atol=2e-4)
    np.testing.assert_allclose(p.sum(1).cpu().numpy(), 1.0, rtol=1e-5)
    eng.close()


def test_eval_logits_kernel_exact_counts(salun_ctx):
    from unlearn_saliency_b200 import _lib
    from unlearn_saliency_b200.tail import _ptr, _stream
    g = torch.Generator().manual_seed(3)
    z = (torch.randn(300, 10, generator=g) * 4).cuda()
    y = torch.randint(0, 10, (300,), generator=g).cuda()
    loss = torch.zeros(1, dtype=torch.float64, device="cuda")
    hit = torch.zeros(1, dtype=torch.int64, device="cuda")
    probs = torch.empty_like(z)
    L = _lib.lib()
    for _ in range(2):   # accumulates
        assert L.salun_eval_logits(salun_ctx.handle, _ptr(z), _ptr(y), 300, 10, _ptr(probs), _ptr(loss), _ptr(hit),
                                   _stream(salun_ctx.device)) == 0
    ce, hits, pr = OD.eval_logits(z.cpu().numpy(), y.cpu().numpy())
    assert int(hit) == 2 * hits
    assert abs(float(loss) - 2 * ce) < 1e-4 * ce
    np.testing.assert_allclose(probs.cpu().numpy(), pr, rtol=2e-6, atol=1e-8)
