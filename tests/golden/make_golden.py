"""Generate the golden fixtures that pin oracle/ against the UNMODIFIED reference.

Run in the build container only (it imports /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors for this path, so the fixtures
are outputs of the reference's own functions on seeded synthetic inputs:

  tiny_mask.npz       generate_mask.save_gradient_ratio on a ~5k-parameter CNN: all 10 masks, complete.
  resnet18_mask.npz   generate_mask.save_gradient_ratio on resnet18 (synthetic weights, 2 x 16 images):
                      packed masks for ratios 0.1 / 0.5, ones-count for all ratios.
  resnet18_grad.npz   logits, loss, per-tensor gradient norms + sampled entries of one eval-mode -CE backward
                      and one train-mode CE backward of the reference resnet18.
  resnet18_rl.npz     unlearn.RL (1 epoch: 2 forget + 2 retain batches of 16) with a 50% mask: sampled final
                      parameters, per-tensor norms, BN running statistics, and the random labels it drew.

The import shims (fake matplotlib, trainer.train_with_rewind, identity .cuda()) are the ones documented in
SURVEY.md Appendix C; no reference file is modified or copied.
"""
import argparse
import hashlib
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/Classification"


def import_reference():
    sys.path.insert(0, REF)
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    pkg = types.ModuleType("trainer")
    pkg.__path__ = [REF + "/trainer"]
    sys.modules["trainer"] = pkg
    sp = importlib.util.spec_from_file_location("trainer.train", REF + "/trainer/train.py")
    tt = importlib.util.module_from_spec(sp)
    sys.modules["trainer.train"] = tt
    sp.loader.exec_module(tt)
    tt.train_with_rewind = lambda *a, **k: None
    sp2 = importlib.util.spec_from_file_location("trainer", REF + "/trainer/__init__.py",
                                                 submodule_search_locations=[REF + "/trainer"])
    tp = importlib.util.module_from_spec(sp2)
    sys.modules["trainer"] = tp
    sp2.loader.exec_module(tp)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    import generate_mask  # noqa
    import unlearn  # noqa
    from models import model_dict  # noqa
    return generate_mask, unlearn, model_dict


def sample_idx(n, k=64):
    g = np.random.default_rng(n)
    return np.sort(g.choice(n, size=min(k, n), replace=False))


def ref_args(save_dir, **kw):
    a = argparse.Namespace(unlearn_lr=0.013, momentum=0.9, weight_decay=5e-4, save_dir=save_dir, dataset="cifar10",
                           num_classes=10, warmup=0, print_freq=1000, unlearn_epochs=1, decreasing_lr="91,136",
                           rewind_epoch=0, imagenet_arch=False, unlearn="RL", batch_size=16, gpu=0, no_l1_epochs=0,
                           alpha=0.0)
    for k, v in kw.items():
        setattr(a, k, v)
    return a


class TinyNet(torch.nn.Module):
    """fixture-only model (any nn.Module works with save_gradient_ratio)"""

    def __init__(self):
        super().__init__()
        self.conv1 = torch.nn.Conv2d(3, 12, 3, padding=1, bias=False)
        self.bn1 = torch.nn.BatchNorm2d(12)
        self.conv2 = torch.nn.Conv2d(12, 24, 3, padding=1, stride=2, bias=False)
        self.fc = torch.nn.Linear(24, 10)

    def forward(self, x):
        x = torch.relu(self.bn1(self.conv1(x)))
        x = torch.relu(self.conv2(x))
        return self.fc(x.mean((2, 3)))


def tiny_state(seed=0):
    g = torch.Generator().manual_seed(seed)
    m = TinyNet()
    sd = m.state_dict()
    for k, v in sd.items():
        if v.dtype.is_floating_point:
            sd[k] = torch.randn(v.shape, generator=g) * 0.2 + (1.0 if k.endswith("running_var") else 0.0)
            if k.endswith("running_var"):
                sd[k] = sd[k].abs() + 0.5
    m.load_state_dict(sd)
    return m


def main():
    torch.set_num_threads(8)
    gm, unlearn, model_dict = import_reference()
    from oracle import classification as OC
    crit = torch.nn.CrossEntropyLoss()

    # ------------------------------------------------------------------ tiny model, complete masks
    g = torch.Generator().manual_seed(7)
    x = torch.rand(48, 3, 8, 8, generator=g)
    y = torch.randint(0, 10, (48,), generator=g)
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=16, shuffle=False)
    with tempfile.TemporaryDirectory() as d:
        gm.save_gradient_ratio({"forget": loader}, tiny_state(0), crit, ref_args(d))
        masks = {}
        for r in OC.THRESHOLDS:
            md = torch.load(os.path.join(d, f"with_{r}.pt"))
            assert all(v.dtype == torch.int64 for v in md.values())
            masks[str(r)] = torch.cat([v.flatten() for v in md.values()]).numpy().astype(np.uint8)
        keys = list(md.keys())
    np.savez_compressed(os.path.join(HERE, "tiny_mask.npz"), x=x.numpy(), y=y.numpy(), keys=np.array(keys),
                        **{f"mask_{k}": v for k, v in masks.items()})
    print("tiny_mask.npz", {k: int(v.sum()) for k, v in masks.items()})

    # ------------------------------------------------------------------ resnet18
    params, buffers = OC.synth_state(10, seed=0)
    sd = OC.state_dict_of(params, buffers)
    model = model_dict["resnet18"](num_classes=10)
    missing = model.load_state_dict(sd, strict=True)
    assert list(dict(model.named_parameters()).keys()) == list(params.keys())
    g = torch.Generator().manual_seed(11)
    x = torch.rand(32, 3, 32, 32, generator=g)
    y = torch.randint(0, 10, (32,), generator=g)
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=16, shuffle=False)

    # one eval-mode -CE backward and one train-mode CE backward
    out = {}
    for tag, train, sign in (("eval", False, -1.0), ("train", True, 1.0)):
        m2 = model_dict["resnet18"](num_classes=10)
        m2.load_state_dict(sd)
        m2.train(train)
        logits = m2(x[:16])
        loss = sign * crit(logits, y[:16])
        loss.backward()
        out[f"{tag}_logits"] = logits.detach().numpy()
        out[f"{tag}_loss"] = np.float32(loss.item())
        out[f"{tag}_gnorm"] = np.array([p.grad.norm().item() for p in m2.parameters()], dtype=np.float64)
        out[f"{tag}_gsample"] = np.concatenate([p.grad.flatten()[sample_idx(p.numel())].numpy() for p in m2.parameters()])
        if train:
            out["train_rm_bn1"] = m2.bn1.running_mean.numpy().copy()
            out["train_rv_l4"] = m2.layer4[1].bn2.running_var.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "resnet18_grad.npz"), **out)
    print("resnet18_grad.npz", out["eval_loss"], out["train_loss"])

    with tempfile.TemporaryDirectory() as d:
        gm.save_gradient_ratio({"forget": loader}, model, crit, ref_args(d))
        res = {}
        for r in OC.THRESHOLDS:
            md = torch.load(os.path.join(d, f"with_{r}.pt"))
            flat = torch.cat([v.flatten() for v in md.values()]).numpy()
            res[f"ones_{r}"] = np.int64(flat.sum())
            res[f"sha_{r}"] = np.array(hashlib.sha256(flat.astype(np.uint8).tobytes()).hexdigest())
            if r in (0.1, 0.5):
                res[f"bits_{r}"] = np.packbits(flat.astype(np.uint8), bitorder="little")
            if r == 0.5:
                mask05 = md
    np.savez_compressed(os.path.join(HERE, "resnet18_mask.npz"), **res)
    print("resnet18_mask.npz", {k: int(v) for k, v in res.items() if k.startswith("ones")})

    # ------------------------------------------------------------------ resnet50, ImageNet stem (BASELINE config 4 shape, 64x64 here)
    p50, b50 = OC.synth_state_bottleneck(10, seed=0, depth=50, imagenet=True)
    m50 = model_dict["resnet50"](num_classes=10, imagenet=True)
    assert [n for n, _ in m50.named_parameters()] == list(p50.keys())
    sd50 = OC.state_dict_of(p50, b50)
    m50.load_state_dict(sd50)
    g = torch.Generator().manual_seed(21)
    x50 = torch.rand(4, 3, 64, 64, generator=g)
    y50 = torch.randint(0, 10, (4,), generator=g)
    out50 = {}
    for tag, train, sign in (("eval", False, -1.0), ("train", True, 1.0)):
        mm = model_dict["resnet50"](num_classes=10, imagenet=True)
        mm.load_state_dict(sd50)
        mm.train(train)
        logits = mm(x50)
        loss = sign * crit(logits, y50)
        loss.backward()
        out50[f"{tag}_logits"] = logits.detach().numpy()
        out50[f"{tag}_loss"] = np.float32(loss.item())
        out50[f"{tag}_gnorm"] = np.array([p.grad.norm().item() for p in mm.parameters()], dtype=np.float64)
    out50["n_params"] = np.int64(sum(p.numel() for p in m50.parameters()))
    out50["n_params_1000"] = np.int64(sum(p.numel() for p in model_dict["resnet50"](num_classes=1000, imagenet=True).parameters()))
    np.savez_compressed(os.path.join(HERE, "resnet50_grad.npz"), **out50)
    print("resnet50_grad.npz", out50["eval_loss"], out50["train_loss"], out50["n_params_1000"])

    # ------------------------------------------------------------------ RL epoch with the 0.5 mask
    model = model_dict["resnet18"](num_classes=10)
    model.load_state_dict(sd)
    fl = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=16, shuffle=False)
    g = torch.Generator().manual_seed(13)
    xr = torch.rand(32, 3, 32, 32, generator=g)
    yr = torch.randint(0, 10, (32,), generator=g)
    rl = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(xr, yr), batch_size=16, shuffle=False)
    drawn = []
    orig_randint = torch.randint

    def logging_randint(*a, **k):
        t = orig_randint(*a, **k)
        drawn.append(t.clone())
        return t

    torch.manual_seed(123)
    torch.randint = logging_randint
    try:
        unlearn.RL({"forget": fl, "retain": rl}, model, crit, ref_args("/tmp"), mask05)
    finally:
        torch.randint = orig_randint
    fin = dict(model.named_parameters())
    bufs = dict(model.named_buffers())
    res = dict(
        rand_labels=torch.stack(drawn).numpy(),
        pnorm=np.array([p.detach().norm().item() for p in fin.values()], dtype=np.float64),
        psample=np.concatenate([p.detach().flatten()[sample_idx(p.numel())].numpy() for p in fin.values()]),
        rm_bn1=bufs["bn1.running_mean"].numpy(), rv_l4=bufs["layer4.1.bn2.running_var"].numpy(),
        nbt=np.int64(bufs["bn1.num_batches_tracked"].item()),
    )
    # a single forget step (tight tolerance: no chaotic amplification of rounding differences yet)
    model = model_dict["resnet18"](num_classes=10)
    model.load_state_dict(sd)
    f1 = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x[:16], y[:16]), batch_size=16, shuffle=False)
    r0 = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x[:0], y[:0]), batch_size=16, shuffle=False)
    drawn1 = []
    drawn, keep = drawn1, drawn
    torch.manual_seed(321)
    torch.randint = lambda *a, **k: (drawn1.append(orig_randint(*a, **k)), drawn1[-1])[1]
    try:
        unlearn.RL({"forget": f1, "retain": r0}, model, crit, ref_args("/tmp"), mask05)
    finally:
        torch.randint = orig_randint
    res["rand_labels_1"] = drawn1[0].numpy()
    res["psample_1"] = np.concatenate([p.detach().flatten()[sample_idx(p.numel())].numpy() for p in model.parameters()])
    np.savez_compressed(os.path.join(HERE, "resnet18_rl.npz"), **res)
    print("resnet18_rl.npz labels", res["rand_labels"].shape, "nbt", res["nbt"])


if __name__ == "__main__":
    main()
