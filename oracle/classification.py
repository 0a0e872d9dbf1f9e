"""oracle/classification.py -- CPU restatement (torch fp32) of the Classification hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates, functionally and without nn.Module
classes, what the reference computes; citations are relative to /root/reference/Classification.
Pinned against outputs of the unmodified reference by tests/golden/make_golden.py ->
tests/golden/*.npz -> tests/test_oracle_golden.py.

  resnet_forward          models/ResNet.py:303-322 (+ BasicBlock.forward :108-124, NormalizeByChannelMeanStd :23-28)
  save_gradient_ratio     generate_mask.py:14-82
  rl_epoch / ga / ft      unlearn/RL.py:109-176, unlearn/GA.py:107-128, unlearn/FT.py:116-144 with
                          _apply_mask_to_grads (RL.py:11-14), torch.optim.SGD (impl.py:68-73),
                          _restore_masked_params (RL.py:17-34)
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Iterable, List, Tuple

import torch
import torch.nn.functional as F

CIFAR_MEAN = (0.4914, 0.4822, 0.4465)  # models/ResNet.py:214-216
CIFAR_STD = (0.2470, 0.2435, 0.2616)

# ---------------------------------------------------------------------------------------------
# architecture table of resnet18 with the CIFAR stem (models/ResNet.py:217-223, 232-243, 336)
# ---------------------------------------------------------------------------------------------
STAGES = [(64, 1), (128, 2), (256, 2), (512, 2)]  # (planes, stride of the first block)
BLOCKS = {18: (2, 2, 2, 2), 34: (3, 4, 6, 3)}      # BasicBlocks per stage: resnet18 (ResNet.py:336), resnet34 (:347)


def resnet18_param_shapes(num_classes: int = 10, depth: int = 18) -> "OrderedDict[str, Tuple[int, ...]]":
    """named_parameters() order of the reference's resnet18 (62 tensors, 11 173 962 elements for 10 classes);
    depth=34 gives resnet34 (BasicBlock [3,4,6,3])."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    s["conv1.weight"] = (64, 3, 3, 3)
    s["bn1.weight"] = (64,)
    s["bn1.bias"] = (64,)
    inpl = 64
    for li, (planes, stride) in enumerate(STAGES, start=1):
        for b in range(BLOCKS[depth][li - 1]):
            pre = f"layer{li}.{b}."
            st = stride if b == 0 else 1
            s[pre + "conv1.weight"] = (planes, inpl, 3, 3)
            s[pre + "bn1.weight"] = (planes,)
            s[pre + "bn1.bias"] = (planes,)
            s[pre + "conv2.weight"] = (planes, planes, 3, 3)
            s[pre + "bn2.weight"] = (planes,)
            s[pre + "bn2.bias"] = (planes,)
            if b == 0 and (st != 1 or inpl != planes):
                s[pre + "downsample.0.weight"] = (planes, inpl, 1, 1)
                s[pre + "downsample.1.weight"] = (planes,)
                s[pre + "downsample.1.bias"] = (planes,)
            inpl = planes
    s["fc.weight"] = (num_classes, 512)
    s["fc.bias"] = (num_classes,)
    return s


def bn_names(shapes) -> List[str]:
    """prefixes of every BatchNorm (they carry running_mean / running_var / num_batches_tracked buffers)"""
    return [k[: -len(".weight")] for k in shapes if (".bn" in k or k.startswith("bn") or "downsample.1" in k) and k.endswith(".weight")]


def synth_state(num_classes: int = 10, seed: int = 0, bn_stats: bool = True, depth: int = 18):
    """Deterministic synthetic weights (formula shared by make_golden.py and the tests; no checkpoint exists offline).

    conv: N(0, 2/fan_out) (the reference's kaiming_normal_(fan_out), ResNet.py:247-249), BN gamma 1+0.1 N, beta 0.1 N,
    running_mean 0.1 N, running_var 1+0.1|N|, fc: N(0, 1/512)."""
    g = torch.Generator().manual_seed(seed)
    shapes = resnet18_param_shapes(num_classes, depth)
    params: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shp in shapes.items():
        if len(shp) == 4:
            fan_out = shp[0] * shp[2] * shp[3]
            params[name] = torch.randn(shp, generator=g) * math.sqrt(2.0 / fan_out)
        elif name == "fc.weight":
            params[name] = torch.randn(shp, generator=g) * math.sqrt(1.0 / shp[1])
        elif name.endswith(".weight"):
            params[name] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            params[name] = 0.1 * torch.randn(shp, generator=g)
    buffers: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for bn in bn_names(shapes):
        c = shapes[bn + ".weight"][0]
        if bn_stats:
            buffers[bn + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
            buffers[bn + ".running_var"] = 1.0 + 0.1 * torch.randn(c, generator=g).abs()
        else:
            buffers[bn + ".running_mean"] = torch.zeros(c)
            buffers[bn + ".running_var"] = torch.ones(c)
        buffers[bn + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return params, buffers


def state_dict_of(params, buffers, mean=CIFAR_MEAN, std=CIFAR_STD):
    """state_dict in the reference's key layout (incl. normalize.mean/std buffers)."""
    sd = OrderedDict()
    sd["normalize.mean"] = torch.tensor(mean)
    sd["normalize.std"] = torch.tensor(std)
    sd.update(params)
    sd.update(buffers)
    return sd


# ---------------------------------------------------------------------------------------------
# forward (autograd supplies the backward -- this is the fp32 CPU reference of the op chain)
# ---------------------------------------------------------------------------------------------
def _bn(x, p, b, name, train: bool, momentum: float = 0.1, eps: float = 1e-5):
    rm, rv = b[name + ".running_mean"], b[name + ".running_var"]
    if train:
        b[name + ".num_batches_tracked"] += 1  # nn.BatchNorm2d bookkeeping (buffers are not masked, Appendix B.1)
    return F.batch_norm(x, rm, rv, p[name + ".weight"], p[name + ".bias"], train, momentum, eps)


class _RoundFwd(torch.autograd.Function):
    """value stored in bf16 by the engine (conv operand); its gradient stays fp32"""

    @staticmethod
    def forward(ctx, t):
        return t.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBoth(torch.autograd.Function):
    """tensor whose value AND whose gradient the engine stores in bf16 (raw conv outputs, activations)"""

    @staticmethod
    def forward(ctx, t):
        return t.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def resnet_forward(p: Dict[str, torch.Tensor], b: Dict[str, torch.Tensor], x: torch.Tensor, train: bool,
                   mean=CIFAR_MEAN, std=CIFAR_STD, emulate_bf16: bool = False) -> torch.Tensor:
    """emulate_bf16=False: the reference's fp32 op chain.
    emulate_bf16=True : the SAME chain in fp32 arithmetic, with a bf16 rounding at every point where the sm_100a engine
    stores a tensor in bf16 (DESIGN.md "precision"): conv operands (normalised input, weights), raw conv outputs and
    activations, and -- on the way back -- the gradients of those outputs / activations.  This is the plain-PyTorch
    reference of the op the kernels actually implement; the fp32 chain measures what bf16 storage costs."""
    rf = _RoundFwd.apply if emulate_bf16 else (lambda t: t)
    rb = _RoundBoth.apply if emulate_bf16 else (lambda t: t)
    m = torch.tensor(mean, dtype=x.dtype, device=x.device)[None, :, None, None]
    s = torch.tensor(std, dtype=x.dtype, device=x.device)[None, :, None, None]
    if emulate_bf16:
        x = (x - m) * (1.0 / s)  # the engine multiplies by 1/std
    else:
        x = x.sub(m).div(s)  # ResNet.py:23-28
    x = rf(x)
    w = lambda k: rf(p[k])
    x = rb(F.relu(_bn(rb(F.conv2d(x, w("conv1.weight"), padding=1)), p, b, "bn1", train)))  # :307-310 (maxpool = Identity)
    inpl = 64
    for li, (planes, stride) in enumerate(STAGES, start=1):
        for blk in range(sum(1 for k in p if k.startswith(f"layer{li}.") and k.endswith(".conv1.weight"))):
            pre = f"layer{li}.{blk}."
            st = stride if blk == 0 else 1
            identity = x
            out = rb(F.relu(_bn(rb(F.conv2d(x, w(pre + "conv1.weight"), stride=st, padding=1)), p, b, pre + "bn1", train)))
            out = _bn(rb(F.conv2d(out, w(pre + "conv2.weight"), padding=1)), p, b, pre + "bn2", train)
            if pre + "downsample.0.weight" in p:
                identity = _bn(rb(F.conv2d(x, w(pre + "downsample.0.weight"), stride=st)), p, b, pre + "downsample.1", train)
            x = rb(F.relu(out + identity))  # :121-122
            inpl = planes
    x = F.adaptive_avg_pool2d(x, 1).flatten(1)  # :317-318
    return F.linear(x, p["fc.weight"], p["fc.bias"])  # :320


def loss_and_grads(p, b, x, y, train: bool, sign: float = 1.0, emulate_bf16: bool = False):
    """sign * mean CE and its gradient w.r.t. every parameter (zero_grad + backward of the reference loops)."""
    leaves = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in p.items())
    out = resnet_forward(leaves, b, x, train, emulate_bf16=emulate_bf16)
    loss = sign * F.cross_entropy(out, y)
    grads = torch.autograd.grad(loss, list(leaves.values()))
    return loss.detach(), out.detach(), OrderedDict(zip(leaves.keys(), grads))


# ---------------------------------------------------------------------------------------------
# (i) generate_mask.py:14-82
# ---------------------------------------------------------------------------------------------
THRESHOLDS = [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0]  # generate_mask.py:50


def accumulate_saliency(p, b, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]], emulate_bf16: bool = False):
    """generate_mask.py:25-48: eval mode, loss = -CE, gradients[name] += grad, abs_ at the end. Returns flat |G|."""
    acc = None
    for x, y in batches:
        _, _, g = loss_and_grads(p, b, x, y, train=False, sign=-1.0, emulate_bf16=emulate_bf16)
        flat = torch.cat([t.flatten() for t in g.values()])
        acc = flat if acc is None else acc + flat
    return acc.abs_()


def masks_from_saliency(absg: torch.Tensor, ratios=THRESHOLDS):
    """generate_mask.py:57-80 with stable sorts (see oracle/salun_oracle.c on ties). Returns {ratio: int64 flat mask}."""
    out = {}
    neg = -absg
    positions = torch.argsort(neg, stable=True)
    ranks = torch.argsort(positions, stable=True)
    for r in ratios:
        k = int(len(neg) * r)
        m = torch.zeros_like(ranks)
        m[ranks < k] = 1
        out[r] = m
    return out


def split_mask(flat_mask: torch.Tensor, shapes) -> "OrderedDict[str, torch.Tensor]":
    out, off = OrderedDict(), 0
    for name, shp in shapes.items():
        n = math.prod(shp)
        out[name] = flat_mask[off: off + n].reshape(shp)
        off += n
    return out


# ---------------------------------------------------------------------------------------------
# (ii) masked SGD steps: RL.py:123-176, GA.py:107-128, FT.py:116-144
# ---------------------------------------------------------------------------------------------
class MaskedSGD:
    """torch.optim.SGD(momentum, wd) + mask multiply + restore, restated per coordinate."""

    def __init__(self, p, mask=None, lr=0.013, momentum=0.9, wd=5e-4):
        self.p, self.mask, self.lr, self.mu, self.wd = p, mask, lr, momentum, wd
        self.v = OrderedDict((k, torch.zeros_like(t)) for k, t in p.items())

    def step(self, grads):
        for k, t in self.p.items():
            g = grads[k]
            m = None if self.mask is None else self.mask[k].to(t.dtype)
            if m is not None:
                g = g * m  # RL.py:11-14
            gp = g + self.wd * t
            v = self.mu * self.v[k] + gp
            newp = t - self.lr * v
            if m is not None:  # RL.py:17-34 (theta0 == current value on masked-out coordinates)
                newp = torch.where(m != 0, newp, t)
                v = v * m
            self.v[k] = v
            self.p[k] = newp


def unlearn_step(p, b, opt: MaskedSGD, x, y, sign: float = 1.0, emulate_bf16: bool = False, l1_alpha: float = 0.0):
    """one loop body of RL/GA/FT: train-mode forward, backward, mask, SGD, restore. Returns (loss, logits).
    l1_alpha: FT_l1's  loss += alpha * ||theta||_1  (FT.py:13-17,133-134)  =>  grad += alpha * sign(theta) before the mask."""
    loss, out, g = loss_and_grads(p, b, x, y, train=True, sign=sign, emulate_bf16=emulate_bf16)
    if l1_alpha:
        g = OrderedDict((k, g[k] + l1_alpha * torch.sign(p[k])) for k in p)
        loss = loss + l1_alpha * sum(t.abs().sum() for t in p.values())
    opt.step(g)
    return loss, out


# ---------------------------------------------------------------------------------------------
# Bottleneck nets (resnet50/101/152: models/ResNet.py:127-177, 358-390) with either stem (:217-230)
# ---------------------------------------------------------------------------------------------
BOTTLENECK_BLOCKS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}


def bottleneck_param_shapes(num_classes: int = 1000, depth: int = 50, imagenet: bool = True):
    """named_parameters() order of the reference's Bottleneck ResNets (resnet50: 161 tensors, 25 557 032 for 1000 classes)."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    s["conv1.weight"] = (64, 3, 7, 7) if imagenet else (64, 3, 3, 3)
    s["bn1.weight"] = (64,)
    s["bn1.bias"] = (64,)
    inpl = 64
    for li, ((planes, stride), nblk) in enumerate(zip(STAGES, BOTTLENECK_BLOCKS[depth]), start=1):
        for b in range(nblk):
            pre = f"layer{li}.{b}."
            st = stride if b == 0 else 1
            s[pre + "conv1.weight"] = (planes, inpl, 1, 1)
            s[pre + "bn1.weight"] = (planes,)
            s[pre + "bn1.bias"] = (planes,)
            s[pre + "conv2.weight"] = (planes, planes, 3, 3)
            s[pre + "bn2.weight"] = (planes,)
            s[pre + "bn2.bias"] = (planes,)
            s[pre + "conv3.weight"] = (planes * 4, planes, 1, 1)
            s[pre + "bn3.weight"] = (planes * 4,)
            s[pre + "bn3.bias"] = (planes * 4,)
            if st != 1 or inpl != planes * 4:
                s[pre + "downsample.0.weight"] = (planes * 4, inpl, 1, 1)
                s[pre + "downsample.1.weight"] = (planes * 4,)
                s[pre + "downsample.1.bias"] = (planes * 4,)
            inpl = planes * 4
    s["fc.weight"] = (num_classes, 2048)
    s["fc.bias"] = (num_classes,)
    return s


def synth_state_bottleneck(num_classes: int = 10, seed: int = 0, depth: int = 50, imagenet: bool = True):
    """same weight formula as synth_state, for the Bottleneck tables"""
    g = torch.Generator().manual_seed(seed)
    shapes = bottleneck_param_shapes(num_classes, depth, imagenet)
    params: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shp in shapes.items():
        if len(shp) == 4:
            params[name] = torch.randn(shp, generator=g) * math.sqrt(2.0 / (shp[0] * shp[2] * shp[3]))
        elif name == "fc.weight":
            params[name] = torch.randn(shp, generator=g) * math.sqrt(1.0 / shp[1])
        elif name.endswith(".weight"):
            params[name] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            params[name] = 0.1 * torch.randn(shp, generator=g)
    buffers: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shp in shapes.items():
        if len(shp) == 1 and name.endswith(".weight"):
            bn = name[: -len(".weight")]
            buffers[bn + ".running_mean"] = 0.1 * torch.randn(shp[0], generator=g)
            buffers[bn + ".running_var"] = 1.0 + 0.1 * torch.randn(shp[0], generator=g).abs()
            buffers[bn + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return params, buffers


def bottleneck_forward(p, b, x, train: bool, imagenet: bool = True, mean=CIFAR_MEAN, std=CIFAR_STD,
                       emulate_bf16: bool = False):
    """ResNet._forward_impl (ResNet.py:303-322) with Bottleneck.forward (:157-177); emulate_bf16 as in resnet_forward."""
    rf = _RoundFwd.apply if emulate_bf16 else (lambda t: t)
    rb = _RoundBoth.apply if emulate_bf16 else (lambda t: t)
    m = torch.tensor(mean, dtype=x.dtype, device=x.device)[None, :, None, None]
    s = torch.tensor(std, dtype=x.dtype, device=x.device)[None, :, None, None]
    x = (x - m) * (1.0 / s) if emulate_bf16 else x.sub(m).div(s)
    x = rf(x)
    w = lambda k: rf(p[k])
    if imagenet:
        x = rb(F.relu(_bn(rb(F.conv2d(x, w("conv1.weight"), stride=2, padding=3)), p, b, "bn1", train)))
        x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    else:
        x = rb(F.relu(_bn(rb(F.conv2d(x, w("conv1.weight"), padding=1)), p, b, "bn1", train)))
    for li, (planes, stride) in enumerate(STAGES, start=1):
        nblk = sum(1 for k in p if k.startswith(f"layer{li}.") and k.endswith(".conv1.weight"))
        for blk in range(nblk):
            pre = f"layer{li}.{blk}."
            st = stride if blk == 0 else 1
            identity = x
            out = rb(F.relu(_bn(rb(F.conv2d(x, w(pre + "conv1.weight"))), p, b, pre + "bn1", train)))
            out = rb(F.relu(_bn(rb(F.conv2d(out, w(pre + "conv2.weight"), stride=st, padding=1)), p, b, pre + "bn2", train)))
            out = _bn(rb(F.conv2d(out, w(pre + "conv3.weight"))), p, b, pre + "bn3", train)
            if pre + "downsample.0.weight" in p:
                identity = _bn(rb(F.conv2d(x, w(pre + "downsample.0.weight"), stride=st)), p, b, pre + "downsample.1", train)
            x = rb(F.relu(out + identity))
    x = F.adaptive_avg_pool2d(x, 1).flatten(1)
    return F.linear(x, p["fc.weight"], p["fc.bias"])


def bottleneck_loss_and_grads(p, b, x, y, train: bool, sign: float = 1.0, imagenet: bool = True, emulate_bf16: bool = False):
    leaves = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in p.items())
    out = bottleneck_forward(leaves, b, x, train, imagenet=imagenet, emulate_bf16=emulate_bf16)
    loss = sign * F.cross_entropy(out, y)
    grads = torch.autograd.grad(loss, list(leaves.values()))
    return loss.detach(), out.detach(), OrderedDict(zip(leaves.keys(), grads))
