"""The reference's unlearning loops in STOCK PyTorch (eager, cuDNN / cuBLAS) on the same B200 -- the denominator of the
north-star "10x the reference PyTorch-eager path" target (SURVEY.md section 8d last bullet).

Nothing here is on the product path and nothing of the product is on this path: torch.nn modules, torch.optim, autograd.
`/root/reference` does not exist on the GPU box, so the reference's statements are restated (with their line numbers):

  ResNet-18 / CIFAR-10, masked RL step      Classification/models/ResNet.py:58-124,180-322 (BasicBlock net, CIFAR stem,
                                            normalize layer), unlearn/RL.py:11-34 (_apply_mask_to_grads,
                                            _restore_masked_params), :123-140 (loop body), unlearn/impl.py:68-73 (SGD)
  DDPM saliency_unlearn iteration           oracle/ddpm.py:saliency_unlearn_step -- the statements of
                                            DDPM/runners/diffusion.py:519-593 around oracle/unet.py (the PyTorch
                                            restatement of Conditional_Model pinned to the reference's outputs); the int64
                                            mask stays on the CPU and is uploaded every step exactly like :589-592

Two arithmetic modes: "tf32" = torch defaults (cuDNN convolutions in TF32, matmuls fp32: what the reference runs on an
Ampere+ GPU) and "fp32" (torch.backends.cudnn.allow_tf32 = False).
"""
from __future__ import annotations

import time
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

CIFAR_MEAN = (0.4914, 0.4822, 0.4465)
CIFAR_STD = (0.2470, 0.2435, 0.2616)


class _Normalize(nn.Module):                       # ResNet.py:12-28
    def __init__(self, mean, std):
        super().__init__()
        self.register_buffer("mean", torch.tensor(mean))
        self.register_buffer("std", torch.tensor(std))

    def forward(self, x):
        return (x - self.mean[None, :, None, None]) / self.std[None, :, None, None]


class _BasicBlock(nn.Module):                      # ResNet.py:77-124
    def __init__(self, inplanes, planes, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = None
        if stride != 1 or inplanes != planes:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))

    def forward(self, x):
        identity = x if self.downsample is None else self.downsample(x)
        out = F.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return F.relu(out + identity)


class ResNet18(nn.Module):                         # ResNet.py:180-322 with imagenet=False (3x3 stem, no max pool)
    def __init__(self, num_classes=10):
        super().__init__()
        self.normalize = _Normalize(CIFAR_MEAN, CIFAR_STD)
        self.conv1 = nn.Conv2d(3, 64, 3, 1, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        layers, inpl = [], 64
        for planes, stride in ((64, 1), (128, 2), (256, 2), (512, 2)):
            layers.append(nn.Sequential(_BasicBlock(inpl, planes, stride), _BasicBlock(planes, planes, 1)))
            inpl = planes
        self.layer1, self.layer2, self.layer3, self.layer4 = layers
        self.fc = nn.Linear(512, num_classes)

    def forward(self, x):
        x = F.relu(self.bn1(self.conv1(self.normalize(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.fc(torch.flatten(F.adaptive_avg_pool2d(x, 1), 1))


def _apply_mask_to_grads(model, mask):             # RL.py:11-14
    for name, param in model.named_parameters():
        if param.grad is not None:
            param.grad *= mask[name]


def _restore_masked_params(model, mask, theta0, optimizer):   # RL.py:17-34
    with torch.no_grad():
        for name, param in model.named_parameters():
            if name not in mask:
                continue
            mask_tensor = mask[name].to(device=param.device, dtype=param.dtype)
            inv_mask_tensor = 1 - mask_tensor
            if torch.count_nonzero(inv_mask_tensor) == 0:
                continue
            param.data.mul_(mask_tensor).add_(theta0[name].to(param.device) * inv_mask_tensor)
            state = optimizer.state.get(param, None)
            if state is not None and "momentum_buffer" in state:
                state["momentum_buffer"].mul_(mask_tensor)


def set_mode(mode: str):
    assert mode in ("tf32", "fp32")
    torch.backends.cudnn.allow_tf32 = mode == "tf32"
    torch.backends.cuda.matmul.allow_tf32 = False


def resnet18_rl_steps_per_sec(batch=256, steps=20, warmup=5, mode="tf32", device="cuda", host_inputs=True):
    """One step = the loop body of RL.py:123-140 (random labels, .cuda() of the batch, forward, CE, backward, mask, SGD
    step, restore) with a 50 % int64 CUDA mask.  Returns steps/s (CUDA events, inputs from pinned host memory as the
    reference's loader delivers them)."""
    set_mode(mode)
    torch.manual_seed(0)
    model = ResNet18(10).to(device).train()
    crit = nn.CrossEntropyLoss()
    opt = torch.optim.SGD(model.parameters(), 0.013, momentum=0.9, weight_decay=5e-4)          # impl.py:68-73
    g = torch.Generator().manual_seed(1)
    mask = OrderedDict((n, (torch.rand(p.shape, generator=g) < 0.5).to(torch.int64).to(device)) for n, p in model.named_parameters())
    theta0 = {n: p.detach().clone() for n, p in model.named_parameters()}                        # RL.py:42-49
    pool = [(torch.rand(batch, 3, 32, 32, generator=g).pin_memory(), torch.randint(0, 10, (batch,), generator=g).pin_memory())
            for _ in range(4)]
    dev_pool = [(x.to(device), y.to(device)) for x, y in pool]
    top1 = 0.0

    def step(i):
        nonlocal top1
        image, target = (pool if host_inputs else dev_pool)[i % 4]
        target = torch.randint(0, 10, target.shape)                                               # RL.py:125
        image, target = image.to(device), target.to(device)                                        # :126-127 (.cuda())
        output = model(image)
        loss = crit(output, target)
        opt.zero_grad()
        loss.backward()
        _apply_mask_to_grads(model, mask)
        opt.step()
        _restore_masked_params(model, mask, theta0, opt)
        top1 = float(loss.item())                                                                  # losses.update(loss.item(), ...) :166

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    set_mode("fp32")
    return 1000.0 / ms, ms


def ddpm_unlearn_its_per_sec(batch=128, steps=5, warmup=2, mode="tf32", device="cuda"):
    """One iteration of Diffusion.saliency_unlearn (rl, alpha 1e-3, 50 % mask held on the CPU and uploaded per step like
    runners/diffusion.py:589-592, clip 1.0, Adam 1e-4) on the cifar10 U-Net with `batch` remain + `batch` forget images."""
    from oracle import ddpm as OD                # the restated statements + the pinned torch network (module docstring)
    from oracle.unet import ConditionalUNet
    from unlearn_saliency_b200.diffusion.config import cifar10_config
    from unlearn_saliency_b200.diffusion.runner import get_beta_schedule
    set_mode(mode)
    torch.manual_seed(0)
    model = ConditionalUNet(cifar10_config()).to(device)
    g = torch.Generator().manual_seed(1)
    mask = {k: (torch.rand(p.shape, generator=g) < 0.5).to(torch.int64) for k, p in model.named_parameters()}   # CPU
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    betas = torch.from_numpy(get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)).float().to(device)
    n, S = batch, 32

    def draw():
        return dict(x_r=torch.rand(n, 3, S, S, generator=g), c_r=torch.randint(1, 10, (n,), generator=g),
                    x_f=torch.rand(n, 3, S, S, generator=g), c_f=torch.zeros(n, dtype=torch.long),
                    t_r=torch.randint(0, 1000, (n,), generator=g), e_r=torch.randn(n, 3, S, S, generator=g),
                    t_f=torch.randint(0, 1000, (n,), generator=g), e_f=torch.randn(n, 3, S, S, generator=g),
                    drop_r=torch.rand(n, generator=g) < 0.1, drop_f=torch.rand(n, generator=g) < 0.1,
                    drop_p=torch.rand(n, generator=g) < 0.1)

    pool = [draw() for _ in range(2)]
    for i in range(warmup):
        OD.saliency_unlearn_step(model, opt, mask, pool[i % 2], betas, alpha=1e-3, method="rl", keep_raw=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        loss, _, _ = OD.saliency_unlearn_step(model, opt, mask, pool[i % 2], betas, alpha=1e-3, method="rl", keep_raw=False)
        float(loss.item())                        # logging.info(f"... loss: {loss.item()}") :595
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    set_mode("fp32")
    del model, opt
    torch.cuda.empty_cache()
    return 1000.0 / ms, ms
