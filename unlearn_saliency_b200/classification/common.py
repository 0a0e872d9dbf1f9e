from __future__ import annotations

import torch

from ..engine import CIFAR_MEAN, CIFAR_STD, ResNetEngine


def check_criterion(criterion):
    """The engine computes mean cross-entropy (nn.CrossEntropyLoss(), main_forget.py:118) -- refuse anything else loudly."""
    if criterion is None:
        return
    if not isinstance(criterion, torch.nn.CrossEntropyLoss) or criterion.reduction != "mean" \
            or criterion.weight is not None or getattr(criterion, "label_smoothing", 0.0) != 0.0:
        raise ValueError("the sm_100a engine implements nn.CrossEntropyLoss() (mean reduction, no weights) only")


def as_engine(model, args=None, max_batch=None, symmetric: bool = False, precision: str = "bf16") -> ResNetEngine:
    """Accept the reference's nn.Module (models/ResNet.py) and move it onto the engine, or pass an engine through.
    symmetric=True: allocate the arenas as NVLink peer-mapped symmetric memory (fused data-parallel step).
    precision: engine build ("bf16" | "split"); args.precision, when set, overrides the caller's default."""
    if isinstance(model, ResNetEngine):
        return model
    if not isinstance(model, torch.nn.Module):
        raise TypeError(f"cannot run {type(model)} on the sm_100a engine")
    arch = getattr(args, "arch", None) or "resnet18"
    sd = model.state_dict()
    num_classes = sd["fc.weight"].shape[0]
    mean = tuple(float(v) for v in sd.get("normalize.mean", torch.tensor(CIFAR_MEAN)).flatten())
    std = tuple(float(v) for v in sd.get("normalize.std", torch.tensor(CIFAR_STD)).flatten())
    imagenet = bool(getattr(args, "imagenet_arch", False))       # models/ResNet.py:224-230 stem (resnet50/101/152)
    image = int(getattr(args, "input_size", None) or (224 if imagenet else 32))
    mb = int(max_batch or getattr(args, "batch_size", 256) or 256)
    precision = getattr(args, "precision", None) or precision
    eng = ResNetEngine(arch, num_classes, image, max_batch=mb, mean=mean, std=std, symmetric=symmetric, imagenet=imagenet,
                       precision=precision)
    eng.load_state_dict(sd)
    eng.train(model.training)
    eng._source_module = model  # written back by sync_to_module()
    return eng


def sync_to_module(engine: ResNetEngine):
    """Copy the engine's weights / BN buffers back into the nn.Module it was created from (so that the reference's
    save_unlearn_checkpoint / validate / SVC_MIA code keeps working on ``model``)."""
    mod = getattr(engine, "_source_module", None)
    if mod is not None:
        dev = next(mod.parameters()).device
        mod.load_state_dict({k: v.to(dev) for k, v in engine.state_dict().items()})
    return mod


def dist_info():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1
