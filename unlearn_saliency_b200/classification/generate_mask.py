"""Saliency-mask generation on the sm_100a engine -- mirror of Classification/generate_mask.py:14-82.

Same signature and side effect as the reference's ``save_gradient_ratio``: writes
``<args.save_dir>/with_{r}.pt`` for r in 0.1 ... 1.0, each a dict {parameter name: int64 0/1 tensor of the
parameter's shape} saved as CUDA tensors (generate_mask.py:76-82; consumers multiply without .to(), RL.py:14).

What differs is where the work runs:
  model(image); loss.backward()              -> salun_resnet_forward_backward (eval-mode BN, loss = -CE)
  gradients[name] += param.grad  (62 adds)   -> one salun_saliency_accumulate_flat over the flat grad arena
  abs_ ; cat ; argsort ; argsort ; ranks<k   -> salun_topk_mask (3-pass radix select, int64 + packed-bit output)
Data parallel: batches are dealt round-robin to ranks (eval-mode per-sample gradients are independent of the batch
composition), each rank accumulates locally and ONE all-reduce(sum) of the accumulator follows.
"""
from __future__ import annotations

import os

import torch

from ..io import default_saver, save_sidecar
from ..tail import topk_count
from .common import as_engine, check_criterion, dist_info

THRESHOLD_LIST = [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0]  # generate_mask.py:50


def accumulate_saliency(engine, forget_loader) -> torch.Tensor:
    """generate_mask.py:25-48 -> |sum_b grad_b| as a flat arena-layout tensor (all ranks hold the global sum)."""
    rank, world = dist_info()
    engine.eval()  # generate_mask.py:25
    acc = torch.zeros_like(engine.params)
    for i, (image, target) in enumerate(forget_loader):
        if i % world != rank:
            continue
        image = image.to(engine.device, non_blocking=True).float().contiguous()
        target = target.to(engine.device, non_blocking=True).long().contiguous()
        engine.forward_backward(image, target, loss_sign=-1.0, train=False)  # loss = -criterion(...), :36
        engine.ctx.saliency_accumulate_flat(engine.grads, acc)              # gradients[name] += grad, :41-44
    if world > 1:
        torch.distributed.all_reduce(acc)
    engine.ctx.abs_(acc)  # :46-48
    return acc


def masks_for_ratio(engine, absg_torch_order: torch.Tensor, ratio: float):
    """generate_mask.py:57-80 for one ratio.  Returns (hard_dict in the reference format, packed bits, info)."""
    n = absg_torch_order.numel()
    k = topk_count(n, ratio)  # int(len(all_elements) * i), :60
    m64, bits, info = engine.ctx.topk_mask(absg_torch_order, k, want_info=True)
    hard_dict, off = {}, 0
    for name, shp in engine.table.items():
        cnt = 1
        for s in shp:
            cnt *= s
        hard_dict[name] = m64[off: off + cnt].reshape(shp)
        off += cnt
    return hard_dict, bits, info


def masks_for_ratios(engine, absg_torch_order: torch.Tensor, ratios):
    """generate_mask.py:50-80 for the whole threshold_list in ONE select sweep (salun_topk_mask_multi): the saliencies are
    read once per radix pass for all ratios.  Returns [(hard_dict, packed bits, info)] in the order of `ratios`."""
    n = absg_torch_order.numel()
    m64s, bitss, infos = engine.ctx.topk_mask_multi(absg_torch_order, [topk_count(n, r) for r in ratios], want_info=True)
    out = []
    for m64, bits, info in zip(m64s, bitss, infos):
        hard_dict, off = {}, 0
        for name, shp in engine.table.items():
            cnt = 1
            for s in shp:
                cnt *= s
            hard_dict[name] = m64[off: off + cnt].reshape(shp)
            off += cnt
        out.append((hard_dict, bits, info))
    return out


def save_gradient_ratio(data_loaders, model, criterion, args):
    check_criterion(criterion)
    # the saliency pass decides an index set: it runs on the split-precision build (fp32-class products: 50 % mask
    # Jaccard 0.9995 against the fp32 reference at BASELINE size vs 0.987 for bf16 and 0.996 for the reference's own TF32
    # GPU path, tests/test_acceptance_gpu.py) unless args.precision selects the fast build
    engine = as_engine(model, args, precision="split")
    rank, _ = dist_info()
    acc = accumulate_saliency(engine, data_loaders["forget"])
    flat = engine.from_native_flat(acc).contiguous()  # named_parameters order & PyTorch layout, as cat(flatten) :57
    os.makedirs(args.save_dir, exist_ok=True)
    infos = {}
    saver = default_saver()   # the ten 89 MB files are pickled / written on a worker thread
    for r, (hard_dict, bits, info) in zip(THRESHOLD_LIST, masks_for_ratios(engine, flat, THRESHOLD_LIST)):
        infos[r] = info
        if rank == 0:
            path = os.path.join(args.save_dir, "with_{}.pt".format(r))
            saver.save(hard_dict, path, cuda_tensors=True)                  # :82 -- CUDA tensors, like the reference's file
            save_sidecar(saver, path, bits, engine.table, ratio=r)          # 1 bit / parameter, same order and layout
    saver.wait()
    return infos
