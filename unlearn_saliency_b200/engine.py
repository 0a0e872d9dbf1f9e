"""Host-side handle of the sm_100a ResNet engine (include/salun.h: salun_resnet_*).

Mirrors what the reference's loops do with ``model`` / ``optimizer``:
  model(image); loss.backward()           -> ResNetEngine.forward_backward      (RL.py:128-132, generate_mask.py:35-39)
  _apply_mask_to_grads + optimizer.step()
      + _restore_masked_params            -> MaskedSGD.step                     (RL.py:134-140, impl.py:68-73)
  model.state_dict() / load_state_dict    -> ResNetEngine.state_dict / load_state_dict (reference key layout)

PyTorch owns the device buffers (flat fp32 arenas) and the stream; all arithmetic is in libsalun.so.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import check
from .tail import SalunContext, _ptr, _stream, mask_words

CIFAR_MEAN = (0.4914, 0.4822, 0.4465)  # Classification/models/ResNet.py:214-216
CIFAR_STD = (0.2470, 0.2435, 0.2616)
_STAGE_BLOCKS = {18: (2, 2, 2, 2), 34: (3, 4, 6, 3), 50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}
_ARCH_DEPTH = {"resnet18": 18, "resnet34": 34, "resnet50": 50, "resnet101": 101, "resnet152": 152}


def _bottleneck_param_table(depth: int, num_classes: int, imagenet: bool) -> "OrderedDict[str, Tuple[int, ...]]":
    """Bottleneck ResNets (ResNet.py:127-177): conv1 1x1, conv2 3x3 (carries the stride), conv3 1x1 x4, projection shortcut."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    s["conv1.weight"] = (64, 3, 7, 7) if imagenet else (64, 3, 3, 3)
    s["bn1.weight"] = (64,)
    s["bn1.bias"] = (64,)
    inpl = 64
    for li, (planes, nblk) in enumerate(zip((64, 128, 256, 512), _STAGE_BLOCKS[depth]), start=1):
        for b in range(nblk):
            pre = f"layer{li}.{b}."
            stride = 2 if (b == 0 and li > 1) else 1
            for nm, shp in (("conv1", (planes, inpl, 1, 1)), ("conv2", (planes, planes, 3, 3)), ("conv3", (planes * 4, planes, 1, 1))):
                s[pre + nm + ".weight"] = shp
                bn = "bn" + nm[-1]
                s[pre + bn + ".weight"] = (shp[0],)
                s[pre + bn + ".bias"] = (shp[0],)
            if stride != 1 or inpl != planes * 4:
                s[pre + "downsample.0.weight"] = (planes * 4, inpl, 1, 1)
                s[pre + "downsample.1.weight"] = (planes * 4,)
                s[pre + "downsample.1.bias"] = (planes * 4,)
            inpl = planes * 4
    s["fc.weight"] = (num_classes, 2048)
    s["fc.bias"] = (num_classes,)
    return s


def resnet_param_table(depth: int, num_classes: int, imagenet: bool = False) -> "OrderedDict[str, Tuple[int, ...]]":
    """named_parameters() order and PyTorch shapes of the reference's ResNets (ResNet.py:180-260)."""
    if depth >= 50:
        return _bottleneck_param_table(depth, num_classes, imagenet)
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    s["conv1.weight"] = (64, 3, 3, 3)
    s["bn1.weight"] = (64,)
    s["bn1.bias"] = (64,)
    inpl = 64
    for li, (planes, nblk) in enumerate(zip((64, 128, 256, 512), _STAGE_BLOCKS[depth]), start=1):
        for b in range(nblk):
            pre = f"layer{li}.{b}."
            stride = 2 if (b == 0 and li > 1) else 1
            s[pre + "conv1.weight"] = (planes, inpl, 3, 3)
            s[pre + "bn1.weight"] = (planes,)
            s[pre + "bn1.bias"] = (planes,)
            s[pre + "conv2.weight"] = (planes, planes, 3, 3)
            s[pre + "bn2.weight"] = (planes,)
            s[pre + "bn2.bias"] = (planes,)
            if stride != 1 or inpl != planes:
                s[pre + "downsample.0.weight"] = (planes, inpl, 1, 1)
                s[pre + "downsample.1.weight"] = (planes,)
                s[pre + "downsample.1.bias"] = (planes,)
            inpl = planes
    s["fc.weight"] = (num_classes, 512)
    s["fc.bias"] = (num_classes,)
    return s


def _bn_prefixes(table) -> list:
    return [k[: -len(".weight")] for k, shp in table.items()
            if k.endswith(".weight") and len(shp) == 1]


class ResNetEngine:
    """The reference's ``model`` for the hot path: parameters live in one flat fp32 arena on the GPU."""

    def __init__(self, arch: str = "resnet18", num_classes: int = 10, image_size: int = 32, max_batch: int = 256,
                 mean=CIFAR_MEAN, std=CIFAR_STD, device=None, ctx: Optional[SalunContext] = None,
                 symmetric: bool = False, imagenet: bool = False, precision: str = "bf16"):
        """symmetric=True allocates the parameter / gradient arenas as torch symmetric memory (NVLink peer-mapped), which
        DistMaskedSGD needs for its fused reduce-scatter + update + all-gather kernel.
        precision: "bf16" (activations and tensor-core operands in bf16) or "split" (bf16 hi/lo pairs, fp32-class
        products; the mode the saliency-mask pass uses to reproduce the fp32 reference's index set)."""
        if arch not in _ARCH_DEPTH:
            raise ValueError(f"arch {arch!r} is not served by the sm_100a engine (supported: {sorted(_ARCH_DEPTH)})")
        self.arch, self.depth = arch, _ARCH_DEPTH[arch]
        if imagenet and self.depth < 50:
            raise ValueError("the ImageNet stem (imagenet=True, ResNet.py:224-230) is served for resnet50/101/152")
        self.imagenet = bool(imagenet)
        self.num_classes, self.image_size, self.max_batch = num_classes, image_size, max_batch
        self.ctx = ctx if ctx is not None else SalunContext(device)
        self.device = self.ctx.device
        self.precision = precision
        self._lib = _lib.lib(precision)
        self.mean, self.std = tuple(float(v) for v in mean), tuple(float(v) for v in std)
        self.cfg = _lib.salun_resnet_cfg(self.depth, num_classes, image_size, max_batch, (C.c_float * 3)(*self.mean),
                                         (C.c_float * 3)(*self.std), 1e-5, 0.1, 1 if imagenet else 0)
        self.table = resnet_param_table(self.depth, num_classes, self.imagenet)
        self.n_params = int(self._lib.salun_resnet_param_count(C.byref(self.cfg)))
        n_bn = int(self._lib.salun_resnet_bn_channels(C.byref(self.cfg)))
        if self.n_params != sum(math.prod(s) for s in self.table.values()):
            raise RuntimeError("parameter table of the host mirror and libsalun disagree")
        dev = self.device
        self.symmetric = symmetric
        if symmetric:
            import torch.distributed._symmetric_memory as symm_mem
            self.params = symm_mem.empty(self.n_params, dtype=torch.float32, device=dev)
            self.grads = symm_mem.empty(self.n_params, dtype=torch.float32, device=dev)
            self.params.zero_()
            self.grads.zero_()
        else:
            self.params = torch.zeros(self.n_params, device=dev)
            self.grads = torch.zeros(self.n_params, device=dev)
        self.running_mean = torch.zeros(n_bn, device=dev)
        self.running_var = torch.ones(n_bn, device=dev)
        self.num_batches_tracked = 0
        self.offsets: Dict[str, int] = {}
        off = 0
        for k, shp in self.table.items():
            self.offsets[k] = off
            off += math.prod(shp)
        self.bn_prefixes = _bn_prefixes(self.table)
        self.bn_offsets: Dict[str, int] = {}
        off = 0
        for p in self.bn_prefixes:
            self.bn_offsets[p] = off
            off += self.table[p + ".weight"][0]
        assert off == n_bn
        h = C.c_void_p()
        check(self._lib.salun_resnet_create(self.ctx.handle, C.byref(self.cfg), _ptr(self.params), _ptr(self.grads),
                                            _ptr(self.running_mean), _ptr(self.running_var), C.byref(h)),
              "salun_resnet_create")
        self._h = h
        self._loss = torch.zeros(1, device=dev)
        self.training = True

    def enable_sync_bn(self, group=None):
        """Train-mode BatchNorm over the GLOBAL mini-batch when the batch is sharded across ranks (sync-BN): the per-BN
        sums are exchanged through NVLink peer-mapped symmetric memory inside the step (salun_resnet_enable_syncbn), so a
        data-parallel step computes what the reference's single-process step computes on the concatenated batch.
        torch.distributed (NCCL) must be initialised; every rank calls this and then runs the same sequence of steps."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        n = int(self._lib.salun_resnet_syncbn_doubles(C.byref(self.cfg)))
        if n <= 0:
            raise RuntimeError("sync-BN is served for resnet18 / resnet34")
        self._sb_sums = symm_mem.empty(n, dtype=torch.float64, device=self.device)
        self._sb_flags = symm_mem.empty(max(world, 8), dtype=torch.int64, device=self.device)
        self._sb_sums.zero_()
        self._sb_flags.zero_()
        self._sb_hs = symm_mem.rendezvous(self._sb_sums, group)
        self._sb_hf = symm_mem.rendezvous(self._sb_flags, group)
        torch.cuda.synchronize(self.device)
        dist.barrier(group)                      # every rank's buffers are zeroed before anyone publishes an epoch
        sums = (C.c_void_p * world)(*[int(p) for p in self._sb_hs.buffer_ptrs])
        flags = (C.c_void_p * world)(*[int(p) for p in self._sb_hf.buffer_ptrs])
        check(self._lib.salun_resnet_enable_syncbn(self._h, sums, flags, rank, world), "salun_resnet_enable_syncbn")
        self.sync_bn = True
        return self

    def close(self):
        if getattr(self, "_h", None):
            self._lib.salun_resnet_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- nn.Module-like surface used by the reference loops ------------------------------
    def train(self, mode: bool = True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def cuda(self, *a, **k):
        return self

    def named_parameters(self):
        """(name, tensor in PyTorch layout) -- copies for conv weights (the arena is OHWI)."""
        for k in self.table:
            yield k, self.get_param(k)

    # ---- layout conversion between the reference (OIHW) and the arena (OHWI) --------------
    def to_native(self, flat_or_dict) -> torch.Tensor:
        """flat tensor / {name: tensor} in named_parameters order & PyTorch layout -> arena-layout flat tensor."""
        out, off = None, 0
        is_dict = isinstance(flat_or_dict, dict)
        for k, shp in self.table.items():
            n = math.prod(shp)
            t = flat_or_dict[k].reshape(shp) if is_dict else flat_or_dict[off: off + n].reshape(shp)
            if out is None:
                out = torch.empty(self.n_params, dtype=t.dtype, device=self.device)
            t = t.to(self.device)
            if len(shp) == 4:
                t = t.permute(0, 2, 3, 1)
            out[off: off + n] = t.reshape(-1)
            off += n
        return out

    def from_native(self, flat: torch.Tensor) -> "OrderedDict[str, torch.Tensor]":
        """arena-layout flat tensor -> {name: tensor in PyTorch layout} (contiguous copies)."""
        res, off = OrderedDict(), 0
        for k, shp in self.table.items():
            n = math.prod(shp)
            t = flat[off: off + n]
            if len(shp) == 4:
                t = t.reshape(shp[0], shp[2], shp[3], shp[1]).permute(0, 3, 1, 2)
            res[k] = t.reshape(shp).contiguous()
            off += n
        return res

    def from_native_flat(self, flat: torch.Tensor) -> torch.Tensor:
        return torch.cat([t.reshape(-1) for t in self.from_native(flat).values()])

    def get_param(self, name: str) -> torch.Tensor:
        shp = self.table[name]
        t = self.params[self.offsets[name]: self.offsets[name] + math.prod(shp)]
        if len(shp) == 4:
            return t.reshape(shp[0], shp[2], shp[3], shp[1]).permute(0, 3, 1, 2).contiguous()
        return t.reshape(shp).clone()

    def grad_dict(self) -> "OrderedDict[str, torch.Tensor]":
        return self.from_native(self.grads)

    def load_state_dict(self, sd, strict: bool = False):
        """Reference checkpoint layout: plain names, normalize.mean/std buffers."""
        sd = {k[7:] if k.startswith("module.") else k: v for k, v in sd.items()}
        missing = [k for k in self.table if k not in sd]
        if missing and strict:
            raise KeyError(f"missing keys: {missing[:5]}...")
        with torch.no_grad():
            self.params.copy_(self.to_native({k: (sd[k].float() if k in sd else self.get_param(k)) for k in self.table}))
            for p in self.bn_prefixes:
                o, c = self.bn_offsets[p], self.table[p + ".weight"][0]
                if p + ".running_mean" in sd:
                    self.running_mean[o: o + c] = sd[p + ".running_mean"].to(self.device)
                    self.running_var[o: o + c] = sd[p + ".running_var"].to(self.device)
            if "bn1.num_batches_tracked" in sd:
                self.num_batches_tracked = int(sd["bn1.num_batches_tracked"])
            if "normalize.mean" in sd and tuple(round(float(v), 6) for v in sd["normalize.mean"]) != tuple(
                    round(v, 6) for v in self.mean):
                raise ValueError("checkpoint normalize.mean differs from the engine's configured mean; "
                                 "construct the engine with the dataset's statistics (utils.py:115-117)")
        return self

    def state_dict(self) -> "OrderedDict[str, torch.Tensor]":
        sd = OrderedDict()
        sd["normalize.mean"] = torch.tensor(self.mean, device=self.device)
        sd["normalize.std"] = torch.tensor(self.std, device=self.device)
        params = self.from_native(self.params)
        bn = set(self.bn_prefixes)
        for k, v in params.items():
            sd[k] = v
            pre = k[: -len(".bias")] if k.endswith(".bias") else None
            if pre in bn:  # nn.BatchNorm2d order: weight, bias, running_mean, running_var, num_batches_tracked
                o, c = self.bn_offsets[pre], self.table[pre + ".weight"][0]
                sd[pre + ".running_mean"] = self.running_mean[o: o + c].clone()
                sd[pre + ".running_var"] = self.running_var[o: o + c].clone()
                sd[pre + ".num_batches_tracked"] = torch.tensor(self.num_batches_tracked, device=self.device)
        return sd

    # ---- compute --------------------------------------------------------------------------
    def _check_x(self, x):
        if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4 and x.shape[1] == 3
                and x.shape[2] == x.shape[3] == self.image_size):
            raise ValueError(f"x must be a contiguous CUDA fp32 tensor [n,3,{self.image_size},{self.image_size}]")
        if not 0 < x.shape[0] <= self.max_batch:
            raise ValueError(f"batch {x.shape[0]} exceeds max_batch {self.max_batch}")

    def forward_backward(self, x: torch.Tensor, y: torch.Tensor, loss_sign: float = 1.0, want_logits: bool = False,
                         train: Optional[bool] = None):
        """loss = loss_sign * CE(model(x), y); grads <- d loss / d params.  Returns (loss[1] device tensor, logits|None)."""
        self._check_x(x)
        if not (y.is_cuda and y.dtype == torch.int64 and y.is_contiguous() and y.numel() == x.shape[0]):
            raise ValueError("y must be a contiguous CUDA int64 tensor [n]")
        train = self.training if train is None else train
        logits = torch.empty(x.shape[0], self.num_classes, device=self.device) if want_logits else None
        check(self._lib.salun_resnet_forward_backward(self._h, _ptr(x), _ptr(y), x.shape[0], 1 if train else 0,
                                                      float(loss_sign), _ptr(self._loss), _ptr(logits),
                                                      _stream(self.device)), "salun_resnet_forward_backward")
        if train:
            self.num_batches_tracked += 1
        return self._loss, logits

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """eval-mode logits (trainer/val.py validate)"""
        self._check_x(x)
        logits = torch.empty(x.shape[0], self.num_classes, device=self.device)
        check(self._lib.salun_resnet_forward(self._h, _ptr(x), x.shape[0], _ptr(logits), _stream(self.device)),
              "salun_resnet_forward")
        return logits

    __call__ = forward

    # ---- masks ------------------------------------------------------------------------------
    def mask_bits_from_dict(self, mask: Dict[str, torch.Tensor]) -> torch.Tensor:
        """reference mask file {name: int64 0/1 tensor} (generate_mask.py:76-82) -> packed bits in arena layout"""
        native = self.to_native({k: mask[k].to(torch.int64) for k in self.table})
        return self.ctx.pack_mask(native.contiguous())

    def mask_bits_from_file(self, path: str) -> torch.Tensor:
        """mask file of the reference (``with_<r>.pt``) or its packed side-car (``with_<r>.pt.bits``, io.py) -> packed bits
        in arena layout.  The side-car is 64x smaller than the int64 dict and is preferred when present."""
        from .io import load_mask
        bits_torch_order = load_mask(path, self.table, self.ctx, self.device)
        m64 = self.ctx.unpack_mask(bits_torch_order, self.n_params)
        return self.ctx.pack_mask(self.to_native(m64).contiguous())

    def mask_dict_from_native_i64(self, flat_i64: torch.Tensor) -> "OrderedDict[str, torch.Tensor]":
        return self.from_native(flat_i64)


class MaskedSGD:
    """torch.optim.SGD(momentum, weight_decay) + mask multiply + restore, fused (impl.py:68-73, RL.py:11-34)."""

    def __init__(self, engine: ResNetEngine, lr: float, momentum: float = 0.9, weight_decay: float = 5e-4,
                 mask_bits: Optional[torch.Tensor] = None):
        self.engine, self.ctx = engine, engine.ctx
        self.param_groups = [{"lr": lr, "momentum": momentum, "weight_decay": weight_decay}]
        self.momentum_buffer = torch.zeros_like(engine.params)
        self.mask_bits = mask_bits
        if mask_bits is not None and mask_bits.numel() != mask_words(engine.n_params):
            raise ValueError("mask_bits has the wrong length")

    def zero_grad(self):  # grads are overwritten by every forward_backward
        pass

    def step(self):
        g = self.param_groups[0]
        self.ctx.masked_sgd_step(self.engine.params, self.engine.grads, self.momentum_buffer, self.mask_bits,
                                 g["lr"], g["momentum"], g["weight_decay"])


class GraphedStep:
    """One SalUn step -- forward_backward + optimizer step (RL.py:128-140) -- captured ONCE into a CUDA graph for a fixed
    batch size and replayed: the ~175 dependent 3-35 us launches of a ResNet-18 step are launch-latency bound, and a graph
    replay removes the per-launch host cost and most of the inter-kernel gaps.  Inputs are copied into static buffers;
    the engine's wgrad side stream is captured through its fork / join events.  The two warm-up steps needed before
    capture run on a snapshot: parameters, momentum and BatchNorm buffers are restored afterwards.
    With a DistMaskedSGD optimizer the two symmetric-memory barriers and the fused reduce-scatter + SGD + all-gather kernel
    are part of the graph (every rank must construct and replay it the same number of times: the barriers pair up)."""

    def __init__(self, engine: "ResNetEngine", opt, batch: int, loss_sign: float = 1.0, train: Optional[bool] = None):
        self.engine, self.opt = engine, opt
        dev, S = engine.device, engine.image_size
        self.x = torch.zeros(batch, 3, S, S, device=dev)
        self.y = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.train = engine.training if train is None else train
        snap = [t.clone() for t in (engine.params, engine.running_mean, engine.running_var)]
        mom = getattr(opt, "momentum_buffer", None)
        if mom is None:
            mom = getattr(opt, "momentum_shard", None)     # DistMaskedSGD: this rank's shard of the momentum
        mom_snap = mom.clone() if mom is not None else None
        nbt = engine.num_batches_tracked
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(2):
                engine.forward_backward(self.x, self.y, loss_sign=loss_sign, train=self.train)
                opt.step()
        cur.wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            engine.forward_backward(self.x, self.y, loss_sign=loss_sign, train=self.train)
            opt.step()
        with torch.no_grad():
            for dst, src in zip((engine.params, engine.running_mean, engine.running_var), snap):
                dst.copy_(src)
            if mom is not None:
                mom.copy_(mom_snap)
        engine.num_batches_tracked = nbt

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """x, y: host (pinned) or device tensors of the captured batch size.  Returns the device loss scalar."""
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        if self.train:
            self.engine.num_batches_tracked += 1
        return self.engine._loss

    def replay(self) -> torch.Tensor:
        """replay on whatever the static buffers self.x / self.y hold (inputs already resident)"""
        self.graph.replay()
        if self.train:
            self.engine.num_batches_tracked += 1
        return self.engine._loss


class DistMaskedSGD:
    """Data-parallel MaskedSGD: all_reduce(grad)/W + mask + SGD + restore as ONE kernel over NVLink peer memory
    (salun_dp_masked_sgd_step): each rank reduces its 1/W shard of every peer's gradient arena, updates that shard of the
    weights (momentum exists only for the shard) and stores the new weights into every peer's parameter arena.
    The engine must have been created with symmetric=True and torch.distributed (NCCL) must be initialised."""

    def __init__(self, engine: ResNetEngine, lr: float, momentum: float = 0.9, weight_decay: float = 5e-4,
                 mask_bits: Optional[torch.Tensor] = None, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        if not engine.symmetric:
            raise ValueError("DistMaskedSGD needs ResNetEngine(symmetric=True)")
        self.engine, self.ctx = engine, engine.ctx
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.param_groups = [{"lr": lr, "momentum": momentum, "weight_decay": weight_decay}]
        self.mask_bits = mask_bits
        self._hp = symm_mem.rendezvous(engine.params, self.group)
        self._hg = symm_mem.rendezvous(engine.grads, self.group)
        lo, hi = C.c_int64(), C.c_int64()
        check(engine._lib.salun_dp_shard(engine.n_params, self.rank, self.world, C.byref(lo), C.byref(hi)), "salun_dp_shard")
        self.lo, self.hi = lo.value, hi.value
        self.momentum_shard = torch.zeros(max(4, self.hi - self.lo), device=engine.device)
        self._pp = (C.c_void_p * self.world)(*[int(p) for p in self._hp.buffer_ptrs])
        self._gp = (C.c_void_p * self.world)(*[int(p) for p in self._hg.buffer_ptrs])

    def zero_grad(self):
        pass

    def step(self):
        g = self.param_groups[0]
        self._hg.barrier(channel=0)  # every rank's backward has written its gradient arena
        check(self.engine._lib.salun_dp_masked_sgd_step(
            self.ctx.handle, self._pp, self._gp, _ptr(self.momentum_shard), _ptr(self.mask_bits), self.engine.n_params,
            self.rank, self.world, float(g["lr"]), float(g["momentum"]), float(g["weight_decay"]),
            _stream(self.engine.device)), "salun_dp_masked_sgd_step")
        self._hp.barrier(channel=1)  # every rank's shard of the new weights has landed in every replica
