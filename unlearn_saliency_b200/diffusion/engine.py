"""Host-side handle of the sm_100a DDPM U-Net engine (include/salun.h: salun_unet_*).

Mirrors what the reference's DDPM loops do with ``model``:
  model(x_t, t.float(), c, mode="train", ...)                -> UNetEngine.forward           (runners/diffusion.py:533-570)
  model(x_t, t.float(), c, cond_scale=s, mode="test")        -> UNetEngine.forward_cfg       (:974-979, models/diffusion.py:340-355)
  loss.backward()                                            -> UNetEngine.backward(d_eps)   (:579-580, :983)
  model.state_dict() / load_state_dict (``module.`` prefix)  -> UNetEngine.state_dict / load_state_dict

PyTorch owns the flat fp32 parameter / gradient arenas and the stream; all arithmetic of the network is in libsalun.so.
The attribute surface (params, grads, n, offsets, shapes, ctx, flat_from_dict, dict_from_flat, zero_grad) is the one
``flat.FlatMaskedAdam`` / ``flat.FlatSaliency`` expect, so the fused clip + mask + Adam and accumulate + top-k tails run
on the engine's arenas directly.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict
from typing import Dict, Optional

import torch

from .. import _lib
from .._lib import check
from ..tail import SalunContext, _ptr, _stream


class salun_unet_cfg(C.Structure):
    _fields_ = [
        ("ch", C.c_int), ("n_levels", C.c_int), ("ch_mult", C.c_int * 8), ("num_res_blocks", C.c_int),
        ("n_attn_res", C.c_int), ("attn_res", C.c_int * 8), ("image_size", C.c_int), ("in_channels", C.c_int),
        ("out_ch", C.c_int), ("n_classes", C.c_int), ("max_batch", C.c_int), ("dropout", C.c_float),
    ]


_P = C.c_void_p
_lib.register_signatures({
    "salun_unet_param_count": [_P],
    "salun_unet_create": [_P, _P, _P, _P, C.POINTER(_P)],
    "salun_unet_destroy": [_P],
    "salun_unet_forward": [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_uint64, C.c_int, _P, _P],
    "salun_unet_backward": [_P, _P, C.c_int, _P],
    "salun_unet_num_tensors": [_P],
    "salun_unet_tensor_info": [_P, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "salun_unet_export_tensor": [_P, C.c_int, C.c_int, _P, _P],
    "salun_ddpm_q_sample": [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P],
    "salun_ddpm_eps_loss_grad": [_P, _P, _P, _P, C.c_int, C.c_int, _P, _P, _P, _P],
}, {"salun_unet_param_count": C.c_int64})


class DDPMLoss:
    """q-sample and the eps-prediction losses + dL/d(eps) of one iteration on the sm_100a kernels
    (salun_ddpm_q_sample, salun_ddpm_eps_loss_grad): functions/losses.py:21-37, runners/diffusion.py:533-580."""

    def __init__(self, betas: torch.Tensor, ctx: SalunContext):
        self.ctx, self._lib = ctx, _lib.lib()
        dev = ctx.device
        abar = (1 - betas.float().to(dev)).cumprod(dim=0)       # the reference's own fp32 statements (losses.py:31)
        self.sqrt_abar = abar.sqrt().contiguous()
        self.sqrt_1m_abar = (1.0 - abar).sqrt().contiguous()
        self.device = dev

    def q_sample(self, x01: torch.Tensor, e: torch.Tensor, t: torch.Tensor, rescale: bool = True) -> torch.Tensor:
        x01, e, t = x01.contiguous(), e.contiguous(), t.contiguous()
        if not (x01.is_cuda and x01.dtype == torch.float32 and e.dtype == torch.float32 and t.dtype == torch.int64
                and e.shape == x01.shape and t.numel() == x01.shape[0]):
            raise ValueError("q_sample: x01 / e must be CUDA fp32 of the same shape, t CUDA int64 [n]")
        out = torch.empty_like(x01)
        n = x01.shape[0]
        check(self._lib.salun_ddpm_q_sample(self.ctx.handle, _ptr(x01), _ptr(e), _ptr(t), _ptr(self.sqrt_abar),
                                            _ptr(self.sqrt_1m_abar), int(self.sqrt_abar.numel()), 1 if rescale else 0, n,
                                            x01.numel() // max(n, 1),
                                            _ptr(out), _stream(self.device)), "salun_ddpm_q_sample")
        return out

    def loss_grad(self, eps: torch.Tensor, target: torch.Tensor, w: torch.Tensor):
        """returns (loss [1], d_eps like eps, per-sample sum of squares [n])"""
        eps, target, w = eps.contiguous(), target.contiguous(), w.contiguous()
        n = eps.shape[0]
        if not (eps.is_cuda and eps.dtype == torch.float32 and target.shape == eps.shape and target.dtype == torch.float32
                and w.dtype == torch.float32 and w.numel() == n):
            raise ValueError("loss_grad: eps / target must be CUDA fp32 of the same shape, w CUDA fp32 [n]")
        d = torch.empty_like(eps)
        ss = torch.empty(n, device=eps.device)
        loss = torch.empty(1, device=eps.device)
        check(self._lib.salun_ddpm_eps_loss_grad(self.ctx.handle, _ptr(eps), _ptr(target), _ptr(w), n, eps.numel() // n,
                                                 _ptr(d), _ptr(ss), _ptr(loss), _stream(self.device)),
              "salun_ddpm_eps_loss_grad")
        return loss, d, ss


from .config import unet_param_table  # noqa: E402  (named_parameters() order and shapes, pure Python)


class UNetEngine:
    """The reference's ``model`` (Conditional_Model) for the hot path: parameters live in one flat fp32 arena."""

    native_layout = True  # conv weights are stored OHWI in the arena (flat.FlatSaliency converts before the top-k)

    def __init__(self, config, max_batch: int = 256, device=None, ctx: Optional[SalunContext] = None,
                 symmetric: bool = False, share_with: Optional["UNetEngine"] = None, precision: Optional[str] = None):
        """symmetric=True allocates the parameter / gradient arenas as torch symmetric memory (NVLink peer-mapped), which
        DistMaskedAdam needs for its fused reduce-scatter + clip + mask + Adam + all-gather kernels.
        share_with=<engine> builds a second set of activation buffers over the SAME parameter arena (a forward-only
        replica, e.g. for running the no-grad pseudo-label pass on another stream next to the main pass).
        precision: "bf16" (default) or "split" (bf16 hi/lo pairs, fp32-class products: the mask-generation mode)."""
        if share_with is not None:
            ctx = share_with.ctx if ctx is None else ctx
            precision = share_with.precision if precision is None else precision
        self.precision = precision or "bf16"
        m, d = config.model, config.data
        if not m.resamp_with_conv:
            raise NotImplementedError("resamp_with_conv=False is not used by the SalUn configs")
        self.config = config
        self.ctx = ctx if ctx is not None else SalunContext(device)
        self.device = self.ctx.device
        self._lib = _lib.lib(self.precision)
        self.max_batch = int(max_batch)
        mult, attn = list(m.ch_mult), list(m.attn_resolutions)
        self.cfg = salun_unet_cfg(m.ch, len(mult), (C.c_int * 8)(*mult), m.num_res_blocks, len(attn), (C.c_int * 8)(*attn),
                                  d.image_size, m.in_channels, m.out_ch, d.n_classes, self.max_batch, float(m.dropout))
        self.image_size, self.n_classes = d.image_size, d.n_classes
        self.cond_drop_prob = m.cond_drop_prob
        self.shapes = unet_param_table(config)
        self.names = list(self.shapes)
        self.n = int(self._lib.salun_unet_param_count(C.byref(self.cfg)))
        if self.n < 0:
            check(-1, "salun_unet_param_count")
        self.numel = sum(math.prod(s) for s in self.shapes.values())
        if self.n != self.numel:
            raise RuntimeError(f"parameter table of the host mirror ({self.numel}) and libsalun ({self.n}) disagree")
        self.offsets, off = {}, 0
        for k, s in self.shapes.items():
            self.offsets[k] = off
            off += math.prod(s)
        dev = self.device
        self.symmetric = bool(symmetric)
        if share_with is not None:
            if share_with.n != self.n:
                raise ValueError("share_with: the two engines must have the same architecture")
            self.params, self.grads, self.symmetric = share_with.params, share_with.grads, share_with.symmetric
        elif symmetric:
            import torch.distributed._symmetric_memory as symm_mem
            self.params = symm_mem.empty(self.n, dtype=torch.float32, device=dev)
            self.grads = symm_mem.empty(self.n, dtype=torch.float32, device=dev)
            self.params.zero_()
            self.grads.zero_()
        else:
            self.params = torch.zeros(self.n, device=dev)
            self.grads = torch.zeros(self.n, device=dev)
        h = C.c_void_p()
        check(self._lib.salun_unet_create(self.ctx.handle, C.byref(self.cfg), _ptr(self.params), _ptr(self.grads),
                                          C.byref(h)), "salun_unet_create")
        self._h = h
        self.training = True
        self._tensor_index = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.salun_unet_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- nn.Module-like surface ---------------------------------------------------------------------------------
    def train(self, mode: bool = True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def named_parameters(self):
        for k in self.shapes:
            yield k, self.get_param(k)

    def zero_grad(self):  # backward(accumulate=False) overwrites the arena
        pass

    # ---- layout conversion between the reference (OIHW) and the arena (OHWI) -------------------------------------
    def to_native(self, d: Dict[str, torch.Tensor], dtype=None) -> torch.Tensor:
        out = None
        for k, shp in self.shapes.items():
            t = d[k].reshape(shp).to(self.device)
            if out is None:
                out = torch.empty(self.n, dtype=dtype or t.dtype, device=self.device)
            if len(shp) == 4:
                t = t.permute(0, 2, 3, 1)
            out[self.offsets[k]: self.offsets[k] + math.prod(shp)] = t.reshape(-1)
        return out

    def from_native(self, flat: torch.Tensor, cpu: bool = False, key_prefix: str = "") -> "OrderedDict[str, torch.Tensor]":
        res = OrderedDict()
        for k, shp in self.shapes.items():
            t = flat[self.offsets[k]: self.offsets[k] + math.prod(shp)]
            if len(shp) == 4:
                t = t.reshape(shp[0], shp[2], shp[3], shp[1]).permute(0, 3, 1, 2)
            t = t.reshape(shp).contiguous().clone()
            res[key_prefix + k] = t.cpu() if cpu else t
        return res

    def from_native_flat(self, flat: torch.Tensor) -> torch.Tensor:
        """arena order -> the reference's flat order (torch.cat of the OIHW tensors, runners/diffusion.py:1006-1008)"""
        return torch.cat([t.reshape(-1) for t in self.from_native(flat).values()])

    # FlatParams-compatible spellings (flat.FlatMaskedAdam / FlatSaliency)
    def flat_from_dict(self, d: Dict[str, torch.Tensor], dtype=None) -> torch.Tensor:
        d = {(k[7:] if k.startswith("module.") else k): v for k, v in d.items()}
        return self.to_native(d, dtype=dtype)

    def dict_from_flat(self, flat: torch.Tensor, cpu: bool = False) -> "OrderedDict[str, torch.Tensor]":
        return self.from_native(flat, cpu=cpu)

    def get_param(self, name: str) -> torch.Tensor:
        shp = self.shapes[name]
        t = self.params[self.offsets[name]: self.offsets[name] + math.prod(shp)]
        if len(shp) == 4:
            return t.reshape(shp[0], shp[2], shp[3], shp[1]).permute(0, 3, 1, 2).contiguous()
        return t.reshape(shp).clone()

    def grad_dict(self) -> "OrderedDict[str, torch.Tensor]":
        return self.from_native(self.grads)

    def load_state_dict(self, sd, strict: bool = True):
        """states[0] of a reference checkpoint (runners/diffusion.py:500-506), with or without ``module.``."""
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
        missing = [k for k in self.shapes if k not in sd]
        if missing and strict:
            raise KeyError(f"missing keys: {missing[:5]}...")
        with torch.no_grad():
            self.params.copy_(self.to_native({k: (sd[k].float() if k in sd else self.get_param(k)) for k in self.shapes}))
        return self

    def state_dict(self, prefix: str = "") -> "OrderedDict[str, torch.Tensor]":
        return self.from_native(self.params, key_prefix=prefix)

    # ---- compute ------------------------------------------------------------------------------------------------
    def _check(self, x, t, c, drop):
        S = self.image_size
        if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4 and x.shape[1] == 3
                and x.shape[2] == x.shape[3] == S):
            raise ValueError(f"x must be a contiguous CUDA fp32 tensor [n,3,{S},{S}]")
        n = x.shape[0]
        if not 0 < n <= self.max_batch:
            raise ValueError(f"batch {n} exceeds max_batch {self.max_batch}")
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n):
            raise ValueError("t must be a contiguous CUDA fp32 tensor [n] (the reference passes t.float())")
        if not (c.is_cuda and c.dtype == torch.int64 and c.is_contiguous() and c.numel() == n):
            raise ValueError("c must be a contiguous CUDA int64 tensor [n]")
        if drop is not None and not (drop.is_cuda and drop.dtype == torch.uint8 and drop.is_contiguous() and drop.numel() == n):
            raise ValueError("drop must be a contiguous CUDA uint8 tensor [n]")

    def forward(self, x: torch.Tensor, t: torch.Tensor, c: torch.Tensor, drop: Optional[torch.Tensor] = None,
                save: bool = False, seed: int = 0, train: Optional[bool] = None) -> torch.Tensor:
        """eps = model._forward(x, t, c) (models/diffusion.py:357-413); drop[i] != 0 swaps in the null class embedding."""
        if drop is not None and drop.dtype == torch.bool:
            drop = drop.to(torch.uint8)
        self._check(x, t, c, drop)
        train = self.training if train is None else train
        eps = torch.empty_like(x)
        check(self._lib.salun_unet_forward(self._h, _ptr(x), _ptr(t), _ptr(c), _ptr(drop), x.shape[0], 1 if train else 0,
                                           int(seed) & 0xFFFFFFFFFFFFFFFF, 1 if save else 0, _ptr(eps),
                                           _stream(self.device)), "salun_unet_forward")
        return eps

    def backward(self, d_eps: torch.Tensor, accumulate: bool = False):
        """grads (=|+=) d loss / d params of the last forward(save=True), given d loss / d eps."""
        if not (d_eps.is_cuda and d_eps.dtype == torch.float32 and d_eps.is_contiguous()):
            raise ValueError("d_eps must be a contiguous CUDA fp32 tensor")
        check(self._lib.salun_unet_backward(self._h, _ptr(d_eps), 1 if accumulate else 0, _stream(self.device)),
              "salun_unet_backward")

    # ---- bring-up / parity tests ----------------------------------------------------------------------------------
    def tensor_names(self):
        if self._tensor_index is None:
            idx = OrderedDict()
            buf = C.create_string_buffer(128)
            cc, hh = C.c_int(), C.c_int()
            for i in range(int(self._lib.salun_unet_num_tensors(self._h))):
                check(self._lib.salun_unet_tensor_info(self._h, i, buf, 128, C.byref(cc), C.byref(hh)), "tensor_info")
                idx[buf.value.decode()] = (i, cc.value, hh.value)
            self._tensor_index = idx
        return self._tensor_index

    def export(self, name: str, n: int, grad: bool = False) -> torch.Tensor:
        i, cc, hh = self.tensor_names()[name]
        out = torch.empty(n, cc, hh, hh, device=self.device)
        check(self._lib.salun_unet_export_tensor(self._h, i, 1 if grad else 0, _ptr(out), _stream(self.device)), "export")
        return out


class DistMaskedAdam:
    """Data-parallel clip + mask + Adam (runners/diffusion.py:582-593 under DDP-averaged gradients) as two kernels over
    NVLink peer memory (salun_dp_grad_reduce_sumsq -> barrier -> salun_dp_masked_adam_step): each rank sums its 1/W shard
    of every peer's gradient arena, the W shard norms give the global pre-clip norm, the shard is updated (Adam moments
    exist only for the shard: ZeRO-1) and the new weights are stored into every replica.  Replaces NCCL all-reduce +
    grad_sumsq + clip_coef + masked_adam_step.  The engine must have been created with symmetric=True."""

    def __init__(self, engine: UNetEngine, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, mask=None,
                 max_norm: Optional[float] = None, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        if not engine.symmetric:
            raise ValueError("DistMaskedAdam needs UNetEngine(symmetric=True)")
        self.flat = self.engine = engine
        self.ctx = engine.ctx
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.lr, self.betas, self.eps, self.wd, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.step_count = 0
        self.mask_bits = None
        if mask is not None:
            self.mask_bits = self.ctx.pack_mask(engine.flat_from_dict(mask, dtype=torch.int64).contiguous())
        lib = engine._lib
        lo, hi = C.c_int64(), C.c_int64()
        check(lib.salun_dp_shard(engine.n, self.rank, self.world, C.byref(lo), C.byref(hi)), "salun_dp_shard")
        self.lo, self.hi = lo.value, hi.value
        m = max(4, self.hi - self.lo)
        dev = engine.device
        self.g_shard = torch.zeros(m, device=dev)
        self.exp_avg = torch.zeros(m, device=dev)        # shard only
        self.exp_avg_sq = torch.zeros(m, device=dev)
        self._norm = symm_mem.empty(2, dtype=torch.float64, device=dev)   # slot 0: this rank's shard sum of squares
        self._norm.zero_()
        self._coef_norm = torch.zeros(2, device=dev)
        self._hp = symm_mem.rendezvous(engine.params, self.group)
        self._hg = symm_mem.rendezvous(engine.grads, self.group)
        self._hn = symm_mem.rendezvous(self._norm, self.group)
        W = self.world
        self._pp = (C.c_void_p * W)(*[int(p) for p in self._hp.buffer_ptrs])
        self._gp = (C.c_void_p * W)(*[int(p) for p in self._hg.buffer_ptrs])
        self._np = (C.c_void_p * W)(*[int(p) for p in self._hn.buffer_ptrs])

    def zero_grad(self):
        pass

    def grad_norm(self) -> torch.Tensor:
        """pre-clip global norm of the averaged gradient (device scalar), as returned by clip_grad_norm_"""
        return self._coef_norm[1]

    def gather_state(self):
        """(exp_avg, exp_avg_sq) as full arena-layout tensors on every rank (checkpoints): all-gather of the shards"""
        import torch.distributed as dist
        per = (self.engine.n + self.world - 1) // self.world
        per = (per + 127) // 128 * 128
        outs = []
        for sh in (self.exp_avg, self.exp_avg_sq):
            pad = torch.zeros(per, device=sh.device)
            pad[: self.hi - self.lo] = sh[: self.hi - self.lo]
            full = torch.empty(per * self.world, device=sh.device)
            dist.all_gather_into_tensor(full, pad, group=self.group)
            outs.append(full[: self.engine.n].contiguous())
        return tuple(outs)

    def step(self):
        self.step_count += 1
        eng, lib, st = self.engine, self.engine._lib, _stream(self.engine.device)
        self._hg.barrier(channel=0)      # every rank's backward has written its gradient arena
        check(lib.salun_dp_grad_reduce_sumsq(self.ctx.handle, self._gp, _ptr(self.g_shard), _ptr(self._norm), eng.n,
                                             self.rank, self.world, st), "salun_dp_grad_reduce_sumsq")
        self._hn.barrier(channel=1)      # every rank's shard norm is visible
        check(lib.salun_dp_masked_adam_step(self.ctx.handle, self._pp, self._np, _ptr(self.g_shard), _ptr(self.exp_avg),
                                            _ptr(self.exp_avg_sq), _ptr(self.mask_bits), eng.n, self.rank, self.world,
                                            float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                                            float(self.wd), self.step_count,
                                            float(self.max_norm) if self.max_norm is not None else -1.0,
                                            _ptr(self._coef_norm), st), "salun_dp_masked_adam_step")
        self._hp.barrier(channel=2)      # every rank's shard of the new weights has landed in every replica
