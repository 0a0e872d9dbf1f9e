"""Host-side wrappers of the HBM-bound tail kernels (include/salun.h) on torch CUDA tensors.

PyTorch is plumbing here: it owns the device buffers and the stream; all arithmetic runs in
libsalun.so.  Mirrors, one to one, the reference statements cited in include/salun.h.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import check, salun_topk_info


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _req(t: torch.Tensor, dtype, name: str):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise ValueError(f"{name} must be a contiguous CUDA tensor of dtype {dtype}")


def topk_count(n: int, ratio: float) -> int:
    """k = int(len(all_elements) * i)  -- Classification/generate_mask.py:60 (python double arithmetic)."""
    return int(n * ratio)


def mask_words(n: int) -> int:
    return (n + 31) // 32


class SalunContext:
    """One libsalun context (device workspaces) bound to a CUDA device.  Not thread-safe."""

    def __init__(self, device: Optional[int | torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("SalunContext needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self._lib = _lib.lib()
        h = C.c_void_p()
        check(self._lib.salun_ctx_create(dev.index, C.byref(h)), "salun_ctx_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.salun_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def ensure_op_scratch(self, nbytes: int = 32 << 20):
        """Context-owned scratch for the split-K path of the op-level convolutions (salun_op_set_scratch); allocated once and
        kept for the life of the context because captured graphs hold its address."""
        if getattr(self, "_op_scratch", None) is None:
            self._op_scratch = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            check(self._lib.salun_op_set_scratch(self._h, C.c_void_p(self._op_scratch.data_ptr()), nbytes), "salun_op_set_scratch")
        return self._op_scratch

    # ---- (i) mask generation tail ------------------------------------------------------
    def saliency_accumulate(self, grads: Sequence[torch.Tensor], accum_flat: torch.Tensor,
                            scale: Optional[torch.Tensor] = None):
        """gradients[name] += param.grad.data for every tensor, into one flat arena (generate_mask.py:41-44)."""
        _req(accum_flat, torch.float32, "accum_flat")
        n_t = len(grads)
        ptrs = (C.c_void_p * n_t)()
        numels = (C.c_int64 * n_t)()
        total = 0
        for i, g in enumerate(grads):
            _req(g, torch.float32, f"grads[{i}]")
            ptrs[i] = g.data_ptr()
            numels[i] = g.numel()
            total += g.numel()
        if total != accum_flat.numel():
            raise ValueError(f"accum_flat has {accum_flat.numel()} elements, grads sum to {total}")
        check(self._lib.salun_saliency_accumulate(self._h, ptrs, numels, n_t, _ptr(accum_flat), _ptr(scale),
                                                  _stream(self.device)), "salun_saliency_accumulate")

    def saliency_accumulate_flat(self, grad: torch.Tensor, accum: torch.Tensor, scale: Optional[torch.Tensor] = None):
        _req(grad, torch.float32, "grad"); _req(accum, torch.float32, "accum")
        if grad.numel() != accum.numel():
            raise ValueError("grad / accum size mismatch")
        check(self._lib.salun_saliency_accumulate_flat(self._h, _ptr(grad), _ptr(accum), accum.numel(), _ptr(scale),
                                                       _stream(self.device)), "salun_saliency_accumulate_flat")

    def abs_(self, a: torch.Tensor):
        _req(a, torch.float32, "a")
        check(self._lib.salun_abs_inplace(self._h, _ptr(a), a.numel(), _stream(self.device)), "salun_abs_inplace")
        return a

    def topk_mask(self, accum: torch.Tensor, k: int, want_i64: bool = True, want_bits: bool = True,
                  want_info: bool = False, out_i64: Optional[torch.Tensor] = None,
                  out_bits: Optional[torch.Tensor] = None):
        """Global top-k mask of |accum| (generate_mask.py:57-80).  Returns (mask_i64, mask_bits, info)."""
        _req(accum, torch.float32, "accum")
        n = accum.numel()
        m64 = bits = None
        if want_i64:
            m64 = out_i64 if out_i64 is not None else torch.empty(n, dtype=torch.int64, device=accum.device)
            _req(m64, torch.int64, "out_i64")
        if want_bits:
            bits = out_bits if out_bits is not None else torch.empty(mask_words(n), dtype=torch.int32, device=accum.device)
            _req(bits, torch.int32, "out_bits")
        info = salun_topk_info() if want_info else None
        check(self._lib.salun_topk_mask(self._h, _ptr(accum), n, int(k), _ptr(m64), _ptr(bits),
                                        C.byref(info) if info is not None else None, _stream(self.device)),
              "salun_topk_mask")
        return m64, bits, info

    def topk_mask_multi(self, accum: torch.Tensor, ks: Sequence[int], want_i64: bool = True, want_bits: bool = True,
                        want_info: bool = False):
        """The reference's sweep over threshold_list (generate_mask.py:50-82) as one call: the saliencies are read once
        per pass for ALL ratios.  Returns (list of mask_i64, list of mask_bits, list of info); entry r is bit-identical to
        topk_mask(accum, ks[r])."""
        _req(accum, torch.float32, "accum")
        n, R = accum.numel(), len(ks)
        if R > 16:
            raise ValueError("at most 16 ratios per call")
        m64 = [torch.empty(n, dtype=torch.int64, device=accum.device) for _ in range(R)] if want_i64 else None
        bits = [torch.empty(mask_words(n), dtype=torch.int32, device=accum.device) for _ in range(R)] if want_bits else None
        kk = (C.c_int64 * R)(*[int(k) for k in ks])
        p64 = (C.c_void_p * R)(*[t.data_ptr() for t in m64]) if want_i64 else None
        pb = (C.c_void_p * R)(*[t.data_ptr() for t in bits]) if want_bits else None
        infos = (salun_topk_info * R)() if want_info else None
        check(self._lib.salun_topk_mask_multi(self._h, _ptr(accum), n, kk, R, p64, pb, infos, _stream(self.device)),
              "salun_topk_mask_multi")
        return m64, bits, (list(infos) if infos is not None else None)

    def pack_mask(self, mask_i64: torch.Tensor) -> torch.Tensor:
        _req(mask_i64, torch.int64, "mask_i64")
        n = mask_i64.numel()
        bits = torch.empty(mask_words(n), dtype=torch.int32, device=mask_i64.device)
        check(self._lib.salun_pack_mask(self._h, _ptr(mask_i64), n, _ptr(bits), _stream(self.device)), "salun_pack_mask")
        return bits

    def unpack_mask(self, bits: torch.Tensor, n: int) -> torch.Tensor:
        _req(bits, torch.int32, "bits")
        out = torch.empty(n, dtype=torch.int64, device=bits.device)
        check(self._lib.salun_unpack_mask(self._h, _ptr(bits), n, _ptr(out), _stream(self.device)), "salun_unpack_mask")
        return out

    # ---- (ii) masked step tail ---------------------------------------------------------
    def apply_mask(self, g: torch.Tensor, bits: torch.Tensor):
        _req(g, torch.float32, "g"); _req(bits, torch.int32, "bits")
        check(self._lib.salun_apply_mask(self._h, _ptr(g), _ptr(bits), g.numel(), _stream(self.device)), "salun_apply_mask")

    def masked_sgd_step(self, p, g, v, bits, lr: float, momentum: float, wd: float):
        for t, nm in ((p, "p"), (g, "g"), (v, "v")):
            _req(t, torch.float32, nm)
        if bits is not None:
            _req(bits, torch.int32, "bits")
        check(self._lib.salun_masked_sgd_step(self._h, _ptr(p), _ptr(g), _ptr(v), _ptr(bits), p.numel(),
                                              float(lr), float(momentum), float(wd), _stream(self.device)),
              "salun_masked_sgd_step")

    def grad_sumsq(self, g: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        _req(g, torch.float32, "g")
        out = out if out is not None else torch.empty(1, dtype=torch.float64, device=g.device)
        check(self._lib.salun_grad_sumsq(self._h, _ptr(g), g.numel(), _ptr(out), _stream(self.device)), "salun_grad_sumsq")
        return out

    def l1_penalty_grad(self, p: torch.Tensor, g: torch.Tensor, alpha: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """g += alpha * sign(p); returns sum |p| (float64 device scalar) -- FT.py:13-17,133-134."""
        _req(p, torch.float32, "p")
        _req(g, torch.float32, "g")
        out = out if out is not None else torch.empty(1, dtype=torch.float64, device=g.device)
        check(self._lib.salun_l1_penalty_grad(self._h, _ptr(p), _ptr(g), p.numel(), float(alpha), _ptr(out),
                                              _stream(self.device)), "salun_l1_penalty_grad")
        return out

    def clip_coef(self, sumsq: torch.Tensor, max_norm: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        out = out if out is not None else torch.empty(1, dtype=torch.float32, device=sumsq.device)
        check(self._lib.salun_clip_coef(self._h, _ptr(sumsq), float(max_norm), _ptr(out), _stream(self.device)), "salun_clip_coef")
        return out

    def masked_adam_step(self, p, g, m1, m2, bits, lr, beta1, beta2, eps, wd, step: int, coef=None):
        for t, nm in ((p, "p"), (g, "g"), (m1, "m1"), (m2, "m2")):
            _req(t, torch.float32, nm)
        check(self._lib.salun_masked_adam_step(self._h, _ptr(p), _ptr(g), _ptr(m1), _ptr(m2), _ptr(bits), p.numel(),
                                               float(lr), float(beta1), float(beta2), float(eps), float(wd), int(step),
                                               _ptr(coef), _stream(self.device)), "salun_masked_adam_step")
