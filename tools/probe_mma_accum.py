"""How exact is the tensor cores' fp32 accumulation?  D = A . B^T with non-negative bf16 operands (no cancellation) through
salun_gemm_bf16_tn (kind::f16, fp32 TMEM accumulator) against the float64 product of the same bf16 values, for growing K.
A truncating (round-toward-zero) accumulator shows up as a NEGATIVE mean relative error that grows linearly with K / 16."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unlearn_saliency_b200 import _lib  # noqa: E402
from unlearn_saliency_b200.tail import SalunContext, _ptr, _stream  # noqa: E402

ctx = SalunContext(0)
L = _lib.lib()
g = torch.Generator(device="cuda").manual_seed(0)
M, N = 256, 128
for K in (64, 256, 1024, 4096, 9216, 16384):
    for tag, lo in (("positive", 0.0), ("signed", -1.0)):
        A = (torch.rand(M, K, device="cuda", generator=g) * (1 - lo) + lo).bfloat16()
        B = (torch.rand(N, K, device="cuda", generator=g) * (1 - lo) + lo).bfloat16()
        out = torch.empty(M, N, device="cuda")
        rc = L.salun_gemm_bf16_tn(ctx.handle, _ptr(A), _ptr(B), _ptr(out), None, M, N, K, _stream(ctx.device))
        assert rc == 0
        ref = A.double() @ B.double().t()
        f32 = (A.float() @ B.float().t()).double()
        scale = (A.double().abs() @ B.double().abs().t())
        e = ((out.double() - ref) / scale)
        e32 = ((f32 - ref) / scale)
        print(f"K={K:6d} {tag:8s} tcgen05: mean {float(e.mean()):+.3e} rms {float(e.square().mean().sqrt()):.3e} | "
              f"torch fp32 matmul: mean {float(e32.mean()):+.3e} rms {float(e32.square().mean().sqrt()):.3e}   (relative to sum |a||b|)")
