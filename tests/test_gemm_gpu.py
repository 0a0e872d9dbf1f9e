"""tcgen05 GEMM / implicit-GEMM convolution kernels against a plain PyTorch fp32 reference of the same op.

Inputs are bf16-representable, accumulation is fp32 in TMEM: the only difference to the fp32 reference is
summation order, so tolerances are tight (rtol 2e-3 on bf16 outputs = 1 bf16 ulp, 1e-4 on fp32 outputs).
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from unlearn_saliency_b200 import _lib
from unlearn_saliency_b200._lib import check

pytestmark = pytest.mark.gpu


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 128, 128), (1024, 64, 576), (4096, 512, 4608),
                                    (200, 64, 64), (8192, 256, 1152)])
def test_gemm_tn(salun_ctx, M, N, K):
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.full((M, N), float("nan"), device="cuda")
    outb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    check(_lib.lib().salun_gemm_bf16_tn(salun_ctx.handle, _p(A), _p(B), _p(out), _p(outb), M, N, K, _st()), "gemm")
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-3 * K ** 0.5 / 8)
    torch.testing.assert_close(outb.float(), ref, rtol=8e-3, atol=2e-2 * K ** 0.5 / 8)


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (256, 256, 128), (512, 128, 576), (4096, 512, 4608),
                                    (200, 128, 64), (8192, 256, 1152), (16384, 1024, 1024), (300, 384, 192)])
def test_gemm2_tn(salun_ctx, M, N, K):
    """CTA-pair kernel (tcgen05 cta_group::2): same contract, including ragged M and an odd number of 128-row tiles."""
    torch.manual_seed(M + N + K + 1)
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.full((M, N), float("nan"), device="cuda")
    outb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    check(_lib.lib().salun_gemm2_bf16_tn(salun_ctx.handle, _p(A), _p(B), _p(out), _p(outb), M, N, K, _st()), "gemm2")
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-3 * K ** 0.5 / 8)
    torch.testing.assert_close(outb.float(), ref, rtol=8e-3, atol=2e-2 * K ** 0.5 / 8)


def _pad_nhwc(x_nchw):
    """fp32 NCHW -> bf16 halo-padded NHWC [N][H+2][W+2][C]"""
    return F.pad(x_nchw.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1)).contiguous().bfloat16()


CONV_CASES = [(2, 32, 32, 64, 64, 3), (4, 16, 16, 128, 128, 3), (8, 8, 8, 256, 256, 3), (16, 4, 4, 512, 512, 3),
              (4, 16, 16, 64, 128, 1), (2, 32, 32, 128, 64, 3)]


@pytest.mark.parametrize("B,H,W,Cin,Cout,ks", CONV_CASES)
def test_conv_fwd(salun_ctx, B, H, W, Cin, Cout, ks):
    torch.manual_seed(B * H + Cin)
    x = torch.randn(B, Cin, H, W, device="cuda").bfloat16().float()
    w = (torch.randn(Cout, Cin, ks, ks, device="cuda") * (2.0 / (Cin * ks * ks)) ** 0.5).bfloat16().float()
    xpad = _pad_nhwc(x)
    wk = w.permute(0, 2, 3, 1).reshape(Cout, ks * ks * Cin).contiguous().bfloat16()  # [Cout][tap][Cin]
    M = B * H * W
    y = torch.empty(M, Cout, device="cuda", dtype=torch.bfloat16)
    yf = torch.empty(M, Cout, device="cuda")
    ssum = torch.zeros(M // 128 * 4, Cout, device="cuda")
    ssq = torch.zeros_like(ssum)
    check(_lib.lib().salun_conv_fwd_bf16(salun_ctx.handle, _p(xpad), _p(wk), _p(y), _p(yf), _p(ssum), _p(ssq),
                                         B, H, W, Cin, Cout, ks, _st()), "conv_fwd")
    torch.cuda.synchronize()
    ref = F.conv2d(x, w, padding=ks // 2).permute(0, 2, 3, 1).reshape(M, Cout)
    torch.testing.assert_close(yf, ref, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(y.float(), ref, rtol=8e-3, atol=2e-2)
    torch.testing.assert_close(ssum.sum(0), ref.sum(0), rtol=1e-3, atol=1e-2 * (M ** 0.5))
    torch.testing.assert_close(ssq.sum(0), (ref * ref).sum(0), rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("B,H,W,Cin,Cout,ks", CONV_CASES)
def test_conv_wgrad(salun_ctx, B, H, W, Cin, Cout, ks):
    torch.manual_seed(B * W + Cout)
    x = torch.randn(B, Cin, H, W, device="cuda").bfloat16().float()
    dy = torch.randn(B, Cout, H, W, device="cuda").bfloat16().float()
    xpad = _pad_nhwc(x)
    M = B * H * W
    dy2 = dy.permute(0, 2, 3, 1).reshape(M, Cout).contiguous().bfloat16()
    dw = torch.zeros(Cout, ks * ks * Cin, device="cuda")
    check(_lib.lib().salun_conv_wgrad_bf16(salun_ctx.handle, _p(dy2), _p(xpad), _p(dw), B, H, W, Cin, Cout, ks, 0, 0,
                                           _st()), "conv_wgrad")
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(x, (Cout, Cin, ks, ks), dy, padding=ks // 2)  # [Cout][Cin][kh][kw]
    ref = ref.permute(0, 2, 3, 1).reshape(Cout, ks * ks * Cin)
    torch.testing.assert_close(dw, ref, rtol=1e-3, atol=2e-3 * M ** 0.5)
