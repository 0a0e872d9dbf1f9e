"""Stand-alone conv forward launches (salun_conv_fwd_bf16, implicit GEMM through 4-D TMA boxes) on the DDPM U-Net's
dominant shapes: CUDA-event timing, and a target for `ncu --set full -k regex:k_conv_gemm_p`."""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unlearn_saliency_b200 import _lib
from unlearn_saliency_b200._lib import check
from unlearn_saliency_b200.tail import SalunContext

ctx = SalunContext(0)
L = _lib.lib()
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
shapes = [(256, 32, 128, 128), (256, 16, 256, 256), (256, 32, 256, 128), (256, 16, 512, 256)]
if len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
for B, H, Cin, Cout in shapes:
    xpad = torch.zeros(B, H + 2, H + 2, Cin, device="cuda", dtype=torch.bfloat16)
    xpad[:, 1:-1, 1:-1] = torch.randn(B, H, H, Cin, device="cuda").bfloat16()
    wk = (torch.randn(Cout, 9 * Cin, device="cuda") * 0.03).bfloat16()
    M = B * H * H
    y = torch.empty(M, Cout, device="cuda", dtype=torch.bfloat16)
    f = lambda: check(L.salun_conv_fwd_bf16(ctx.handle, p(xpad), p(wk), p(y), None, None, None, B, H, H, Cin, Cout, 3, st()), "conv")
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    t = tot / iters
    fl = 2.0 * M * Cout * 9 * Cin / 1e9
    print(f"conv B={B} H={H} Cin={Cin} Cout={Cout}: {t*1e3:.1f} us {fl/t:.0f} TFLOP/s", flush=True)
