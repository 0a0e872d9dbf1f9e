"""Time the 1-CTA persistent GEMM against the CTA-pair (cta_group::2) GEMM on the step's GEMM shapes."""
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

def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters

shapes = [(16384, 256, 2304), (4096, 512, 4608), (65536, 128, 1152), (8192, 8192, 8192), (16384, 1024, 1024),
          (200704, 256, 64), (12544, 512, 1024), (3136, 2048, 512)]
if len(sys.argv) > 1 and sys.argv[1] == "unet":  # conv GEMM shapes of the DDPM U-Net at batch 256 (M = n*H*H, N = Cout, K = 9*Cin)
    shapes = [(262144, 128, 1152), (262144, 128, 3456), (262144, 128, 2304), (65536, 256, 2304), (65536, 256, 4608),
              (65536, 256, 3456), (16384, 256, 4608), (4096, 256, 4608), (262144, 384, 1152), (65536, 512, 2304)]
for M, N, K in shapes:
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = torch.randn(N, K, device="cuda").bfloat16()
    o1 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    o2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    f1 = lambda: check(L.salun_gemm_bf16_tn(ctx.handle, p(A), p(B), None, p(o1), M, N, K, st()), "g1")
    f2 = lambda: check(L.salun_gemm2_bf16_tn(ctx.handle, p(A), p(B), None, p(o2), M, N, K, st()), "g2")
    f3 = lambda: torch.matmul(A, B.t())
    t1, t2, t3 = timeit(f1), timeit(f2), timeit(f3)
    fl = 2.0 * M * N * K / 1e9
    same = torch.equal(o1, o2)
    print(f"M={M} N={N} K={K}: 1cta {t1*1e3:.1f} us {fl/t1:.0f} TF | 2cta {t2*1e3:.1f} us {fl/t2:.0f} TF | cublas {t3*1e3:.1f} us {fl/t3:.0f} TF | equal={same}", flush=True)
