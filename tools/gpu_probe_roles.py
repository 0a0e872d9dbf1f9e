"""role timing of the tensor-core kernels: where do the producer / MMA issuer / epilogue spend their cycles?"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unlearn_saliency_b200 import _lib
from unlearn_saliency_b200.tail import SalunContext
ctx = SalunContext(0); L = _lib.lib()
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
buf = torch.zeros(8 * 2048, dtype=torch.int64, device="cuda")

def show(name, nctas):
    torch.cuda.synchronize()
    b = buf.view(-1, 8)[:nctas].double()
    m = b.mean(0).tolist()
    print(f"{name:46s} producer: wait_empty {m[0]:8.0f} / total {m[1]:8.0f} | mma: wait_full {m[2]:8.0f} wait_tmem {m[3]:7.0f} / total {m[4]:8.0f} | epi: wait_tfull {m[5]:8.0f} / total {m[6]:8.0f}  (cycles, mean over {nctas} CTAs)", flush=True)
    buf.zero_()

def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3

for (B, H, Cin, Cout) in [(256, 32, 64, 64), (256, 16, 128, 128), (256, 8, 256, 256), (256, 4, 512, 512)]:
    W = H; M = B * H * W
    x = torch.randn(B, H + 2, W + 2, Cin, device="cuda").bfloat16()
    wk = torch.randn(Cout, 9 * Cin, device="cuda").bfloat16()
    y = torch.empty(M, Cout, device="cuda", dtype=torch.bfloat16)
    dw = torch.zeros(Cout, 9 * Cin, device="cuda")
    f_gemm = lambda: L.salun_conv_fwd_bf16(ctx.handle, p(x), p(wk), p(y), None, None, None, B, H, W, Cin, Cout, 3, st())
    f_wg = lambda: L.salun_conv_wgrad_bf16(ctx.handle, p(y), p(x), p(dw), B, H, W, Cin, Cout, 3, 0, 0, st())
    t = timeit(f_gemm)
    L.salun_debug_role_timing(p(buf)); f_gemm(); show(f"conv_gemm_p {H}x{W} {Cin}->{Cout}  {t:6.1f} us", min(148, (M // 128) * max(1, Cout // 128)))
    L.salun_debug_role_timing(None)
    if H in (16, 32) and Cin in (64, 128):
        f_rw = lambda: L.salun_conv_rw_fwd_bf16(ctx.handle, p(x), p(wk), p(y), None, None, B, H, W, Cin, Cout, st())
        t = timeit(f_rw)
        L.salun_debug_role_timing(p(buf)); f_rw(); show(f"conv_rw     {H}x{W} {Cin}->{Cout}  {t:6.1f} us", 148 // (Cout // 64) * (Cout // 64))
        L.salun_debug_role_timing(None)
    t = timeit(f_wg)
    L.salun_debug_role_timing(p(buf)); f_wg(); show(f"wgrad       {H}x{W} {Cin}->{Cout}  {t:6.1f} us", 290)
    L.salun_debug_role_timing(None)
