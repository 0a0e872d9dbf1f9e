"""Role timing of the persistent GEMM on the short-K shapes of the DDPM attention blocks (q/k/v/proj 1x1 convolutions:
M = 65536, N = 256, K = 256): who waits on whom -- TMA producer, MMA issuer or the epilogue warps?"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unlearn_saliency_b200 import _lib
from unlearn_saliency_b200.tail import SalunContext
ctx = SalunContext(0); L = _lib.lib()
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
buf = torch.zeros(8 * 2048, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=10):
    fn(); torch.cuda.synchronize(); tot = 0.0
    for _ in range(n):
        flush.zero_(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n * 1e3
for (M, N, K, f32) in [(65536, 256, 256, False), (65536, 256, 256, True), (65536, 128, 256, False), (65536, 256, 1024, False), (65536, 256, 2304, False)]:
    A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16()
    ob = None if f32 else torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    of = torch.empty(M, N, device="cuda") if f32 else None
    f = lambda: L.salun_gemm_bf16_tn(ctx.handle, p(A), p(B), p(of), p(ob), M, N, K, st())
    t = timeit(f)
    L.salun_debug_role_timing(p(buf)); f(); torch.cuda.synchronize()
    b = buf.view(-1, 8)[:148].double().mean(0).tolist(); buf.zero_(); L.salun_debug_role_timing(None)
    print(f"M={M} N={N} K={K} out={'f32' if f32 else 'bf16'}: {t:6.1f} us {2.0*M*N*K/t/1e6:6.0f} TFLOP/s | producer wait_empty {b[0]:7.0f}/{b[1]:7.0f} | "
          f"mma wait_full {b[2]:7.0f} wait_tmem {b[3]:7.0f} /{b[4]:7.0f} | epi wait_tfull {b[5]:7.0f}/{b[6]:7.0f}", flush=True)
