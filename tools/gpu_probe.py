"""Bring-up probe run on the B200 box: prints max errors of the tcgen05 kernels, including descriptor variants."""
import ctypes as C
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from unlearn_saliency_b200 import _lib
from unlearn_saliency_b200.tail import SalunContext

ctx = SalunContext(0)
L = _lib.lib()
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
print(torch.cuda.get_device_name(0), flush=True)
GROUP = sys.argv[1] if len(sys.argv) > 1 else "all"
def want(g):
    return GROUP in ("all", g)


def run(name, fn):
    try:
        r = fn()
        torch.cuda.synchronize()
        print(name, "->", r, flush=True)
    except Exception as e:  # noqa
        print(name, "FAILED:", repr(e)[:300], flush=True)


def gemm(M, N, K):
    A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.full((M, N), float("nan"), device="cuda")
    rc = L.salun_gemm_bf16_tn(ctx.handle, p(A), p(B), p(out), None, M, N, K, st())
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    return rc, float((out - ref).abs().max()), float(ref.abs().max()), int(torch.isnan(out).sum())


for shp in [] if not want("gemm") else [(128, 64, 64), (128, 64, 128), (256, 128, 256), (4096, 256, 2304)]:
    run(f"gemm{shp}", lambda: gemm(*shp))


def conv(B, H, W, Cin, Cout, ks):
    x = torch.randn(B, Cin, H, W, device="cuda").bfloat16().float()
    w = (torch.randn(Cout, Cin, ks, ks, device="cuda") * 0.05).bfloat16().float()
    xpad = F.pad(x.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1)).contiguous().bfloat16()
    wk = w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous().bfloat16()
    M = B * H * W
    yf = torch.full((M, Cout), float("nan"), device="cuda")
    rc = L.salun_conv_fwd_bf16(ctx.handle, p(xpad), p(wk), None, p(yf), None, None, B, H, W, Cin, Cout, ks, st())
    torch.cuda.synchronize()
    ref = F.conv2d(x, w, padding=ks // 2).permute(0, 2, 3, 1).reshape(M, Cout)
    return rc, float((yf - ref).abs().max()), float(ref.abs().max())


for c in [] if not want("conv") else [(2, 32, 32, 64, 64, 3), (4, 16, 16, 128, 128, 3), (16, 4, 4, 512, 512, 3), (4, 16, 16, 64, 128, 1)]:
    run(f"conv{c}", lambda: conv(*c))


def wgrad(B, H, W, Cin, Cout, ks, swap, splits=0):
    x = torch.randn(B, Cin, H, W, device="cuda").bfloat16().float()
    dy = torch.randn(B, Cout, H, W, device="cuda").bfloat16().float()
    xpad = F.pad(x.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1)).contiguous().bfloat16()
    M = B * H * W
    dy2 = dy.permute(0, 2, 3, 1).reshape(M, Cout).contiguous().bfloat16()
    dw = torch.zeros(Cout, ks * ks * Cin, device="cuda")
    rc = L.salun_conv_wgrad_bf16(ctx.handle, p(dy2), p(xpad), p(dw), B, H, W, Cin, Cout, ks, splits, swap, st())
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(x, (Cout, Cin, ks, ks), dy, padding=ks // 2).permute(0, 2, 3, 1).reshape(Cout, -1)
    return rc, float((dw - ref).abs().max()), float(ref.abs().max())


for swap in (0, 1):
    if not want(f"wgrad{swap}"):
        continue
    for c in [(2, 32, 32, 64, 64, 3), (4, 16, 16, 128, 128, 3), (4, 16, 16, 64, 128, 1), (16, 4, 4, 512, 512, 3)]:
        run(f"wgrad{c} swap={swap}", lambda: wgrad(*c, swap))

# quick timing of the big shapes
def timeit(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for (B, H, W, Cin, Cout) in [] if not want("time") else [(256, 32, 32, 64, 64), (256, 16, 16, 128, 128), (256, 8, 8, 256, 256), (256, 4, 4, 512, 512)]:
    x = torch.randn(B, H + 2, W + 2, Cin, device="cuda").bfloat16()
    wk = torch.randn(Cout, 9 * Cin, device="cuda").bfloat16()
    M = B * H * W
    y = torch.empty(M, Cout, device="cuda", dtype=torch.bfloat16)
    dw = torch.zeros(Cout, 9 * Cin, device="cuda")
    try:
        t = timeit(lambda: L.salun_conv_fwd_bf16(ctx.handle, p(x), p(wk), p(y), None, None, None, B, H, W, Cin, Cout, 3, st()))
        fl = 2.0 * M * Cout * 9 * Cin
        print(f"conv_fwd {B}x{H}x{W} {Cin}->{Cout}: {t*1e3:.1f} us  {fl/t/1e9:.1f} TFLOP/s", flush=True)
        t = timeit(lambda: L.salun_conv_wgrad_bf16(ctx.handle, p(y), p(x), p(dw), B, H, W, Cin, Cout, 3, 0, 0, st()))
        print(f"conv_wgrad {B}x{H}x{W} {Cin}->{Cout}: {t*1e3:.1f} us  {fl/t/1e9:.1f} TFLOP/s", flush=True)
        xc = x[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2).float().contiguous(memory_format=torch.channels_last)
        wc = wk.view(Cout, 3, 3, Cin).permute(0, 3, 1, 2).float().contiguous(memory_format=torch.channels_last)
        t = timeit(lambda: F.conv2d(xc, wc, padding=1))
        print(f"  torch fp32(tf32={torch.backends.cudnn.allow_tf32}) conv2d: {t*1e3:.1f} us  {fl/t/1e9:.1f} TFLOP/s", flush=True)
        xb, wb = xc.bfloat16(), wc.bfloat16()
        t = timeit(lambda: F.conv2d(xb, wb, padding=1))
        print(f"  torch bf16 conv2d: {t*1e3:.1f} us  {fl/t/1e9:.1f} TFLOP/s", flush=True)
    except Exception as e:
        print("timing failed", repr(e)[:200], flush=True)
