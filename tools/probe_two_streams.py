"""Does running two independent half-batch U-Net passes on two streams overlap the HBM-bound kernels of one with the
tensor-core kernels of the other?  fwd+bwd of 2 x 128 images: one engine at batch 256 vs two engines at batch 128
sequentially vs concurrently."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unlearn_saliency_b200.diffusion.engine import UNetEngine
from unlearn_saliency_b200.diffusion.config import cifar10_config

cfg = cifar10_config()
B = 128
def mk(n):
    e = UNetEngine(cfg, max_batch=n)
    e.params.normal_(0, 0.02)
    return e
def inputs(n):
    return (torch.randn(n, 3, 32, 32, device="cuda"), torch.randint(0, 1000, (n,), device="cuda").float(),
            torch.randint(0, 10, (n,), device="cuda"), torch.randn(n, 3, 32, 32, device="cuda") / n)
def timeit(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(it): fn()
    torch.cuda.synchronize(); return (time.time() - t0) / it * 1e3

big = mk(2 * B); xb = inputs(2 * B)
def run(e, x):
    e.forward(x[0], x[1], x[2], save=True, train=True, seed=1); e.backward(x[3])
print(f"one engine, batch 256: {timeit(lambda: run(big, xb)):.2f} ms")
big.close(); del big
e1, e2 = mk(B), mk(B); x1, x2 = inputs(B), inputs(B)
print(f"two engines, batch 128, sequential: {timeit(lambda: (run(e1, x1), run(e2, x2))):.2f} ms")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def conc():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): run(e1, x1)
    with torch.cuda.stream(s2): run(e2, x2)
    cur.wait_stream(s1); cur.wait_stream(s2)
print(f"two engines, batch 128, two streams: {timeit(conc):.2f} ms")
# host-side launch cost of one pass (no sync inside)
torch.cuda.synchronize(); t0 = time.time(); run(e1, x1); t1 = time.time(); torch.cuda.synchronize()
print(f"host time to enqueue one fwd+bwd: {(t1 - t0) * 1e3:.2f} ms")
