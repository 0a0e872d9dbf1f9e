"""host enqueue time of one engine step (no sync) vs the device time of the step"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unlearn_saliency_b200.engine import ResNetEngine, MaskedSGD
from unlearn_saliency_b200.tail import SalunContext

dev = torch.device("cuda:0")
ctx = SalunContext(0)
eng = ResNetEngine("resnet18", 10, 32, 256, device=dev, ctx=ctx)
g = torch.Generator(device="cpu").manual_seed(0)
sd = {}
for k, shp in eng.table.items():
    if len(shp) == 4:
        sd[k] = torch.randn(shp, generator=g) * (2.0 / (shp[0] * shp[2] * shp[3])) ** 0.5
    elif k == "fc.weight":
        sd[k] = (torch.rand(shp, generator=g) * 2 - 1) / shp[1] ** 0.5
    elif k.endswith(".weight"):
        sd[k] = torch.ones(shp)
    else:
        sd[k] = torch.zeros(shp)
eng.load_state_dict(sd)
opt = MaskedSGD(eng, 0.013, 0.9, 5e-4, mask_bits=None)
eng.train(True)
x = torch.rand(256, 3, 32, 32, device=dev); y = torch.randint(0, 10, (256,), device=dev)
for _ in range(10):
    eng.forward_backward(x, y); opt.step()
torch.cuda.synchronize()
host = []
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(50):
    t0 = time.perf_counter()
    eng.forward_backward(x, y); opt.step()
    host.append(time.perf_counter() - t0)
e1.record(); torch.cuda.synchronize()
host.sort()
print(f"host enqueue per step: median {host[25]*1e3:.3f} ms, min {host[0]*1e3:.3f} ms, max {host[-1]*1e3:.3f} ms; device {e0.elapsed_time(e1)/50:.3f} ms/step")
# one isolated step: host enqueue when the GPU is idle
torch.cuda.synchronize()
t0 = time.perf_counter(); eng.forward_backward(x, y); opt.step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"isolated step: enqueue {(t1-t0)*1e3:.3f} ms, until idle {(t2-t0)*1e3:.3f} ms")
