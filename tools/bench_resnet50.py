"""ResNet-50 / ImageNet-shape masked SGD step (BASELINE.json configs[3] architecture) on one B200: steps/s at a given batch."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unlearn_saliency_b200.engine import MaskedSGD, ResNetEngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
eng = ResNetEngine("resnet50", 1000, 224, max_batch=B, imagenet=True, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225))
g = torch.Generator().manual_seed(0)
sd = {}
for k, shp in eng.table.items():
    if len(shp) == 4: sd[k] = torch.randn(shp, generator=g) * (2.0 / (shp[0] * shp[2] * shp[3])) ** 0.5
    elif k == "fc.weight": sd[k] = torch.randn(shp, generator=g) * 0.01
    elif k.endswith(".weight"): sd[k] = torch.ones(shp)
    else: sd[k] = torch.zeros(shp)
sd["normalize.mean"] = torch.tensor(eng.mean); sd["normalize.std"] = torch.tensor(eng.std)
eng.load_state_dict(sd)
bits = eng.ctx.pack_mask((torch.rand(eng.n_params, device="cuda") < 0.5).to(torch.int64))
opt = MaskedSGD(eng, 0.01, 0.9, 5e-4, bits)
x = torch.rand(B, 3, 224, 224, device="cuda"); y = torch.randint(0, 1000, (B,), device="cuda")
eng.train(True)
for _ in range(3): eng.forward_backward(x, y, loss_sign=-1.0); opt.step()   # GA step (GA.py:107-128)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
for _ in range(steps): eng.forward_backward(x, y, loss_sign=-1.0); opt.step()
e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / steps
gflop = 3 * 8.1784 * B
print(json.dumps({"metric": "ResNet-50/ImageNet-shape masked GA step (224x224, 1000 classes)", "batch": B, "ms_per_step": ms,
                  "steps_per_s": 1000 / ms, "images_per_s": B * 1000 / ms, "tflops": gflop / ms, "loss": float(eng._loss.item()),
                  "mem_GB": torch.cuda.max_memory_allocated() / 1e9, "params": eng.n_params}))
