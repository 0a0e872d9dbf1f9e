"""One ResNet-50 / ImageNet-shape masked GA step between cudaProfilerStart/Stop (ncu --profile-from-start off ...)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unlearn_saliency_b200.engine import MaskedSGD, ResNetEngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
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
for _ in range(2): eng.forward_backward(x, y, loss_sign=-1.0); opt.step()
torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart()
eng.forward_backward(x, y, loss_sign=-1.0); opt.step()
torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
