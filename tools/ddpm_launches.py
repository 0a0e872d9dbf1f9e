"""One DDPM saliency_unlearn iteration (cifar10 U-Net, 128 remain + 128 forget images, rl) between cudaProfilerStart/Stop:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_ddpm_launches.csv \\
      python tools/ddpm_launches.py [bf16|split]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unlearn_saliency_b200.diffusion.config import cifar10_config               # noqa: E402
from unlearn_saliency_b200.diffusion.engine import UNetEngine                   # noqa: E402
from unlearn_saliency_b200.diffusion.runner import DDPMEngineUnlearner, get_beta_schedule   # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = 128
dev = torch.device("cuda:0")
eng = UNetEngine(cifar10_config(), max_batch=2 * B, device=dev, precision=precision)
g = torch.Generator().manual_seed(0)
sd = {}
for k, shp in eng.shapes.items():
    if "norm" in k:
        sd[k] = torch.ones(shp) if k.endswith("weight") else torch.zeros(shp)
    elif len(shp) >= 2:
        fan_in = 1
        for d in shp[1:]:
            fan_in *= d
        sd[k] = (torch.rand(shp, generator=g) * 2 - 1) / fan_in ** 0.5
    else:
        sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * 0.05
eng.load_state_dict(sd)
betas = torch.from_numpy(get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)).float()
un = DDPMEngineUnlearner(eng, betas, lr=1e-4, grad_clip=1.0)
un.opt.mask_bits = eng.ctx.pack_mask((torch.rand(eng.n, generator=g) < 0.5).to(torch.int64).to(dev))
un._overlap = False
xr, xf = torch.rand(B, 3, 32, 32, device=dev), torch.rand(B, 3, 32, 32, device=dev)
cr, cf = torch.randint(1, 10, (B,), device=dev), torch.zeros(B, dtype=torch.long, device=dev)
for _ in range(2):
    un.saliency_unlearn_step(xr, cr, xf, cf, alpha=1e-3, method="rl")
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
un.saliency_unlearn_step(xr, cr, xf, cf, alpha=1e-3, method="rl")
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
