"""DDPM SalUn iterations/sec (cifar10 config, 128 remain + 128 forget images per iteration, method rl) on one GPU:
PyTorch forward/backward + fused sm_100a tail vs the reference's tail statements in stock PyTorch."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unlearn_saliency_b200.diffusion.runner import DDPMUnlearner, eps_loss, get_beta_schedule, q_sample
from unlearn_saliency_b200.diffusion.unet import ConditionalUNet, cifar10_config

B = int(os.environ.get("DDPM_BATCH", "128")); steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
torch.manual_seed(0)
betas = torch.from_numpy(get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)).float()
g = torch.Generator().manual_seed(1)
xr, cr = torch.rand(B, 3, 32, 32, generator=g), torch.randint(1, 10, (B,), generator=g)
xf, cf = torch.rand(B, 3, 32, 32, generator=g), torch.zeros(B, dtype=torch.long)

def timeit(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / steps

model = ConditionalUNet(cifar10_config()).cuda()
mask = {"module." + n: (torch.rand(p.shape, generator=g) < 0.5).to(torch.int64) for n, p in model.named_parameters()}  # CPU int64 like torch.load
un = DDPMUnlearner(model, betas, mask=mask)
ms_fused = timeit(lambda: un.saliency_unlearn_step(xr, cr, xf, cf, alpha=1e-3, method="rl"))
# tail only
def tail_fused():
    un.opt.step()
ms_tail_fused = timeit(tail_fused)
ref = ConditionalUNet(cifar10_config()).cuda()
opt = torch.optim.Adam(ref.parameters(), lr=1e-4)
bd = betas.cuda()
def ref_tail():
    torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
    for n, p in ref.named_parameters():
        if p.grad is not None:
            p.grad *= mask["module." + n].to(p.device)      # the per-step 309 MB H2D of runners/diffusion.py:589-592
    opt.step()
def ref_step():
    ref.train()
    a, b_, c_, d_ = 2 * xr.cuda() - 1, cr.cuda(), 2 * xf.cuda() - 1, cf.cuda()
    t = torch.randint(0, 1000, (B,), device="cuda"); e = torch.randn_like(a)
    remain = eps_loss(ref, a, t, b_, e, bd)
    xt = q_sample(c_, t, e, bd)
    out = ref(xt, t.float(), d_, mode="train"); pseudo = ref(xt, t.float(), (d_ + 1) % 10, mode="train").detach()
    loss = torch.nn.functional.mse_loss(out, pseudo) + 1e-3 * remain
    opt.zero_grad(); loss.backward(); ref_tail()
ms_ref = timeit(ref_step)
ms_tail_ref = timeit(ref_tail)
print(json.dumps({"metric": "DDPM saliency_unlearn iterations/sec (cifar10 U-Net, 128+128 images, rl)", "n_gpus": 1,
                  "fused_tail_its_per_s": 1000 / ms_fused, "stock_pytorch_its_per_s": 1000 / ms_ref,
                  "ms_per_it_fused": ms_fused, "ms_per_it_stock": ms_ref, "tail_ms_fused": ms_tail_fused,
                  "tail_ms_stock": ms_tail_ref, "params": un.flat.numel,
                  "note": "U-Net fwd/bwd is PyTorch (cuDNN, TF32) in both arms; only the clip+mask+Adam tail differs"}))
