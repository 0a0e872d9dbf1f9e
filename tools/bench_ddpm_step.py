"""DDPM SalUn saliency_unlearn iterations/sec (BASELINE.json configs[2] on ONE GPU): cifar10 U-Net, 128 remain + 128
forget images per iteration, method rl, dropout 0.1, 50% mask, clip 1.0, Adam -- the sm_100a engine (U-Net forward /
backward + fused tail, all libsalun kernels) vs the reference's statements in stock PyTorch on the same GPU.

    python tools/bench_ddpm_step.py [steps] [--no-ref] [--profile]

Algorithmic work: 7 forward-equivalents x 128 images x 12.449 GFLOP = 11.15 TFLOP / iteration.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from unlearn_saliency_b200 import _lib  # noqa: E402
from unlearn_saliency_b200.diffusion.engine import UNetEngine  # noqa: E402
from unlearn_saliency_b200.diffusion.runner import DDPMEngineUnlearner, eps_loss, get_beta_schedule, q_sample  # noqa: E402
from oracle.unet import ConditionalUNet, cifar10_config  # noqa: E402

B = int(os.environ.get("DDPM_BATCH", "128"))
args = [a for a in sys.argv[1:] if not a.startswith("--")]
steps = int(args[0]) if args else 10
IT_TFLOP = 7 * B * 12.449e-3


def timeit(fn, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    torch.manual_seed(0)
    cfg = cifar10_config()  # dropout 0.1, cond_drop_prob 0.1
    betas = torch.from_numpy(get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)).float()
    g = torch.Generator().manual_seed(1)
    xr, cr = torch.rand(B, 3, 32, 32, generator=g).pin_memory(), torch.randint(1, 10, (B,), generator=g).pin_memory()
    xf, cf = torch.rand(B, 3, 32, 32, generator=g).pin_memory(), torch.zeros(B, dtype=torch.long).pin_memory()
    ref = ConditionalUNet(cfg).cuda()
    mask = {"module." + n: (torch.rand(p.shape, generator=g) < 0.5).to(torch.int64) for n, p in ref.named_parameters()}
    eng = UNetEngine(cfg, max_batch=2 * B)
    eng.load_state_dict(ref.state_dict())
    un = DDPMEngineUnlearner(eng, betas, lr=1e-4, grad_clip=1.0, mask=mask)
    lib = _lib.lib()
    res = {"metric": "DDPM saliency_unlearn iterations/sec (cifar10 U-Net, 128+128 images, rl, dropout 0.1)", "n_gpus": 1,
           "batch": [B, B], "tflop_per_it": IT_TFLOP}
    losses = []

    def step_engine():
        losses.append(un.saliency_unlearn_step(xr, cr, xf, cf, alpha=1e-3, method="rl"))

    if "--profile" in sys.argv:  # under ncu: two iterations, nothing else
        step_engine()
        step_engine()
        torch.cuda.synchronize()
        return
    l0 = lib.salun_launch_count()
    ms = timeit(step_engine)
    res.update(engine_ms_per_it=ms, engine_its_per_s=1000 / ms, engine_tflops=IT_TFLOP / ms * 1e3,
               launches_per_it=(lib.salun_launch_count() - l0) / (steps + 3),
               engine_loss_first=float(losses[0]), engine_loss_last=float(losses[-1]))
    # device-resident inputs (no H2D inside the step)
    xr_d, cr_d, xf_d, cf_d = xr.cuda(), cr.cuda(), xf.cuda(), cf.cuda()
    ms_d = timeit(lambda: un.saliency_unlearn_step(xr_d, cr_d, xf_d, cf_d, alpha=1e-3, method="rl"))
    res.update(engine_ms_per_it_resident=ms_d)
    x2 = torch.randn(2 * B, 3, 32, 32, device="cuda")
    t2 = torch.randint(0, 1000, (2 * B,), device="cuda").float()
    c2 = torch.randint(0, 10, (2 * B,), device="cuda")
    d2 = torch.randn(2 * B, 3, 32, 32, device="cuda") / B
    res["fwd_ms_256"] = timeit(lambda: eng.forward(x2, t2, c2, save=True, train=True, seed=1))
    res["fwd_bwd_ms_256"] = timeit(lambda: (eng.forward(x2, t2, c2, save=True, train=True, seed=1), eng.backward(d2)))
    res["tail_ms"] = timeit(un.opt.step)
    # mask generation (runners/diffusion.py:959-1039): one forget batch of 128 = conditional + null pass as a batch of 256,
    # backward, clip, accumulate; then |.| + global top-k over 38.6 M saliencies and the int64 mask dict
    res["maskgen_batch_ms_128"] = timeit(lambda: un.generate_mask_batch(xf_d, cf_d, cond_scale=2.0))
    import time
    torch.cuda.synchronize()
    t0 = time.time()
    m, info = un.finish_mask(None, 0.5)
    torch.cuda.synchronize()
    res["maskgen_select_and_export_ms"] = (time.time() - t0) * 1e3
    res["maskgen_ones"] = int(sum(int(v.sum()) for v in m.values()))
    if "--no-ref" not in sys.argv:
        opt = torch.optim.Adam(ref.parameters(), lr=1e-4)
        bd = betas.cuda()
        ref.train()

        def ref_step():  # the statements of runners/diffusion.py:519-593 (DataParallel over one GPU is a pass-through)
            a, b_ = 2 * xr.cuda() - 1, cr.cuda()
            t = torch.randint(0, 1000, (B,), device="cuda")
            e = torch.randn_like(a)
            remain = eps_loss(ref, a, t, b_, e, bd)
            c_, d_ = 2 * xf.cuda() - 1, cf.cuda()
            t = torch.randint(0, 1000, (B,), device="cuda")
            e = torch.randn_like(c_)
            xt = q_sample(c_, t, e, bd)
            out = ref(xt, t.float(), d_, mode="train")
            pseudo = ref(xt, t.float(), (d_ + 1) % 10, mode="train").detach()
            loss = torch.nn.functional.mse_loss(out, pseudo) + 1e-3 * remain
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
            for n, p in ref.named_parameters():
                if p.grad is not None:
                    p.grad *= mask["module." + n].to(p.device)  # the per-step 309 MB H2D of :589-592
            opt.step()

        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            ms_ref = timeit(ref_step, warm=2)
            key = "stock_pytorch_tf32" if tf32 else "stock_pytorch_fp32"
            res[key + "_ms_per_it"] = ms_ref
            res[key + "_its_per_s"] = 1000 / ms_ref
        res["speedup_vs_stock_fp32"] = res["stock_pytorch_fp32_ms_per_it"] / ms
        res["speedup_vs_stock_tf32"] = res["stock_pytorch_tf32_ms_per_it"] / ms
    res["mem_gb"] = torch.cuda.max_memory_allocated() / 1e9
    free, total = torch.cuda.mem_get_info()
    res["device_mem_used_gb"] = (total - free) / 1e9
    print(json.dumps(res))


if __name__ == "__main__":
    main()
