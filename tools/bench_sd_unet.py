"""Full-size Stable-Diffusion v1.4 U-Net (859.5 M parameters) forward on the engine, timed beside the stock-PyTorch statement
of the same network (oracle/sd_unet.py as the eager baseline: fp32 and TF32), with whole-network parity at full size and one
50-step DDIM sample with guidance (the shape ESD's quick_sample_till_t runs: batch 1 -> U-Net batch 2, 64x64 latents, 77x768
context).  Prints one JSON line; run on the GPU box:  python tools/bench_sd_unet.py > gpurun_out/r2_sd_unet.json"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import sd_unet as OS                                   # noqa: E402  (baseline leg + checker)
from unlearn_saliency_b200 import _lib                             # noqa: E402
from unlearn_saliency_b200.sd.engine import SDUNetEngine, sd_v1_config   # noqa: E402
from unlearn_saliency_b200.sd.sampler import EngineDDIMSampler     # noqa: E402


def synth(table, dev, seed=3):
    g = torch.Generator(device=dev).manual_seed(seed)
    P = {}
    for k, shp in table.items():
        if len(shp) == 1:
            is_norm_w = k.endswith(".weight")
            P[k] = (1.0 + 0.1 * torch.randn(shp, device=dev, generator=g)) if is_norm_w else 0.05 * torch.randn(shp, device=dev, generator=g)
        else:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            P[k] = torch.randn(shp, device=dev, generator=g) * (0.8 / fan_in ** 0.5)
    return P


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def measure(dev=None, torch_fp32=True, iters=20):
    dev = torch.device("cuda:0") if dev is None else dev
    cfg = sd_v1_config()
    n, S, L, D = 2, 64, 77, cfg["context_dim"]
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(n, 4, S, S, device=dev, generator=g)
    t = torch.tensor([481.0, 37.0], device=dev)
    ctx = torch.randn(n, L, D, device=dev, generator=g)
    out = {"workload": "SD v1.4 U-Net forward, batch 2 (uncond | cond), 64x64x4 latents, 77x768 context", "params": None}
    P = None
    ref = None
    for precision in [p for p in ("bf16", "split") if p in _lib.available_precisions()]:
        t0 = time.time()
        eng = SDUNetEngine(cfg, latent_size=S, max_batch=n, context_len=L, device=dev, precision=precision)
        if P is None:
            P = synth(eng.table, dev)
            out["params"] = eng.n_params
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
            with torch.no_grad():
                ref = OS.unet_forward(P, cfg, x, t, ctx)
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        eng.load_state_dict(P)
        eps = eng(x, t, ctx)
        torch.cuda.synchronize()
        build_s = time.time() - t0
        rel = float((eps - ref).norm() / ref.norm())
        ms = timed(lambda: eng(x, t, ctx), iters)
        sampler = EngineDDIMSampler(eng)
        torch.cuda.synchronize()
        t1 = time.time()
        z, _ = sampler.sample(S=50, batch_size=1, shape=[4, S, S], conditioning=ctx[1:], x_T=x[:1], eta=0.0,
                              unconditional_guidance_scale=3.0, unconditional_conditioning=ctx[:1])
        torch.cuda.synchronize()
        out[precision] = {"forward_ms": round(ms, 3), "rel_err_vs_fp32": rel, "launches_per_forward": eng.launches_per_forward(n),
                          "build_and_first_forward_s": round(build_s, 2), "ddim50_cfg_s": round(time.time() - t1, 3),
                          "ddim_finite": bool(torch.isfinite(z).all()),
                          "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
        del eng, sampler
        torch.cuda.empty_cache()
    with torch.no_grad():
        for name, tf32 in ((("torch_fp32", False),) if torch_fp32 else ()) + (("torch_tf32", True),):
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = tf32
            ms = timed(lambda: OS.unet_forward(P, cfg, x, t, ctx), 10)
            e = OS.unet_forward(P, cfg, x, t, ctx)
            out[name] = {"forward_ms": round(ms, 3), "rel_err_vs_fp32": float((e - ref).norm() / ref.norm())}
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True
        Ph = {k: v.to(torch.bfloat16) for k, v in P.items()}
        xb, cb = x.to(torch.bfloat16), ctx.to(torch.bfloat16)

        def bf():
            return OS.unet_forward(Ph, cfg, xb, t, cb)
        try:
            ms = timed(bf, 10)
            out["torch_bf16"] = {"forward_ms": round(ms, 3), "rel_err_vs_fp32": float((bf().float() - ref).norm() / ref.norm())}
        except Exception as ex:          # dtype plumbing of the restatement, not a product path
            out["torch_bf16"] = {"error": str(ex)[:120]}
    best = min(out[k]["forward_ms"] for k in ("torch_tf32", "torch_bf16") if "forward_ms" in out.get(k, {}))
    out["speedup_vs_best_torch_eager"] = {p: best / out[p]["forward_ms"] for p in ("bf16", "split") if p in out}
    return out


if __name__ == "__main__":
    print(json.dumps(measure()))
