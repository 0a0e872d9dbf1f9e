"""One ESD iteration (SD/train-scripts/train-esd.py:268-323) at Stable-Diffusion v1.4 size on one B200, BASELINE.json configs[4]:
the reference's statements with every U-Net evaluation in stock PyTorch eager (torch defaults: TF32 convolutions, fp32 matmuls) against the same loop with the
no-grad evaluations on the engine (train_esd(engine=...): DDIM partial sampling + the frozen model's e_0 / e_p; the one
autograd pass stays PyTorch in both arms).  The LatentDiffusion object is a stand-in (random-init U-Net with the reference's
parameter names over the torch restatement oracle/sd_unet.py; deterministic fake text embeddings) -- the SD stack, CLIP and
the checkpoint are not on the box.  t_enc is fixed to ddim_steps / 2 (the mean of the reference's uniform draw) so both arms
do the same work.  Prints one JSON line."""
import json
import os
import sys
import time
import zlib

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sd_unet as OS                                            # noqa: E402  (stock-PyTorch arm / stand-in)
from tools.bench_sd_unet import synth                                       # noqa: E402
from unlearn_saliency_b200.sd import SDTail, esd_iteration                  # noqa: E402
from unlearn_saliency_b200.sd.engine import sd_unet_param_table, sd_v1_config   # noqa: E402
from unlearn_saliency_b200.sd.loops import engine_passes                    # noqa: E402
from unlearn_saliency_b200.tail import SalunContext                         # noqa: E402


class TorchUNet(nn.Module):
    def __init__(self, P, cfg):
        super().__init__()
        self.cfg, self.num_heads = cfg, cfg["num_heads"]
        for k, v in P.items():
            m, parts = self, k.split(".")
            for p in parts[:-1]:
                if p not in m._modules:
                    m.add_module(p, nn.Module())
                m = m._modules[p]
            m.register_parameter(parts[-1], nn.Parameter(v.clone()))

    def forward(self, x, t, context):
        return OS.unet_forward(dict(self.named_parameters()), self.cfg, x, t.float(), context)


class LDM(nn.Module):
    def __init__(self, P, cfg):
        super().__init__()
        self.model = nn.Module()
        self.model.diffusion_model = TorchUNet(P, cfg)
        self.D = cfg["context_dim"]

    def get_learned_conditioning(self, prompts):
        out = []
        for p in prompts:
            g = torch.Generator().manual_seed(zlib.crc32(p.encode()))
            out.append(torch.randn(77, self.D, generator=g))
        return torch.stack(out).cuda()

    def apply_model(self, x, t, cond):
        return self.model.diffusion_model(x, t, cond)


def main():
    dev = torch.device("cuda:0")
    cfg = sd_v1_config()
    steps, iters, method = 50, int(sys.argv[1]) if len(sys.argv) > 1 else 4, "xattn"
    P = synth(sd_unet_param_table(cfg), dev)
    ctx = SalunContext(dev)
    out = {"workload": "ESD iteration, SD v1.4 U-Net (859.5 M parameters), 512x512 (64x64 latents), ddim_steps 50, t_enc 25, "
                       "train_method xattn, batch 1 (guided sampling = U-Net batch 2)", "iterations_timed": iters}
    for arm in ("torch_eager", "engine_bf16", "engine_split"):
        model, frozen = LDM(P, cfg).cuda(), LDM(P, cfg).cuda()
        for p in frozen.parameters():
            p.requires_grad_(False)
        tail = SDTail(model, lr=1e-5, train_method=method, ctx=ctx)
        uncond = model.get_learned_conditioning([""])
        if arm == "torch_eager":
            fro = frozen
            sample_fn = lambda emb, s, code, t: OS.ddim_sample(model.apply_model, emb, uncond, code, steps, s, till_T=t)
        else:
            fro, sample_fn = engine_passes(model, frozen, tail, arm.split("_")[1], image_size=512, ddim_steps=steps, ctx=ctx)

        def one(i):
            g = torch.Generator().manual_seed(i)
            rng = dict(t_enc=torch.tensor([steps // 2], device=dev), start_code=torch.randn(1, 4, 64, 64, generator=g).to(dev),
                       t_enc_ddpm=torch.tensor([500], device=dev))
            return esd_iteration(model, fro, sample_fn, tail, "Van Gogh", 3.0, 1.0, image_size=512, ddim_steps=steps, rng=rng)

        one(0)
        torch.cuda.synchronize()
        t0 = time.time()
        losses = [float(one(1 + i)) for i in range(iters)]
        torch.cuda.synchronize()
        ms = (time.time() - t0) * 1e3 / iters
        out[arm] = {"ms_per_iteration": round(ms, 1), "iterations_per_s": round(1e3 / ms, 3), "last_loss": losses[-1],
                    "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}
        del model, frozen, tail, fro, sample_fn
        torch.cuda.empty_cache()
    base = out["torch_eager"]["ms_per_iteration"]
    out["speedup_vs_torch_eager"] = {k: round(base / out[k]["ms_per_iteration"], 2) for k in ("engine_bf16", "engine_split")}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
