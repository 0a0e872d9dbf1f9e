"""Golden vectors for the SD U-Net forward from the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_sd.py  ->  tests/golden/sd_unet.npz

``UNetModel`` (SD/ldm/modules/diffusionmodules/openaimodel.py:413-846) imports with a 3-line omegaconf stub (SURVEY.md
Appendix C).  Weights are synthetic and seeded (the reference zero-initialises proj_out / out_layers.3 / out.2, which would
make every golden output trivially zero); inputs are seeded; the file holds eps = model(x, t, context) for
  "a": two levels (64 / 128 channels), attention at both, 2 heads (d = 32 / 64), 77 x 128 context, 16 x 16 latents, batch 2
  "b": one level at 320 channels, 8 heads of width 40 (the SD v1.4 head geometry), 77 x 64 context, 8 x 8 latents, batch 3
plus the named_parameters() keys of the SD v1.4 configuration."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CONFIGS = {
    "a": dict(cfg=dict(in_channels=4, out_channels=4, model_channels=64, attention_resolutions=[2, 1], num_res_blocks=1,
                       channel_mult=[1, 2], num_heads=2, transformer_depth=1, context_dim=128), latent=16, n=2, ctx_len=77),
    "b": dict(cfg=dict(in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[1], num_res_blocks=1,
                       channel_mult=[1], num_heads=8, transformer_depth=1, context_dim=64), latent=8, n=3, ctx_len=77),
}


def sd_synth_weights(table, seed=0):
    """{name: tensor}: norms ~ 1 + 0.1 N(0,1) / 0.1 N(0,1), weights N(0, 1/fan_in) x 0.8, biases 0.05 N(0,1)"""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in table.items():
        if len(shp) == 1:
            is_norm = any(s in k for s in (".norm", "in_layers.0", "out_layers.0", "out.0"))
            r = torch.randn(shp, generator=g)
            sd[k] = (1.0 + 0.1 * r) if (is_norm and k.endswith(".weight")) else (0.1 * r if is_norm else 0.05 * r)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            sd[k] = torch.randn(shp, generator=g) * (0.8 / fan_in ** 0.5)
    return sd


def sd_inputs(tag):
    c = CONFIGS[tag]
    g = torch.Generator().manual_seed(100 + ord(tag))
    n, S = c["n"], c["latent"]
    x = torch.randn(n, 4, S, S, generator=g)
    t = torch.randint(0, 1000, (n,), generator=g).float()
    ctx = torch.randn(n, c["ctx_len"], c["cfg"]["context_dim"], generator=g)
    return x, t, ctx


def import_unet():
    sys.path.insert(0, "/root/reference/SD")
    oc, lc = types.ModuleType("omegaconf"), types.ModuleType("omegaconf.listconfig")

    class ListConfig(list):
        pass

    lc.ListConfig = ListConfig
    oc.listconfig = lc
    sys.modules["omegaconf"], sys.modules["omegaconf.listconfig"] = oc, lc
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    return UNetModel


def main():
    torch.set_num_threads(8)
    UNetModel = import_unet()
    out = {}
    for tag, c in CONFIGS.items():
        m = UNetModel(image_size=32, use_spatial_transformer=True, use_checkpoint=False, legacy=False, **c["cfg"]).eval()
        table = {k: tuple(p.shape) for k, p in m.named_parameters()}
        m.load_state_dict(sd_synth_weights(table, seed=7), strict=True)
        x, t, ctx = sd_inputs(tag)
        with torch.no_grad():
            eps = m(x, timesteps=t, context=ctx)
        out[f"{tag}_eps"] = eps.numpy()
        out[f"{tag}_keys"] = np.array(list(table.keys()))
        print(tag, "params", sum(p.numel() for p in m.parameters()), "eps rms", float(eps.square().mean().sqrt()))
    with torch.device("meta"):
        full = UNetModel(image_size=32, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
                         num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True,
                         transformer_depth=1, context_dim=768, use_checkpoint=True, legacy=False)
    out["v14_keys"] = np.array([k for k, _ in full.named_parameters()])
    out["v14_numel"] = np.int64(sum(p.numel() for p in full.parameters()))
    np.savez_compressed(os.path.join(HERE, "sd_unet.npz"), **out)


if __name__ == "__main__":
    main()
