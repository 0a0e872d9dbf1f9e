"""TEST INFRASTRUCTURE -- plain-PyTorch (fp32, any device) restatement of the Stable-Diffusion U-Net forward and of the DDIM
sampler, working on the reference's own parameter names.  Only tests/, __graft_entry__.smoke() and bench.py's baseline legs
import it; the product path is unlearn_saliency_b200/sd/engine.py on libsalun.

  unet_forward(params, cfg, x, t, context)   UNetModel.forward            SD/ldm/modules/diffusionmodules/openaimodel.py:814-846
                                             ResBlock._forward            :268-288 ; Downsample / Upsample :87-160
                                             SpatialTransformer, BasicTransformerBlock, CrossAttention, GEGLU
                                                                          SD/ldm/modules/attention.py:37-66,168-303
                                             timestep_embedding           SD/ldm/modules/diffusionmodules/util.py:173-197
  ddim_sample(...)                           DDIMSampler.make_schedule / ddim_sampling / p_sample_ddim
                                                                          SD/ldm/models/diffusion/ddim.py:37-100,179-362
Pinned to outputs of the unmodified reference by tests/golden/sd_unet.npz (tests/test_oracle_sd_cpu.py).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def timestep_embedding(t, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    return torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1) if dim % 2 else emb


def _resblock(P, pre, x, emb):
    h = F.conv2d(F.silu(F.group_norm(x, 32, P[pre + ".in_layers.0.weight"], P[pre + ".in_layers.0.bias"], 1e-5)),
                 P[pre + ".in_layers.2.weight"], P[pre + ".in_layers.2.bias"], padding=1)
    eo = F.linear(F.silu(emb), P[pre + ".emb_layers.1.weight"], P[pre + ".emb_layers.1.bias"])
    h = h + eo[:, :, None, None]
    h = F.conv2d(F.silu(F.group_norm(h, 32, P[pre + ".out_layers.0.weight"], P[pre + ".out_layers.0.bias"], 1e-5)),
                 P[pre + ".out_layers.3.weight"], P[pre + ".out_layers.3.bias"], padding=1)
    if pre + ".skip_connection.weight" in P:
        x = F.conv2d(x, P[pre + ".skip_connection.weight"], P[pre + ".skip_connection.bias"])
    return x + h


def _attention(P, pre, x, context, heads):
    q = F.linear(x, P[pre + ".to_q.weight"])
    k = F.linear(context, P[pre + ".to_k.weight"])
    v = F.linear(context, P[pre + ".to_v.weight"])
    b, n, c = q.shape
    d = c // heads
    sp = lambda z: z.view(b, -1, heads, d).permute(0, 2, 1, 3)
    sim = sp(q) @ sp(k).transpose(-1, -2) * d ** -0.5
    out = (sim.softmax(dim=-1) @ sp(v)).permute(0, 2, 1, 3).reshape(b, n, c)
    return F.linear(out, P[pre + ".to_out.0.weight"], P[pre + ".to_out.0.bias"])


def _transformer(P, pre, x, context, heads, depth):
    b, c, hh, ww = x.shape
    x_in = x
    x = F.group_norm(x, 32, P[pre + ".norm.weight"], P[pre + ".norm.bias"], 1e-6)
    x = F.conv2d(x, P[pre + ".proj_in.weight"], P[pre + ".proj_in.bias"])
    x = x.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
    for i in range(depth):
        t = f"{pre}.transformer_blocks.{i}"
        ln = lambda z, k: F.layer_norm(z, (c,), P[f"{t}.norm{k}.weight"], P[f"{t}.norm{k}.bias"], 1e-5)
        n1 = ln(x, 1)
        x = _attention(P, t + ".attn1", n1, n1, heads) + x
        x = _attention(P, t + ".attn2", ln(x, 2), context, heads) + x
        a, gate = F.linear(ln(x, 3), P[t + ".ff.net.0.proj.weight"], P[t + ".ff.net.0.proj.bias"]).chunk(2, dim=-1)
        x = F.linear(a * F.gelu(gate), P[t + ".ff.net.2.weight"], P[t + ".ff.net.2.bias"]) + x
    x = x.reshape(b, hh, ww, c).permute(0, 3, 1, 2)
    return F.conv2d(x, P[pre + ".proj_out.weight"], P[pre + ".proj_out.bias"]) + x_in


def unet_forward(P, cfg, x, timesteps, context):
    mc, heads, depth = cfg["model_channels"], cfg["num_heads"], cfg.get("transformer_depth", 1)
    emb = F.linear(timestep_embedding(timesteps, mc).to(P["time_embed.0.weight"].dtype), P["time_embed.0.weight"], P["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), P["time_embed.2.weight"], P["time_embed.2.bias"])

    def block(prefix, h):
        k = 0
        while True:
            pre = f"{prefix}.{k}"
            if pre + ".in_layers.0.weight" in P:
                h = _resblock(P, pre, h, emb)
            elif pre + ".norm.weight" in P:
                h = _transformer(P, pre, h, context, heads, depth)
            elif pre + ".op.weight" in P:
                h = F.conv2d(h, P[pre + ".op.weight"], P[pre + ".op.bias"], stride=2, padding=1)
            elif pre + ".conv.weight" in P:
                h = F.conv2d(F.interpolate(h, scale_factor=2, mode="nearest"), P[pre + ".conv.weight"], P[pre + ".conv.bias"], padding=1)
            else:
                return h
            k += 1

    h = F.conv2d(x, P["input_blocks.0.0.weight"], P["input_blocks.0.0.bias"], padding=1)
    hs = [h]
    i = 1
    while any(k.startswith(f"input_blocks.{i}.") for k in P):
        h = block(f"input_blocks.{i}", h)
        hs.append(h)
        i += 1
    h = block("middle_block", h)
    j = 0
    while any(k.startswith(f"output_blocks.{j}.") for k in P):
        h = block(f"output_blocks.{j}", torch.cat([h, hs.pop()], dim=1))
        j += 1
    h = F.silu(F.group_norm(h, 32, P["out.0.weight"], P["out.0.bias"], 1e-5))
    return F.conv2d(h, P["out.2.weight"], P["out.2.bias"], padding=1)


# ---------------------------------------------------------------------------------------------------------------------
def sd_alphas_cumprod(n_timestep=1000, linear_start=0.00085, linear_end=0.012):
    """make_beta_schedule("linear") of util.py:20-30 with the v1-inference.yaml values, cumulated (ddpm.py:119-131)"""
    betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2).numpy()
    return np.cumprod(1.0 - betas, axis=0)


def ddim_schedule(alphas_cumprod, ddim_num_steps, eta, num_ddpm_timesteps=1000):
    """make_ddim_timesteps("uniform") + make_ddim_sampling_parameters (util.py:56-96)"""
    c = num_ddpm_timesteps // ddim_num_steps
    steps = np.asarray(list(range(0, num_ddpm_timesteps, c))) + 1
    alphas = alphas_cumprod[steps]
    alphas_prev = np.asarray([alphas_cumprod[0]] + alphas_cumprod[steps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return steps, alphas, alphas_prev, sigmas


def ddim_sample(apply_model, cond, uncond, x_T, ddim_steps, scale, eta=0.0, t_start=-1, till_T=None, noises=None):
    """DDIMSampler.sample -> ddim_sampling -> p_sample_ddim (ddim.py:179-362) as train-esd.py's sample_model calls it:
    timesteps[:t_start], stop when index + 1 == till_T.  apply_model(x, t_long, c) -> eps."""
    ac = sd_alphas_cumprod()
    steps, alphas, alphas_prev, sigmas = ddim_schedule(ac, ddim_steps, eta)
    timesteps = steps[:t_start]
    img, b = x_T, x_T.shape[0]
    total = timesteps.shape[0]
    till = till_T if till_T is not None else 0
    for i, step in enumerate(np.flip(timesteps)):
        index = total - i - 1
        ts = torch.full((b,), int(step), device=x_T.device, dtype=torch.long)
        if uncond is None or scale == 1.0:
            e_t = apply_model(img, ts, cond)
        else:
            e_u, e_c = apply_model(torch.cat([img] * 2), torch.cat([ts] * 2), torch.cat([uncond, cond])).chunk(2)
            e_t = e_u + scale * (e_c - e_u)
        a_t, a_prev, sig = float(alphas[index]), float(alphas_prev[index]), float(sigmas[index])
        pred_x0 = (img - math.sqrt(1.0 - a_t) * e_t) / math.sqrt(a_t)
        dir_xt = math.sqrt(1.0 - a_prev - sig ** 2) * e_t
        noise = sig * (noises[i] if noises is not None else torch.zeros_like(img))
        img = math.sqrt(a_prev) * pred_x0 + dir_xt + noise
        if index + 1 == till:
            break
    return img
