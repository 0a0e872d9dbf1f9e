"""Bring-up probe of the sm_100a DDPM U-Net engine: every tape tensor that has a module-level counterpart and every
parameter gradient against the PyTorch fp32 restatement (oracle/unet.py, itself pinned to the
reference by tests/test_ddpm_cpu.py) on the same GPU.  Prints relative errors; exit code 1 if anything is off.

    python tools/gpu_unet_probe.py [tiny|small|full] [n]
"""
import sys
import time
from types import SimpleNamespace

import torch

sys.path.insert(0, ".")
from unlearn_saliency_b200.diffusion.engine import UNetEngine  # noqa: E402
from oracle.unet import ConditionalUNet, cifar10_config  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def cfg_of(name):
    if name == "full":
        return cifar10_config(dropout=0.0)
    mult, attn, size, nrb = ([1, 1], [4], 8, 1) if name == "tiny" else ([1, 2], [8], 16, 1)
    if name == "mid":
        mult, attn, size, nrb = [1, 2, 2], [16], 32, 2
    return SimpleNamespace(
        model=SimpleNamespace(type="conditional", in_channels=3, out_ch=3, ch=128, ch_mult=mult, num_res_blocks=nrb,
                              attn_resolutions=attn, dropout=0.0, resamp_with_conv=True, cond_drop_prob=0.1),
        data=SimpleNamespace(image_size=size, channels=3, n_classes=10),
        diffusion=SimpleNamespace(beta_schedule="linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000))


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-20))


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    cfg = cfg_of(which)
    torch.manual_seed(0)
    model = ConditionalUNet(cfg).cuda()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for k, p in model.named_parameters():
            if "norm" in k and k.endswith("weight"):
                p.copy_((1.0 + 0.1 * torch.randn(p.shape, generator=g)).cuda())
            elif k.endswith("bias"):
                p.copy_((0.05 * torch.randn(p.shape, generator=g)).cuda())
    model.eval()
    S = cfg.data.image_size
    x = torch.randn(n, 3, S, S, generator=g).cuda()
    t = torch.randint(0, 1000, (n,), generator=g).cuda()
    c = torch.randint(0, 10, (n,), generator=g).cuda()
    drop = (torch.rand(n, generator=g) < 0.3).cuda()
    d_eps = (torch.randn(n, 3, S, S, generator=g) / n).cuda()

    # ---- torch reference with hooks ----
    acts, grads = {}, {}

    def hook(name):
        def f(mod, inp, out):
            acts[name] = out.detach()
            out.register_hook(lambda gr, nm=name: grads.__setitem__(nm, gr.detach()))
        return f

    def reg_block(prefix, blk):
        blk.register_forward_hook(hook(prefix + ".out"))
        if hasattr(blk, "nin_shortcut"):
            blk.nin_shortcut.register_forward_hook(hook(prefix + ".sc"))

    def reg_attn(prefix, at):
        at.register_forward_hook(hook(prefix + ".out"))
        at.norm.register_forward_hook(hook(prefix + ".xn"))
        at.q.register_forward_hook(hook(prefix + ".q"))
        at.k.register_forward_hook(hook(prefix + ".k"))
        at.v.register_forward_hook(hook(prefix + ".v"))

    model.conv_in.register_forward_hook(hook("conv_in"))
    for li, lvl in enumerate(model.down):
        for i, b in enumerate(lvl.block):
            reg_block(f"down.{li}.block.{i}", b)
        for i, a in enumerate(lvl.attn):
            reg_attn(f"down.{li}.attn.{i}", a)
        if hasattr(lvl, "downsample"):
            lvl.downsample.register_forward_hook(hook(f"down.{li}.downsample"))
    reg_block("mid.block_1", model.mid.block_1)
    reg_attn("mid.attn_1", model.mid.attn_1)
    reg_block("mid.block_2", model.mid.block_2)
    for li, lvl in enumerate(model.up):
        for i, b in enumerate(lvl.block):
            reg_block(f"up.{li}.block.{i}", b)
        for i, a in enumerate(lvl.attn):
            reg_attn(f"up.{li}.attn.{i}", a)
        if hasattr(lvl, "upsample"):
            lvl.upsample.register_forward_hook(hook(f"up.{li}.upsample"))

    model.zero_grad()
    eps_ref = model(x, t.float(), c, mode="train", drop_mask=drop)
    (eps_ref * d_eps).sum().backward()
    gref = {k: p.grad.detach() if p.grad is not None else torch.zeros_like(p) for k, p in model.named_parameters()}

    # ---- engine ----
    eng = UNetEngine(cfg, max_batch=max(n, 8))
    eng.load_state_dict(model.state_dict())
    eng.eval()
    eps = eng.forward(x, t.float(), c, drop=drop, save=True)
    torch.cuda.synchronize()
    bad = 0
    print(f"== {which} n={n}: eps rel err {rel(eps, eps_ref):.4e}  (|ref| {eps_ref.norm():.3f})")
    names = eng.tensor_names()
    for name in names:
        if name in acts:
            e = eng.export(name, n)
            r = rel(e, acts[name])
            flag = "" if r < 0.03 else "  <-- ACT"
            bad += r >= 0.03
            print(f"  act  {name:28s} {r:.3e}{flag}")
    eng.backward(d_eps)
    torch.cuda.synchronize()
    for name in reversed(list(names)):
        if name in grads and not name.endswith((".xn", ".q", ".k", ".v")):
            e = eng.export(name, n, grad=True)
            r = rel(e, grads[name])
            flag = "" if r < 0.06 else "  <-- GRAD"
            bad += r >= 0.06
            print(f"  dact {name:28s} {r:.3e}{flag}")
        elif name in grads:
            e = eng.export(name, n, grad=True)
            r = rel(e, grads[name])
            flag = "" if r < 0.06 else "  <-- GRAD"
            bad += r >= 0.06
            print(f"  dact {name:28s} {r:.3e}{flag}")
    gd = eng.grad_dict()
    worst = []
    for k in gref:
        r = rel(gd[k], gref[k])
        worst.append((r, k))
        if r >= 0.08:
            bad += 1
    worst.sort(reverse=True)
    print("  param grads: worst 12 of", len(worst))
    for r, k in worst[:12]:
        print(f"    {k:48s} {r:.3e}  |ref| {gref[k].norm():.3e}" + ("  <-- PGRAD" if r >= 0.08 else ""))
    tot = torch.cat([v.reshape(-1) for v in gd.values()])
    totr = torch.cat([v.reshape(-1) for v in gref.values()])
    print(f"  whole-gradient rel err {rel(tot, totr):.4e}; finite {bool(torch.isfinite(tot).all())}")
    # timing
    for _ in range(2):
        eng.forward(x, t.float(), c, drop=drop, save=True)
        eng.backward(d_eps)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(5):
        eng.forward(x, t.float(), c, drop=drop, save=True)
        eng.backward(d_eps)
    torch.cuda.synchronize()
    print(f"  engine fwd+bwd {1e3 * (time.time() - t0) / 5:.2f} ms at n={n}")
    print("PROBE", which, "BAD" if bad else "OK", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
