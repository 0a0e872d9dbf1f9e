"""SD loop mirrors (unlearn_saliency_b200/sd: train_esd / certain_label / generate_mask on the fused sm_100a tail) against
the reference's statements (SD/train-scripts/train-esd.py:268-323, random_label.py:77-139, generate_mask.py:33-108) in
stock PyTorch, around a small stand-in for LatentDiffusion (the SD stack itself is not importable):
same parameter-name patterns (attn1 / attn2 / time_embed / out.), same call surface."""
import copy
import zlib

import numpy as np
import pytest
import torch
from torch import nn

from oracle import tail as OT

pytestmark = pytest.mark.gpu


class _Block(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.attn1 = nn.Conv2d(c, c, 1)
        self.attn2 = nn.Linear(16, c)
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x, ctx):
        h = x + self.attn1(x)
        h = h + self.attn2(ctx.mean(dim=1))[:, :, None, None]
        return h + torch.tanh(self.conv(h))


class _UNet(nn.Module):
    def __init__(self, c=16):
        super().__init__()
        self.time_embed = nn.Linear(1, c)
        self.input_blocks = nn.ModuleList([nn.Conv2d(4, c, 3, padding=1), _Block(c)])
        self.output_blocks = nn.ModuleList([_Block(c)])
        self.out = nn.Sequential(nn.Conv2d(c, 4, 3, padding=1))

    def forward(self, x, t, ctx):
        h = self.input_blocks[0](x) + self.time_embed(t.float()[:, None] / 1000.0)[:, :, None, None]
        h = self.input_blocks[1](h, ctx)
        h = self.output_blocks[0](h, ctx)
        return self.out(h)


class _LDM(nn.Module):
    """the slice of LatentDiffusion's surface the SalUn scripts call"""
    num_timesteps, first_stage_key = 1000, "jpg"

    def __init__(self):
        super().__init__()
        self.model = nn.Module()
        self.model.diffusion_model = _UNet()
        betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000) ** 2
        self.register_buffer("abar", (1 - betas).cumprod(0))

    def get_learned_conditioning(self, prompts):
        dev = self.abar.device
        out = []
        for p in prompts:
            g = torch.Generator().manual_seed(zlib.crc32(p.encode()))
            out.append(torch.randn(77, 16, generator=g))
        return torch.stack(out).to(dev)

    def apply_model(self, x, t, cond):
        return self.model.diffusion_model(x, t, cond)

    def get_input(self, batch, key):
        img = batch[key].permute(0, 3, 1, 2)
        z = torch.nn.functional.avg_pool2d(img, 8)
        z = torch.cat([z, z[:, :1]], dim=1)
        return z, self.get_learned_conditioning(batch["txt"])

    def q_sample(self, x_start, t, noise):
        a = self.abar[t].view(-1, 1, 1, 1)
        return a.sqrt() * x_start + (1 - a).sqrt() * noise

    def shared_step(self, batch):
        z, c = self.get_input(batch, self.first_stage_key)
        g = torch.Generator(device=z.device).manual_seed(11)
        t = torch.randint(0, 1000, (z.shape[0],), device=z.device, generator=g)
        noise = torch.randn(z.shape, device=z.device, generator=g)
        return torch.nn.functional.mse_loss(self.apply_model(self.q_sample(z, t, noise), t, c), noise), {}


def _mask_for(model, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {n: (torch.rand(p.shape, generator=g) < 0.5).to(torch.int64) for n, p in model.model.diffusion_model.named_parameters()}


def _close(mine, ref, p0, mask, selected, lr, steps):
    tot = bad = 0
    for (n, p), (_, q) in zip(ref.model.diffusion_model.named_parameters(), mine.model.diffusion_model.named_parameters()):
        keep = mask[n].cuda().bool() if n in selected else torch.zeros_like(p, dtype=torch.bool)
        assert torch.equal(q[~keep], p0[n][~keep]), n          # masked-out / unselected coordinates never move
        d = (q - p).abs()
        assert float(d.max()) <= steps * lr * 1.05, (n, float(d.max()))
        bad += int((d > 2e-7 + 1e-4 * p.abs()).sum())
        tot += p.numel()
    assert bad / tot < 5e-3, bad / tot   # Adam turns a last-bit gradient difference of a ~0 gradient into a full +-lr step


def test_esd_iterations_match_reference_statements(salun_ctx):
    from unlearn_saliency_b200.sd import SDTail, esd_iteration, select_parameters
    torch.manual_seed(0)
    mine = _LDM().cuda()
    ref, frozen = copy.deepcopy(mine), copy.deepcopy(mine)
    mask = _mask_for(mine)
    lr, method = 1e-4, "xattn"
    names = [n for n, _ in mine.model.diffusion_model.named_parameters()]
    selected = set(select_parameters(names, method))
    assert selected and len(selected) < len(names)
    tail = SDTail(mine, lr=lr, train_method=method, mask=mask, ctx=salun_ctx)
    opt = torch.optim.Adam([p for n, p in ref.model.diffusion_model.named_parameters() if n in selected], lr=lr)
    p0 = {n: p.detach().clone() for n, p in ref.model.diffusion_model.named_parameters()}
    sample = lambda emb, s, code, t: code * 0.5 + emb.mean() * 0.1      # stands in for the DDIM partial sampler (no grad)
    for it in range(3):
        g = torch.Generator().manual_seed(40 + it)
        rng = dict(t_enc=torch.randint(50, (1,), generator=g).cuda(), start_code=torch.randn(1, 4, 8, 8, generator=g).cuda())
        rng["t_enc_ddpm"] = torch.randint(0, 1000, (1,), generator=g).cuda()
        loss_m = esd_iteration(mine, frozen, sample, tail, "Van Gogh", 3.0, 1.0, image_size=64, rng=rng)
        # reference statements, train-esd.py:268-323
        emb_0, emb_p = ref.get_learned_conditioning([""]), ref.get_learned_conditioning(["Van Gogh"])
        opt.zero_grad()
        with torch.no_grad():
            z = sample(emb_p, 3.0, rng["start_code"], int(rng["t_enc"]))
            e_0 = frozen.apply_model(z, rng["t_enc_ddpm"], emb_0)
            e_p = frozen.apply_model(z, rng["t_enc_ddpm"], emb_p)
        e_n = ref.apply_model(z, rng["t_enc_ddpm"], emb_p)
        loss = torch.nn.functional.mse_loss(e_n, e_0 - (1.0 * (e_p - e_0)))
        loss.backward()
        for n, p in ref.named_parameters():
            if p.grad is not None:
                p.grad *= mask[n.split("model.diffusion_model.")[-1]].to("cuda")
        opt.step()
        assert abs(float(loss_m) - float(loss)) <= 1e-5 * abs(float(loss)) + 1e-7
    _close(mine, ref, p0, mask, selected, lr, 3)


def test_certain_label_and_generate_mask(salun_ctx, tmp_path, monkeypatch):
    from unlearn_saliency_b200.sd import certain_label, generate_mask
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(1)
    mine = _LDM().cuda()
    ref = copy.deepcopy(mine)
    g = torch.Generator().manual_seed(3)
    desc = [f"an image of class {k}" for k in range(10)]
    forget = [(torch.rand(2, 3, 64, 64, generator=g), torch.zeros(2, dtype=torch.long)) for _ in range(2)]
    remain = [(torch.rand(2, 3, 64, 64, generator=g), torch.randint(1, 10, (2,), generator=g)) for _ in range(2)]
    # ---- mask generation: file format, and selection == argsort definition on the accumulated |gradient| ----
    info = generate_mask(0, 7.5, 2, 1, 1e-5, None, None, None, "cuda", image_size=64, model=mine, loader=forget,
                         descriptions=desc, ctx=salun_ctx)
    m = torch.load(str(tmp_path / "mask" / "0" / "with_0.5.pt"))
    names = [n for n, _ in ref.model.diffusion_model.named_parameters()]
    assert list(m.keys()) == names and all(v.dtype == torch.int64 and v.device.type == "cpu" for v in m.values())
    flat = torch.cat([v.flatten() for v in m.values()]).numpy()
    assert flat.sum() == int(flat.size * 0.5)
    # ---- certain_label: one epoch against the reference statements (random_label.py:77-139) ----
    mask = m
    p0 = {n: p.detach().clone() for n, p in ref.model.diffusion_model.named_parameters()}
    lr = 1e-4
    torch.manual_seed(5)
    certain_label(0, "full", 0.5, 2, 1, lr, None, None, str(tmp_path / "mask" / "0" / "with_0.5.pt"), None, "cuda",
                  image_size=64, model=mine, loaders=(remain, forget), descriptions=desc, ctx=salun_ctx)
    torch.manual_seed(5)
    opt = torch.optim.Adam(ref.model.diffusion_model.parameters(), lr=lr)
    ref.train()
    for (fi, fl), (ri, rl) in zip(forget, remain):
        opt.zero_grad()
        fi, ri = fi.cuda(), ri.cuda()
        remain_loss = ref.shared_step({"jpg": ri.permute(0, 2, 3, 1), "txt": [desc[int(l)] for l in rl]})[0]
        f_in, f_emb = ref.get_input({"jpg": fi.permute(0, 2, 3, 1), "txt": [desc[int(l)] for l in fl]}, "jpg")
        p_in, p_emb = ref.get_input({"jpg": fi.permute(0, 2, 3, 1), "txt": [desc[1] for _ in fl]}, "jpg")
        t = torch.randint(0, 1000, (f_in.shape[0],), device="cuda").long()
        noise = torch.randn_like(f_in)
        f_out = ref.apply_model(ref.q_sample(f_in, t, noise), t, f_emb)
        p_out = ref.apply_model(ref.q_sample(p_in, t, noise), t, p_emb).detach()
        loss = torch.nn.functional.mse_loss(f_out, p_out) + 0.5 * remain_loss
        loss.backward()
        for n, p in ref.named_parameters():
            if p.grad is not None:
                p.grad *= mask[n.split("model.diffusion_model.")[-1]].to("cuda")
        opt.step()
    _close(mine, ref, p0, mask, set(names), lr, 2)
