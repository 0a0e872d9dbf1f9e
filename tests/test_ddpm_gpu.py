"""DDPM SalUn loop bodies with the fused sm_100a tail against the reference's statements run with stock PyTorch
(same model weights, same externally drawn t / e / class-dropout decisions): DDPM/runners/diffusion.py:519-593, 959-1039."""
import copy

import numpy as np
import pytest
import torch

from oracle import tail as OT
from tests.golden.make_golden_ddpm import synth_weights, tiny_config

pytestmark = pytest.mark.gpu


torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _models():
    from oracle.unet import ConditionalUNet
    m = ConditionalUNet(tiny_config())
    m.load_state_dict(synth_weights(m))
    return m.cuda(), copy.deepcopy(m).cuda()


def _draw(seed, n=8, size=8):
    g = torch.Generator().manual_seed(seed)
    r = dict(x_r=torch.rand(n, 3, size, size, generator=g), c_r=torch.randint(1, 10, (n,), generator=g),
             x_f=torch.rand(n, 3, size, size, generator=g), c_f=torch.zeros(n, dtype=torch.long),
             t_r=torch.randint(0, 1000, (n,), generator=g), e_r=torch.randn(n, 3, size, size, generator=g),
             t_f=torch.randint(0, 1000, (n,), generator=g), e_f=torch.randn(n, 3, size, size, generator=g),
             drop_r=torch.rand(n, generator=g) < 0.1, drop_f=torch.rand(n, generator=g) < 0.1,
             drop_p=torch.rand(n, generator=g) < 0.1)
    return r


def test_saliency_unlearn_step_matches_reference_statements(salun_ctx):
    from tests.ddpm_torch_helper import DDPMUnlearner
    from unlearn_saliency_b200.diffusion.runner import eps_loss, get_beta_schedule, q_sample
    mine, ref = _models()
    betas = torch.from_numpy(get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)).float()
    g = torch.Generator().manual_seed(3)
    mask = {"module." + n: (torch.rand(p.shape, generator=g) < 0.5).to(torch.int64) for n, p in ref.named_parameters()}
    un = DDPMUnlearner(mine, betas, lr=1e-4, grad_clip=1.0, mask=mask, ctx=salun_ctx)
    opt = torch.optim.Adam(ref.parameters(), lr=1e-4, weight_decay=0.0, betas=(0.9, 0.999), amsgrad=False, eps=1e-8)
    bd = betas.cuda()
    p0 = {n: p.detach().clone() for n, p in ref.named_parameters()}
    for step in range(3):
        r = _draw(10 + step)
        loss_mine = un.saliency_unlearn_step(r["x_r"], r["c_r"], r["x_f"], r["c_f"], alpha=1e-3, method="rl",
                                             rng={k: r[k] for k in ("t_r", "e_r", "t_f", "e_f", "drop_r", "drop_f", "drop_p")})
        # the reference's statements (runners/diffusion.py:523-593) with stock PyTorch
        ref.train()
        xr, xf = 2 * r["x_r"].cuda() - 1, 2 * r["x_f"].cuda() - 1
        remain = eps_loss(ref, xr, r["t_r"].cuda(), r["c_r"].cuda(), r["e_r"].cuda(), bd, drop_mask=r["drop_r"].cuda())
        xt = q_sample(xf, r["t_f"].cuda(), r["e_f"].cuda(), bd)
        out = ref(xt, r["t_f"].cuda().float(), r["c_f"].cuda(), mode="train", drop_mask=r["drop_f"].cuda())
        pseudo = ref(xt, r["t_f"].cuda().float(), (r["c_f"].cuda() + 1) % 10, mode="train", drop_mask=r["drop_p"].cuda()).detach()
        loss = torch.nn.functional.mse_loss(out, pseudo) + 1e-3 * remain
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
        for n, p in ref.named_parameters():
            if p.grad is not None:
                p.grad *= mask["module." + n].to(p.device)
        opt.step()
        assert abs(float(loss_mine) - float(loss)) <= 1e-4 * abs(float(loss)) + 1e-6
    # Both arms run the same PyTorch forward/backward; cuDNN/atomics make two runs differ in the last bits, and Adam's
    # m/sqrt(v) normalisation turns a sign flip of a ~0 gradient into a full +-lr step.  So: every element within
    # 3 steps * lr, and all but a vanishing fraction within fp32 rounding of the reference statements.
    tot = bad = 0
    for (n, p), (_, q) in zip(ref.named_parameters(), mine.named_parameters()):
        d = (q - p).abs()
        assert float(d.max()) <= 3.1e-4, (n, float(d.max()))
        bad += int((d > 2e-6 + 1e-4 * p.abs()).sum())
        tot += p.numel()
        m = mask["module." + n].cuda()
        assert torch.equal(q[m == 0], p0[n][m == 0])
    assert bad / tot < 2e-3, bad / tot


def test_generate_mask_matches_reference_statements(salun_ctx, tmp_path):
    from tests.ddpm_torch_helper import DDPMUnlearner
    from unlearn_saliency_b200.diffusion.runner import get_beta_schedule, q_sample
    mine, ref = _models()
    betas = torch.from_numpy(get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)).float()
    un = DDPMUnlearner(mine, betas, ctx=salun_ctx)
    bd = betas.cuda()
    grads = {n: 0 for n, _ in ref.named_parameters()}
    ref.eval()
    for b in range(2):
        r = _draw(30 + b)
        un.generate_mask_batch(r["x_f"], r["c_f"], cond_scale=2.0, t=r["t_f"], e=r["e_f"])
        x = 2 * r["x_f"].cuda() - 1
        xt = q_sample(x, r["t_f"].cuda(), r["e_f"].cuda(), bd)
        out = ref(xt, r["t_f"].cuda().float(), r["c_f"].cuda(), cond_scale=2.0, mode="test")
        loss = (r["e_f"].cuda() - out).square().sum(dim=(1, 2, 3)).mean(dim=0)       # :980
        ref.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)                          # :985-990
        for n, p in ref.named_parameters():
            if p.grad is not None:
                grads[n] = grads[n] + p.grad.data.cpu()                                # :992-996
    path = str(tmp_path / "mask" / "0" / "with_0.5.pt")
    un.finish_mask(path, 0.5)
    m = torch.load(path)
    assert list(m.keys()) == ["module." + n for n in grads]
    flat = torch.cat([torch.as_tensor(g).abs().flatten() if not isinstance(g, int) else torch.zeros(0) for g in grads.values()])
    mine_flat = torch.cat([v.flatten() for v in m.values()]).numpy()
    k = int(mine_flat.size * 0.5)
    assert mine_flat.sum() == k
    ref_mask = OT.topk_mask_argsort(flat.numpy(), k)
    assert (mine_flat != ref_mask).mean() < 2e-3   # identical up to elements whose |g| differs in the last bits
