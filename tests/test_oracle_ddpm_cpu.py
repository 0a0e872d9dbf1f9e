"""oracle/ddpm.py (restated DDPM SalUn loop bodies) against the golden outputs of the UNMODIFIED reference
(tests/golden/ddpm_tiny.npz, made by tests/golden/make_golden_ddpm.py from /root/reference/DDPM).  CPU only."""
import os

import numpy as np
import torch

from oracle import ddpm as OD
from tests.golden.make_golden_ddpm import inputs, sample_idx, synth_weights, tiny_config
from oracle.unet import ConditionalUNet

G = os.path.join(os.path.dirname(__file__), "golden", "ddpm_tiny.npz")


def _model():
    m = ConditionalUNet(tiny_config())
    m.load_state_dict(synth_weights(m))
    return m


def test_eps_loss_and_gradients_equal_reference():
    z = np.load(G)
    m = _model().eval()
    x0, e, t, c = inputs()
    betas = torch.from_numpy(z["betas"])
    m.zero_grad()
    loss = OD.eps_loss(m, x0, t, c, e, betas, cond_drop_prob=0.0)
    loss.backward()
    np.testing.assert_allclose(loss.item(), z["loss"], rtol=1e-5)
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in m.parameters()]
    np.testing.assert_allclose(np.array([g.norm().item() for g in grads]), z["gnorm"], rtol=1e-3, atol=1e-6)
    gs = np.concatenate([g.flatten()[sample_idx(g.numel())].numpy() for g in grads])
    np.testing.assert_allclose(gs, z["gsample"], rtol=1e-2, atol=1e-4)


def test_generate_mask_batch_uses_cfg_output_and_clips():
    """the mask-generation loss is built on the reference's "test"-mode output (golden eps_test) and the accumulated
    gradient is the clipped one (norm <= 1)"""
    z = np.load(G)
    m = _model()
    x0, e, t, c = inputs()
    betas = torch.from_numpy(z["betas"])
    x01 = (x0 + 1) / 2                      # generate_mask_batch applies 2x - 1 itself
    grads = {}
    loss = OD.generate_mask_batch(m, x01, c, t, e, betas, grads, cond_scale=2.0)
    ref = ((e - torch.from_numpy(z["eps_test"])) ** 2).sum(dim=(1, 2, 3)).mean()
    np.testing.assert_allclose(loss.item(), ref.item(), rtol=1e-4)
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values()))
    assert float(total) <= 1.0 + 1e-4
    assert set(grads) == {k for k, _ in m.named_parameters()}


def test_unlearn_step_masks_and_steps_like_adam():
    z = np.load(G)
    m = _model()
    p0 = {k: p.detach().clone() for k, p in m.named_parameters()}
    g = torch.Generator().manual_seed(4)
    mask = {k: (torch.rand(p.shape, generator=g) < 0.5).to(torch.int64) for k, p in m.named_parameters()}
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    n, S = 4, 8
    r = dict(x_r=torch.rand(n, 3, S, S, generator=g), c_r=torch.randint(1, 10, (n,), generator=g),
             x_f=torch.rand(n, 3, S, S, generator=g), c_f=torch.zeros(n, dtype=torch.long),
             t_r=torch.randint(0, 1000, (n,), generator=g), e_r=torch.randn(n, 3, S, S, generator=g),
             t_f=torch.randint(0, 1000, (n,), generator=g), e_f=torch.randn(n, 3, S, S, generator=g),
             drop_r=torch.rand(n, generator=g) < 0.1, drop_f=torch.rand(n, generator=g) < 0.1,
             drop_p=torch.rand(n, generator=g) < 0.1)
    loss, norm, raw = OD.saliency_unlearn_step(m, opt, mask, r, torch.from_numpy(z["betas"]), alpha=1e-3, method="rl")
    assert torch.isfinite(loss) and float(norm) > 0
    moved = tot = 0
    for k, p in m.named_parameters():
        mk = mask[k].bool()
        assert torch.equal(p.detach()[~mk], p0[k][~mk])                 # masked-out coordinates keep theta_0
        sel = mk & (raw[k].abs() > 1e-7)                                # |g| >> eps: the first Adam step is lr * sign(g)
        moved += int((p.detach() != p0[k])[sel].sum())
        tot += int(sel.sum())
    assert moved / tot > 0.99, moved / tot
