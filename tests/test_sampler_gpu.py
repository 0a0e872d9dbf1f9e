"""Class-conditional sampling on the engine (diffusion/sampler.py: EngineSampler, salun_ddim_step) against the restated
reference sampler (oracle/ddpm.py:generalized_steps_conditional = DDPM/functions/denoising.py:72-95) around the pinned
torch network."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, r):
    return float((a.float().cpu() - r.float().cpu()).norm() / r.float().cpu().norm())


@pytest.mark.parametrize("precision,tol", [("split", 2e-3), ("bf16", 0.12)])
def test_generalized_sampler_matches_reference_statements(salun_ctx, precision, tol):
    from oracle import ddpm as OD
    from oracle.unet import ConditionalUNet
    from tests.golden.make_golden_ddpm import small_config, synth_weights
    from unlearn_saliency_b200 import _lib
    from unlearn_saliency_b200.diffusion.engine import UNetEngine
    from unlearn_saliency_b200.diffusion.runner import get_beta_schedule
    from unlearn_saliency_b200.diffusion.sampler import EngineSampler, timestep_sequence
    if precision not in _lib.available_precisions():
        pytest.skip("build missing")
    cfg = small_config()
    ref = ConditionalUNet(cfg)
    ref.load_state_dict(synth_weights(ref))
    ref = ref.cuda().eval()
    eng = UNetEngine(cfg, max_batch=16, ctx=salun_ctx, precision=precision).eval()
    eng.load_state_dict(ref.state_dict())
    betas = torch.from_numpy(get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)).float().cuda()
    g = torch.Generator().manual_seed(4)
    n, S = 5, cfg.data.image_size
    x = torch.randn(n, 3, S, S, generator=g).cuda()
    c = torch.randint(0, 10, (n,), generator=g).cuda()
    seq = timestep_sequence(1000, 4)                      # 4 steps: t = 750, 500, 250, 0
    assert seq == [0, 250, 500, 750]
    noises = [torch.randn(n, 3, S, S, generator=g).cuda() for _ in seq]
    torch.backends.cudnn.allow_tf32 = False
    xs_ref, x0_ref = OD.generalized_steps_conditional(x, c, seq, ref, betas, cond_scale=2.0, eta=0.7, noises=noises)
    sm = EngineSampler(eng, betas)
    xs, x0s = sm.generalized_steps_conditional(x, c, seq, cond_scale=2.0, eta=0.7, noise_fn=lambda k, like: noises[k], keep=True)
    assert len(xs) == len(xs_ref) == 5
    for k in range(1, 5):
        assert _rel(xs[k], xs_ref[k]) < tol, (k, _rel(xs[k], xs_ref[k]))
        assert _rel(x0s[k - 1], x0_ref[k - 1]) < 2 * tol
    last = sm.sample_image(x, c, 2.0, timesteps=4, eta=0.0)
    xs0, _ = OD.generalized_steps_conditional(x, c, seq, ref, betas, cond_scale=2.0, eta=0.0)
    assert _rel(last, xs0[-1]) < tol
    imgs = sm.sample_visualization(10, 10, 5, 2.0, S, timesteps=2, eta=0.0)
    assert imgs.shape == (10, 3, S, S) and float(imgs.min()) >= 0 and float(imgs.max()) <= 1
    eng.close()


def test_ddim_step_kernel_vs_torch_statements(salun_ctx):
    from unlearn_saliency_b200 import _lib
    from unlearn_saliency_b200.tail import _ptr, _stream
    g = torch.Generator().manual_seed(2)
    n, chw = 7, 3 * 16 * 16
    ec, en, xt, nz = (torch.randn(n, chw, generator=g).cuda() for _ in range(4))
    at = (torch.rand(n, generator=g) * 0.8 + 0.1).cuda()
    an = (at + (1 - at) * torch.rand(n, generator=g).cuda() * 0.9).contiguous()
    out, x0 = torch.empty_like(xt), torch.empty_like(xt)
    s, eta = 2.0, 0.6
    assert _lib.lib().salun_ddim_step(salun_ctx.handle, _ptr(ec), _ptr(en), _ptr(xt), _ptr(nz), _ptr(at), _ptr(an), s, eta, n, chw,
                                      _ptr(out), _ptr(x0), _stream(salun_ctx.device)) == 0
    a, b = at[:, None], an[:, None]
    et = (1 + s) * ec - s * en
    x0_t = (xt - et * (1 - a).sqrt()) / a.sqrt()
    c1 = eta * ((1 - a / b) * (1 - b) / (1 - a)).sqrt()
    c2 = ((1 - b) - c1 ** 2).sqrt()
    want = b.sqrt() * x0_t + c1 * nz + c2 * et
    np.testing.assert_allclose(out.cpu().numpy(), want.cpu().numpy(), rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(x0.cpu().numpy(), x0_t.cpu().numpy(), rtol=2e-6, atol=2e-6)
