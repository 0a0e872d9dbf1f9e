"""The split-precision build (libsalun_split.so: activations as bf16 hi/lo pairs, four tensor-core partial products per
multiply) against the fp32 oracle -- tight, per-tensor tolerances (the bf16 build's tolerance model
in tests/test_resnet_gpu.py does not apply: there is no bf16 rounding left to model)."""
import numpy as np
import pytest
import torch

from oracle import classification as OC

pytestmark = pytest.mark.gpu


def _rel(a, r):
    a, r = a.float().flatten().cpu(), r.float().flatten().cpu()
    return float((a - r).norm() / (r.norm() + 1e-30))


def _data(n, seed=11, size=32):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 3, size, size, generator=g), torch.randint(0, 10, (n,), generator=g)


@pytest.fixture(scope="module")
def split_engine(salun_ctx):
    from unlearn_saliency_b200 import _lib
    from unlearn_saliency_b200.engine import ResNetEngine
    if "split" not in _lib.available_precisions():
        pytest.skip("libsalun_split.so not built")
    eng = ResNetEngine("resnet18", 10, 32, max_batch=64, ctx=salun_ctx, precision="split")
    yield eng
    eng.close()


@pytest.mark.parametrize("train,sign,n", [(False, -1.0, 32), (True, 1.0, 32), (True, 1.0, 13), (False, -1.0, 64)])
def test_resnet18_split_forward_backward_vs_fp32_oracle(split_engine, train, sign, n):
    eng = split_engine
    params, buffers = OC.synth_state(10, seed=0)
    eng.load_state_dict(OC.state_dict_of(params, buffers))
    x, y = _data(n)
    b = {k: v.clone() for k, v in buffers.items()}
    lref, oref, gref = OC.loss_and_grads(params, b, x, y, train=train, sign=sign)
    eng.train(train)
    loss, logits = eng.forward_backward(x.cuda(), y.cuda(), loss_sign=sign, want_logits=True)
    torch.cuda.synchronize()
    assert _rel(logits, oref) < 2e-4, _rel(logits, oref)
    assert abs(float(loss) - float(lref)) < 1e-4 * max(1.0, abs(float(lref)))
    gd = eng.grad_dict()
    worst = max(((_rel(gd[k], r), k) for k, r in gref.items()))
    print("split build, worst per-tensor gradient error:", worst)
    flat_ref = torch.cat([r.flatten() for r in gref.values()])
    whole = _rel(torch.cat([gd[k].flatten() for k in gref]), flat_ref)
    # The yardstick is the reference's own GPU arithmetic: the same statements with torch's default TF32 convolutions.
    # Split operands carry 2^-18; what is left is the tensor cores' truncating fp32 accumulation (~2^-24 of the running sum
    # per MMA, profiles/r2_mma_accum.txt), amplified -- like every rounding on this random-weight / noise-image case -- by
    # the early layers' cancelling sums (train-mode BatchNorm most of all; conv1.weight is the worst tensor).
    torch.backends.cudnn.allow_tf32 = True
    pc = {k: v.cuda() for k, v in params.items()}
    bc = {k: v.clone().cuda() for k, v in buffers.items()}
    _, _, gtf = OC.loss_and_grads(pc, bc, x.cuda(), y.cuda(), train=train, sign=sign)
    torch.backends.cudnn.allow_tf32 = False
    whole_tf32 = _rel(torch.cat([t.flatten() for t in gtf.values()]), flat_ref)
    print(f"whole-gradient error vs fp32: split build {whole:.3e} | reference statements with TF32 convolutions {whole_tf32:.3e}")
    assert whole < 0.5 * whole_tf32, (whole, whole_tf32)
    assert whole < (3e-2 if train else 5e-3), whole
    for k, r in gref.items():
        assert _rel(gd[k], r) < (0.15 if train else 3e-2), (k, _rel(gd[k], r))
    if train:
        sd = eng.state_dict()
        np.testing.assert_allclose(sd["bn1.running_mean"].cpu().numpy(), b["bn1.running_mean"].numpy(), rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(sd["layer4.1.bn2.running_var"].cpu().numpy(), b["layer4.1.bn2.running_var"].numpy(),
                                   rtol=1e-3, atol=1e-7)


def test_resnet50_split_forward_backward(salun_ctx):
    """Bottleneck runtime (flat activations, explicit patch matrices) in the split build"""
    from unlearn_saliency_b200 import _lib
    from unlearn_saliency_b200.engine import ResNetEngine
    if "split" not in _lib.available_precisions():
        pytest.skip("libsalun_split.so not built")
    params, buffers = OC.synth_state_bottleneck(10, seed=0, depth=50, imagenet=True)
    eng = ResNetEngine("resnet50", 10, 64, max_batch=8, ctx=salun_ctx, imagenet=True, precision="split")
    eng.load_state_dict(OC.state_dict_of(params, buffers))
    x, y = _data(4, seed=21, size=64)
    for train, sign in ((False, -1.0), (True, 1.0)):
        b = {k: v.clone() for k, v in buffers.items()}
        lref, oref, gref = OC.bottleneck_loss_and_grads(params, b, x, y, train=train, sign=sign)
        # yardstick: the same statements with torch's default TF32 convolutions on this GPU
        torch.backends.cudnn.allow_tf32 = True
        pc = {k: v.cuda() for k, v in params.items()}
        bc = {k: v.clone().cuda() for k, v in buffers.items()}
        _, otf, gtf = OC.bottleneck_loss_and_grads(pc, bc, x.cuda(), y.cuda(), train=train, sign=sign)
        torch.backends.cudnn.allow_tf32 = False
        flat_ref = torch.cat([r.flatten() for r in gref.values()])
        whole_tf32 = _rel(torch.cat([t.flatten() for t in gtf.values()]), flat_ref)
        eng.train(train)
        loss, logits = eng.forward_backward(x.cuda(), y.cuda(), loss_sign=sign, want_logits=True)
        gd = eng.grad_dict()
        whole = _rel(torch.cat([gd[k].flatten() for k in gref]), flat_ref)
        print(f"resnet50 {'train' if train else 'eval'}: logits {_rel(logits, oref):.3e} (TF32 {_rel(otf, oref):.3e}); "
              f"whole gradient {whole:.3e} (TF32 {whole_tf32:.3e})")
        assert _rel(logits, oref) < max(2e-3, _rel(otf, oref)), _rel(logits, oref)
        assert whole < whole_tf32, (whole, whole_tf32)            # closer to fp32 than the reference's own GPU arithmetic
    eng.close()


def test_unet_split_matches_reference_golden(salun_ctx):
    """DDPM U-Net engine, split build, against the reference-generated golden (tests/golden/ddpm_small.npz: channel
    change, 384-wide concat, down / up sampling, 64-token attention) and the fp32 torch statements."""
    import os
    from oracle import ddpm as OD
    from tests.golden.make_golden_ddpm import inputs, small_config, synth_weights
    from unlearn_saliency_b200 import _lib
    from unlearn_saliency_b200.diffusion.engine import UNetEngine
    from unlearn_saliency_b200.diffusion.runner import get_beta_schedule
    from oracle.unet import ConditionalUNet
    if "split" not in _lib.available_precisions():
        pytest.skip("libsalun_split.so not built")
    cfg = small_config()
    ref = ConditionalUNet(cfg)
    ref.load_state_dict(synth_weights(ref))
    ref.eval()
    x0, e, t, c = inputs(seed=2, n=6, size=cfg.data.image_size)
    betas = torch.from_numpy(get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)).float()
    ref.zero_grad()
    lref = OD.eps_loss(ref, x0, t, c, e, betas, cond_drop_prob=0.0)
    lref.backward()
    eng = UNetEngine(cfg, max_batch=8, ctx=salun_ctx, precision="split").eval()
    eng.load_state_dict(ref.state_dict())
    n = x0.shape[0]
    xt = OD.q_sample(x0, t, e, betas).cuda().contiguous()
    eps = eng.forward(xt, t.float().cuda(), c.cuda(), drop=torch.zeros(n, dtype=torch.uint8, device="cuda"), save=True)
    with torch.no_grad():
        eps_ref = ref(OD.q_sample(x0, t, e, betas), t.float(), c, mode="train", cond_drop_prob=0.0)
    assert _rel(eps, eps_ref) < 3e-4, _rel(eps, eps_ref)
    ed = e.cuda()
    eng.backward(((-2.0 / n) * (ed - eps)).contiguous())
    torch.cuda.synchronize()
    gd = eng.grad_dict()
    worst = (0.0, None)
    for k, p in ref.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        if float(g.norm()) < 1e-7 or k.endswith(".k.bias"):
            continue   # softmax is invariant to a constant added to every key: d/d(k.bias) is exactly 0, numerically noise
        r = _rel(gd[k], g)
        worst = max(worst, (r, k))
    print("unet split worst per-tensor gradient error", worst)
    mine = torch.cat([gd[k].reshape(-1).cpu() for k, _ in ref.named_parameters()])
    want = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for _, p in ref.named_parameters()])
    assert _rel(mine, want) < 2e-3, _rel(mine, want)
    assert worst[0] < 2e-2, worst
    eng.close()
