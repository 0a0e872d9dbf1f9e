"""sm_100a ResNet engine (through the C ABI) against the oracle (oracle/classification.py, torch fp32 on the CPU,
itself pinned to the unmodified reference by tests/test_oracle_golden.py).

Precision contract (DESIGN.md "precision"): tensor-core operands, raw conv outputs, activations and their gradients
are stored in bf16 (8-bit mantissa); accumulation, BatchNorm statistics, master weights, weight gradients and the
optimizer are fp32.  Two references, both from oracle/classification.py:
  * emulate_bf16=False -- the reference's fp32 chain (F).
  * emulate_bf16=True  -- the same chain in fp32 arithmetic with a bf16 rounding at exactly the engine's storage
    points (E): the plain-PyTorch reference OF THE OP THE KERNELS IMPLEMENT.
On this deliberately harsh case (random weights, iid-noise images, <= 32 samples, BatchNorm backward cancelling the
dominant gradient component) bf16 rounding noise is chaotically amplified: E itself sits 1-18% (eval) / 15-36% (train)
away from F, and two bf16 evaluations that differ only in summation order sit equally far from each other.  A
per-element tolerance against E or F would therefore either be vacuous or flaky; the stated tolerance is
    err(engine, F) <= 1.3 * err(E, F) + 0.02   per parameter-gradient tensor (relative L2), and
    cos(engine, F) >= cos(E, F) - 0.03,
i.e. the kernels are as close to the fp32 reference as a faithful bf16 evaluation of the reference is.  The last
layers, where no amplification has happened yet, are additionally held to 3% of E.  Logits: 0.06 + 2% of F.
Masked-out weights are bit-identical in all cases; the tail kernels are bit-exact (tests/test_tail_gpu.py).
"""
import math

import numpy as np
import pytest
import torch

from oracle import classification as OC

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(salun_ctx):
    from unlearn_saliency_b200.engine import ResNetEngine
    eng = ResNetEngine("resnet18", 10, 32, max_batch=64, ctx=salun_ctx)
    yield eng
    eng.close()


def _load(engine, seed=0):
    params, buffers = OC.synth_state(10, seed=seed)
    engine.load_state_dict(OC.state_dict_of(params, buffers))
    return params, buffers


def _data(n, seed=11):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 3, 32, 32, generator=g), torch.randint(0, 10, (n,), generator=g)


def _rel_cos(a, r):
    a, r = a.float().flatten(), r.float().flatten()
    return float((a - r).norm() / (r.norm() + 1e-12)), float(torch.dot(a, r) / (a.norm() * r.norm() + 1e-20))


TIGHT = ("fc.weight", "fc.bias", "layer4.1.bn2.weight", "layer4.1.bn2.bias")  # before any amplification


def _cmp_grads(eng_grads, g_fp32, g_emu):
    """engine-vs-fp32 error bounded by the bf16-emulation-vs-fp32 error, tensor by tensor (module docstring)."""
    worst = (0.0, None, 0.0)
    for k, r in g_fp32.items():
        e = eng_grads[k].cpu()
        rel_e, cos_e = _rel_cos(e, r)
        rel_m, cos_m = _rel_cos(g_emu[k], r)
        if rel_e > worst[0]:
            worst = (rel_e, k, rel_m)
        assert rel_e <= 1.3 * rel_m + 0.02, (k, rel_e, rel_m)
        assert cos_e >= cos_m - 0.03, (k, cos_e, cos_m)
        if k in TIGHT:
            rel_t, _ = _rel_cos(e, g_emu[k])
            assert rel_t <= 3e-2, (k, rel_t)
    return worst


def test_state_dict_roundtrip(engine):
    params, buffers = _load(engine)
    sd = engine.state_dict()
    ref = OC.state_dict_of(params, buffers)
    assert set(sd.keys()) == set(ref.keys())
    assert [k for k in sd if k in params] == list(params.keys())  # named_parameters order
    for k in ref:
        assert torch.equal(sd[k].cpu().to(ref[k].dtype), ref[k]), k


@pytest.mark.parametrize("train,sign,n", [(False, -1.0, 32), (True, 1.0, 32), (True, 1.0, 20), (False, -1.0, 3)])
def test_forward_backward_vs_oracle(engine, train, sign, n):
    params, buffers = _load(engine)
    x, y = _data(n)
    b = {k: v.clone() for k, v in buffers.items()}
    loss_ref, logits_ref, g_ref = OC.loss_and_grads(params, b, x, y, train=train, sign=sign)
    b2 = {k: v.clone() for k, v in buffers.items()}
    loss_emu, logits_emu, g_emu = OC.loss_and_grads(params, b2, x, y, train=train, sign=sign, emulate_bf16=True)
    engine.train(train)
    loss, logits = engine.forward_backward(x.cuda(), y.cuda(), loss_sign=sign, want_logits=True)
    torch.cuda.synchronize()
    err = (logits.cpu() - logits_ref).abs().max().item()
    assert err <= 0.06 + 0.02 * logits_ref.abs().max().item(), err
    assert (logits.cpu() - logits_emu).abs().max().item() <= 0.02 + 0.01 * logits_emu.abs().max().item()
    assert abs(loss.item() - loss_ref.item()) <= 2e-2 * max(1.0, abs(loss_ref.item()))
    assert abs(loss.item() - loss_emu.item()) <= 3e-3 * max(1.0, abs(loss_emu.item()))
    worst = _cmp_grads(engine.grad_dict(), g_ref, g_emu)
    print("worst (rel-L2 engine vs fp32, tensor, rel-L2 bf16-emulation vs fp32):", worst, "| logit err vs fp32", err)
    if train:  # BatchNorm buffers advance
        sd = engine.state_dict()
        for k in ("bn1.running_mean", "layer4.1.bn2.running_var", "layer2.0.downsample.1.running_mean"):
            torch.testing.assert_close(sd[k].cpu(), b[k], rtol=2e-2, atol=2e-3)


def test_eval_inference_matches_forward_backward_logits(engine):
    _load(engine)
    x, y = _data(16)
    engine.eval()
    _, l1 = engine.forward_backward(x.cuda(), y.cuda(), want_logits=True)
    l2 = engine.forward(x.cuda())
    assert torch.equal(l1, l2)


def test_masked_rl_steps_vs_oracle(engine):
    """three RL-style steps (RL.py:123-140): masked-out coordinates bit-identical; the update of every tensor is as close
    to the fp32 oracle's as the bf16-emulating oracle's update is (same criterion as the gradients)."""
    from unlearn_saliency_b200.engine import MaskedSGD
    params, buffers = _load(engine)
    params = {k: v.clone() for k, v in params.items()}
    b = {k: v.clone() for k, v in buffers.items()}
    g = torch.Generator().manual_seed(5)
    flat_mask = (torch.rand(engine.n_params, generator=g) < 0.5).to(torch.int64)
    mask = OC.split_mask(flat_mask, OC.resnet18_param_shapes(10))
    bits = engine.mask_bits_from_dict({k: v.cuda() for k, v in mask.items()})
    opt = MaskedSGD(engine, lr=0.013, momentum=0.9, weight_decay=5e-4, mask_bits=bits)
    ref_opt = OC.MaskedSGD(params, mask, lr=0.013, momentum=0.9, wd=5e-4)
    p0 = {k: v.clone() for k, v in params.items()}
    params_e = {k: v.clone() for k, v in params.items()}
    b_e = {k: v.clone() for k, v in buffers.items()}
    emu_opt = OC.MaskedSGD(params_e, mask, lr=0.013, momentum=0.9, wd=5e-4)
    engine.train(True)
    for s in range(3):
        x, y = _data(32, seed=100 + s)
        engine.forward_backward(x.cuda(), y.cuda())
        opt.step()
        OC.unlearn_step(params, b, ref_opt, x, y)
        OC.unlearn_step(params_e, b_e, emu_opt, x, y, emulate_bf16=True)
    torch.cuda.synchronize()
    for k, ref in params.items():
        e = engine.get_param(k).cpu()
        m = mask[k]
        assert torch.equal(e[m == 0], p0[k][m == 0]), k          # restore is exact (RL.py:17-34)
        upd_ref, upd, upd_emu = (ref - p0[k])[m == 1], (e - p0[k])[m == 1], (params_e[k] - p0[k])[m == 1]
        rel, _ = _rel_cos(upd, upd_ref)
        rel_m, _ = _rel_cos(upd_emu, upd_ref)
        assert rel <= 1.3 * rel_m + 0.03, (k, rel, rel_m)


def test_saliency_mask_end_to_end(engine, salun_ctx):
    """generate_mask.py:30-82 on the engine: Jaccard of the 50% mask against the fp32 oracle (bf16 operands)."""
    params, buffers = _load(engine)
    x, y = _data(64)
    b = {k: v.clone() for k, v in buffers.items()}
    absg = OC.accumulate_saliency(params, b, [(x[:32], y[:32]), (x[32:], y[32:])])
    ref = OC.masks_from_saliency(absg, [0.5])[0.5]
    absg_emu = OC.accumulate_saliency(params, b, [(x[:32], y[:32]), (x[32:], y[32:])], emulate_bf16=True)
    ref_emu = OC.masks_from_saliency(absg_emu, [0.5])[0.5]
    engine.eval()
    acc = torch.zeros_like(engine.params)
    for i in (0, 32):
        engine.forward_backward(x[i:i + 32].cuda(), y[i:i + 32].cuda(), loss_sign=-1.0)
        salun_ctx.saliency_accumulate_flat(engine.grads, acc)
    flat = engine.from_native_flat(acc).contiguous()
    k = int(flat.numel() * 0.5)
    m64, _, info = salun_ctx.topk_mask(flat, k, want_info=True)
    assert int(m64.sum()) == k
    m = m64.cpu()
    jac = float((m & ref).sum()) / float((m | ref).sum())
    jac_emu = float((m & ref_emu).sum()) / float((m | ref_emu).sum())
    jac_oo = float((ref_emu & ref).sum()) / float((ref_emu | ref).sum())
    print("end-to-end Jaccard: engine vs fp32 oracle", jac, "| engine vs bf16-emulating oracle", jac_emu,
          "| bf16-emulating vs fp32 oracle", jac_oo)
    # the index set is bit-exact on identical |G| (tests/test_tail_gpu.py); end to end, bf16 storage flips the elements
    # whose |G| sits within the rounding noise of the threshold
    assert jac_emu >= 0.95
    assert jac >= jac_oo - 0.02


def test_resnet34_forward_backward(salun_ctx):
    """resnet34 (ResNet.py:347, BasicBlock [3,4,6,3]) through the same engine: same tolerance model, batch 8."""
    from unlearn_saliency_b200.engine import ResNetEngine
    eng = ResNetEngine("resnet34", 10, 32, max_batch=8, ctx=salun_ctx)
    params, buffers = OC.synth_state(10, seed=1, depth=34)
    assert list(params.keys()) == list(eng.table.keys())
    eng.load_state_dict(OC.state_dict_of(params, buffers))
    x, y = _data(8, seed=4)
    for train, sign in ((False, -1.0), (True, 1.0)):
        b = {k: v.clone() for k, v in buffers.items()}
        _, logits_ref, g_ref = OC.loss_and_grads(params, b, x, y, train=train, sign=sign)
        b2 = {k: v.clone() for k, v in buffers.items()}
        _, _, g_emu = OC.loss_and_grads(params, b2, x, y, train=train, sign=sign, emulate_bf16=True)
        eng.train(train)
        _, logits = eng.forward_backward(x.cuda(), y.cuda(), loss_sign=sign, want_logits=True)
        assert (logits.cpu() - logits_ref).abs().max().item() <= 0.08 + 0.03 * logits_ref.abs().max().item()
        gd = eng.grad_dict()
        for k in ("fc.weight", "layer4.2.conv2.weight", "layer3.5.bn2.weight", "layer1.0.conv1.weight"):
            rel_e, cos_e = _rel_cos(gd[k].cpu(), g_ref[k])
            rel_m, cos_m = _rel_cos(g_emu[k], g_ref[k])
            assert rel_e <= 1.4 * rel_m + 0.03 and cos_e >= cos_m - 0.05, (k, rel_e, rel_m)
    eng.close()


@pytest.mark.parametrize("imagenet,size,n,classes", [(True, 64, 6, 10), (False, 32, 5, 10), (True, 64, 4, 200)])
def test_resnet50_forward_backward(salun_ctx, imagenet, size, n, classes):
    """resnet50 (Bottleneck, ResNet.py:358) with the ImageNet stem (7x7/2 + max pool, BASELINE config 4 architecture) and
    with the CIFAR stem: flat-activation runtime (csrc/salun_resnetb.cu), same tolerance model as resnet18.  200 classes x
    2048 features takes the wide-head path (logits / dW / dpooled as fp32 GEMMs + k_ce_rows)."""
    from unlearn_saliency_b200.engine import ResNetEngine
    eng = ResNetEngine("resnet50", classes, size, max_batch=8, ctx=salun_ctx, imagenet=imagenet)
    params, buffers = OC.synth_state_bottleneck(classes, seed=0, depth=50, imagenet=imagenet)
    assert list(params.keys()) == list(eng.table.keys()) and eng.n_params == sum(v.numel() for v in params.values())
    eng.load_state_dict(OC.state_dict_of(params, buffers))
    g = torch.Generator().manual_seed(21)
    x = torch.rand(n, 3, size, size, generator=g)
    y = torch.randint(0, classes, (n,), generator=g)
    for train, sign in ((False, -1.0), (True, 1.0)):
        b = {k: v.clone() for k, v in buffers.items()}
        loss_ref, logits_ref, g_ref = OC.bottleneck_loss_and_grads(params, b, x, y, train=train, sign=sign, imagenet=imagenet)
        b2 = {k: v.clone() for k, v in buffers.items()}
        _, logits_emu, g_emu = OC.bottleneck_loss_and_grads(params, b2, x, y, train=train, sign=sign, imagenet=imagenet,
                                                             emulate_bf16=True)
        eng.train(train)
        loss, logits = eng.forward_backward(x.cuda(), y.cuda(), loss_sign=sign, want_logits=True)
        torch.cuda.synchronize()
        scale = logits_ref.abs().max().item()
        err, err_emu = (logits.cpu() - logits_ref).abs().max().item(), (logits_emu - logits_ref).abs().max().item()
        assert err <= 1.5 * err_emu + 0.02 * scale + 0.05, (train, err, err_emu, scale)
        gd = eng.grad_dict()
        worst = (0.0, None, 0.0)
        for k, r in g_ref.items():
            rel_e, cos_e = _rel_cos(gd[k].cpu(), r)
            rel_m, cos_m = _rel_cos(g_emu[k], r)
            if rel_e > worst[0]:
                worst = (rel_e, k, rel_m)
            # 53 conv layers deep, <= 6 samples, 2x2 final feature maps: in train mode bf16 noise swamps the early-layer
            # gradients of the bf16 EMULATION itself (cos 0.2 against fp32), so direction is only compared where the
            # emulation is meaningful; magnitude of the error is always bounded by the emulation's
            assert rel_e <= 1.5 * rel_m + 0.05, (train, k, rel_e, rel_m)
            if cos_m >= 0.9:
                assert cos_e >= cos_m - 0.06, (train, k, cos_e, cos_m)
        print("resnet50", "imagenet" if imagenet else "cifar", "train" if train else "eval", "worst", worst, "logit err", err, err_emu)
    sd = eng.state_dict()
    assert set(sd.keys()) == set(OC.state_dict_of(params, buffers).keys())
    eng.close()


def test_graphed_step_equals_eager_step(salun_ctx):
    """engine.GraphedStep (CUDA graph of forward_backward + MaskedSGD.step) is bit-identical to the eager calls, and the
    capture leaves the model state untouched."""
    from unlearn_saliency_b200.engine import GraphedStep, MaskedSGD, ResNetEngine
    params, buffers = OC.synth_state(10, seed=0)
    sd = OC.state_dict_of(params, buffers)
    g = torch.Generator().manual_seed(5)
    n = 32
    xs = [torch.rand(n, 3, 32, 32, generator=g).cuda() for _ in range(3)]
    ys = [torch.randint(0, 10, (n,), generator=g).cuda() for _ in range(3)]
    res = []
    for graphed in (False, True):
        eng = ResNetEngine("resnet18", 10, 32, max_batch=n, ctx=salun_ctx)
        eng.load_state_dict(sd)
        bits = eng.ctx.pack_mask((torch.rand(eng.n_params, generator=torch.Generator().manual_seed(6)) < 0.5).to(torch.int64).cuda())
        opt = MaskedSGD(eng, 0.013, 0.9, 5e-4, mask_bits=bits)
        eng.train()
        p_before = eng.params.clone()
        step = GraphedStep(eng, opt, n) if graphed else None
        assert torch.equal(eng.params, p_before)          # capture + warm-up restored the snapshot
        losses = []
        for x, y in zip(xs, ys):
            if graphed:
                losses.append(float(step(x, y)))
            else:
                loss, _ = eng.forward_backward(x, y)
                opt.step()
                losses.append(float(loss))
        torch.cuda.synchronize()
        res.append((eng.params.clone(), eng.running_mean.clone(), losses))
        eng.close()
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]) and res[0][2] == res[1][2]
