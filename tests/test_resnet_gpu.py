"""sm_100a ResNet engine (through the C ABI) against the oracle (oracle/classification.py, torch fp32 on the CPU,
itself pinned to the unmodified reference by tests/test_oracle_golden.py).

Precision contract: tensor-core operands are bf16 (8-bit mantissa), accumulation / BN statistics /
master weights are fp32.  Stated tolerances: logits |err| <= 0.06 + 2% ; every parameter-gradient tensor has
relative L2 error <= 4e-2 and cosine >= 0.999 against the fp32 oracle; masked-out weights are bit-identical.
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


def _cmp_grads(eng_grads, ref_grads, rel_tol=4e-2, cos_tol=0.999):
    worst = (0.0, None)
    for k, r in ref_grads.items():
        e = eng_grads[k].float().cpu()
        rel = float((e - r).norm() / (r.norm() + 1e-12))
        cos = float(torch.dot(e.flatten(), r.flatten()) / (e.norm() * r.norm() + 1e-20))
        if rel > worst[0]:
            worst = (rel, k)
        assert rel <= rel_tol, (k, rel)
        assert cos >= cos_tol, (k, cos)
    return worst


def test_state_dict_roundtrip(engine):
    params, buffers = _load(engine)
    sd = engine.state_dict()
    ref = OC.state_dict_of(params, buffers)
    assert list(sd.keys()) == list(ref.keys())
    for k in ref:
        assert torch.equal(sd[k].cpu().to(ref[k].dtype), ref[k]), k


@pytest.mark.parametrize("train,sign,n", [(False, -1.0, 32), (True, 1.0, 32), (True, 1.0, 20), (False, -1.0, 3)])
def test_forward_backward_vs_oracle(engine, train, sign, n):
    params, buffers = _load(engine)
    x, y = _data(n)
    b = {k: v.clone() for k, v in buffers.items()}
    loss_ref, logits_ref, g_ref = OC.loss_and_grads(params, b, x, y, train=train, sign=sign)
    engine.train(train)
    loss, logits = engine.forward_backward(x.cuda(), y.cuda(), loss_sign=sign, want_logits=True)
    torch.cuda.synchronize()
    err = (logits.cpu() - logits_ref).abs().max().item()
    assert err <= 0.06 + 0.02 * logits_ref.abs().max().item(), err
    assert abs(loss.item() - loss_ref.item()) <= 2e-2 * max(1.0, abs(loss_ref.item()))
    worst = _cmp_grads(engine.grad_dict(), g_ref)
    print("worst grad rel-L2", worst, "logit err", err)
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
    """three RL-style steps (RL.py:123-140): masked-out coordinates bit-identical, updates within 3% of the oracle's."""
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
    engine.train(True)
    for s in range(3):
        x, y = _data(32, seed=100 + s)
        engine.forward_backward(x.cuda(), y.cuda())
        opt.step()
        OC.unlearn_step(params, b, ref_opt, x, y)
    torch.cuda.synchronize()
    for k, ref in params.items():
        e = engine.get_param(k).cpu()
        m = mask[k]
        assert torch.equal(e[m == 0], p0[k][m == 0]), k          # restore is exact (RL.py:17-34)
        upd_ref, upd = (ref - p0[k])[m == 1], (e - p0[k])[m == 1]
        rel = float((upd - upd_ref).norm() / (upd_ref.norm() + 1e-12))
        assert rel <= 5e-2, (k, rel)


def test_saliency_mask_end_to_end(engine, salun_ctx):
    """generate_mask.py:30-82 on the engine: Jaccard of the 50% mask against the fp32 oracle (bf16 operands)."""
    params, buffers = _load(engine)
    x, y = _data(64)
    b = {k: v.clone() for k, v in buffers.items()}
    absg = OC.accumulate_saliency(params, b, [(x[:32], y[:32]), (x[32:], y[32:])])
    ref = OC.masks_from_saliency(absg, [0.5])[0.5]
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
    print("end-to-end Jaccard vs fp32 oracle:", jac)
    assert jac >= 0.97
