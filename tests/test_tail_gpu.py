"""Parity of the CUDA tail kernels (through the C ABI) against oracle/salun_oracle.c -- BIT-EXACT."""
import numpy as np
import pytest
import torch

from oracle import tail as O

pytestmark = pytest.mark.gpu


def _dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def _saliency_like(rng, n):
    g = rng.standard_normal(n).astype(np.float32) * np.exp(rng.standard_normal(n).astype(np.float32) * 3)
    g[rng.integers(0, n, max(1, n // 30))] = 0.0  # dead-ReLU zeros
    return g


@pytest.mark.parametrize("n", [1, 3, 31, 32, 33, 1000, 1024, 4097, 1 << 20, 11173962])
def test_accumulate_abs_bitexact(salun_ctx, n):
    rng = np.random.default_rng(n)
    acc = rng.standard_normal(n).astype(np.float32)
    g = rng.standard_normal(n).astype(np.float32)
    dacc, dg = _dev(acc), _dev(g)
    salun_ctx.saliency_accumulate_flat(dg, dacc)
    O.saliency_accumulate(acc, g)
    assert np.array_equal(dacc.cpu().numpy(), acc)
    salun_ctx.abs_(dacc)
    O.abs_inplace(acc)
    assert np.array_equal(dacc.cpu().numpy(), acc)


def test_accumulate_multi_tensor(salun_ctx):
    rng = np.random.default_rng(1)
    shapes = [(64, 3, 3, 3), (64,), (64,), (128, 64, 3, 3), (10, 512), (10,), (1,), (7, 5)]
    grads = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    n = sum(g.size for g in grads)
    acc = rng.standard_normal(n).astype(np.float32)
    dacc = _dev(acc)
    scale = torch.tensor([0.37], device="cuda")
    for _ in range(2):
        salun_ctx.saliency_accumulate([_dev(g) for g in grads], dacc)
        O.saliency_accumulate(acc, np.concatenate([g.ravel() for g in grads]))
    assert np.array_equal(dacc.cpu().numpy(), acc)
    salun_ctx.saliency_accumulate([_dev(g) for g in grads], dacc, scale=scale)
    flat = np.concatenate([g.ravel() for g in grads]) * np.float32(0.37)
    O.saliency_accumulate(acc, flat.astype(np.float32))
    assert np.array_equal(dacc.cpu().numpy(), acc)


@pytest.mark.parametrize("n,ratio", [(1, 0.5), (5, 0.5), (100, 0.1), (1000, 0.5), (4097, 0.3), (65536, 0.9),
                                      (1 << 20, 0.5), (11173962, 0.5), (11173962, 0.1), (11173962, 1.0),
                                      (11173962, 0.97)])
def test_topk_mask_bitexact(salun_ctx, n, ratio):
    rng = np.random.default_rng(n + int(ratio * 100))
    g = _saliency_like(rng, n)
    if n > 50:
        g[5] = np.nan
        g[7] = np.inf
        g[9] = -np.inf
    k = int(n * ratio)
    m64, bits, info = salun_ctx.topk_mask(_dev(g), k, want_info=True)
    om, ob, thr, ngt, neq = O.topk_mask(np.abs(g), k)
    assert np.array_equal(m64.cpu().numpy(), om)
    assert np.array_equal(bits.cpu().numpy().view(np.uint32), ob)
    assert int(m64.sum()) == k
    if 0 < k < n:
        assert (info.thr_key, info.n_greater, info.n_equal) == (thr, ngt, neq)


def test_topk_ties_flat_order(salun_ctx):
    # many exact ties at the threshold: the first (k - n_gt) in flat order must be taken
    rng = np.random.default_rng(0)
    n = 300000
    g = rng.integers(0, 4, n).astype(np.float32)  # values 0..3, huge tie classes
    for k in [1, 1000, n // 2, n - 5]:
        m64, bits, info = salun_ctx.topk_mask(_dev(g), k, want_info=True)
        om, ob, *_ = O.topk_mask(g, k)
        assert np.array_equal(m64.cpu().numpy(), om), k
        assert np.array_equal(bits.cpu().numpy().view(np.uint32), ob), k
    small = rng.integers(0, 3, 2000).astype(np.float32)
    m64, _, _ = salun_ctx.topk_mask(_dev(small), 700)
    assert np.array_equal(m64.cpu().numpy(), O.topk_mask_argsort(small, 700))


@pytest.mark.parametrize("n", [1, 37, 4097, 300000, 11173962])
def test_topk_mask_multi_equals_single_and_oracle(salun_ctx, n):
    """salun_topk_mask_multi (one sweep for the reference's whole threshold_list) against the C oracle and against the
    single-ratio entry point: bit-identical masks, bits and info, including k = 0, k = n, huge tie classes and NaN / inf."""
    rng = np.random.default_rng(n)
    for kind in ("saliency", "ties"):
        g = _saliency_like(rng, n) if kind == "saliency" else rng.integers(0, 4, n).astype(np.float32)
        if n > 50 and kind == "saliency":
            g[5], g[7], g[9] = np.nan, np.inf, -np.inf
        ks = [int(n * r) for r in (0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0)] + [0, max(0, n - 5), 1]
        d = _dev(g)
        m64s, bitss, infos = salun_ctx.topk_mask_multi(d, ks, want_info=True)
        for k, m64, bits, info in zip(ks, m64s, bitss, infos):
            om, ob, thr, ngt, neq = O.topk_mask(np.abs(g), k)
            assert np.array_equal(m64.cpu().numpy(), om), (n, kind, k)
            assert np.array_equal(bits.cpu().numpy().view(np.uint32), ob), (n, kind, k)
            if 0 < k < n:
                assert (info.thr_key, info.n_greater, info.n_equal) == (thr, ngt, neq), (n, kind, k)
                s64, sbits, sinfo = salun_ctx.topk_mask(d, k, want_info=True)
                assert torch.equal(s64, m64) and torch.equal(sbits, bits)
                assert (sinfo.thr_key, sinfo.thr_value, sinfo.n_greater, sinfo.n_equal) == \
                       (info.thr_key, info.thr_value, info.n_greater, info.n_equal)
    # bits only / int64 only
    m64s, bitss, _ = salun_ctx.topk_mask_multi(d, ks[:3], want_i64=False)
    assert m64s is None and len(bitss) == 3
    with pytest.raises(ValueError):
        salun_ctx.topk_mask_multi(d, list(range(17)))


def test_pack_unpack(salun_ctx):
    rng = np.random.default_rng(3)
    for n in [1, 31, 32, 33, 100003]:
        m = (rng.random(n) < 0.4).astype(np.int64)
        bits = salun_ctx.pack_mask(_dev(m))
        assert np.array_equal(bits.cpu().numpy().view(np.uint32), O.pack_mask(m))
        back = salun_ctx.unpack_mask(bits, n)
        assert np.array_equal(back.cpu().numpy(), m)


@pytest.mark.parametrize("n", [1, 5, 4097, 11173962])
@pytest.mark.parametrize("masked", [True, False])
def test_masked_sgd_bitexact(salun_ctx, n, masked):
    rng = np.random.default_rng(n)
    p = rng.standard_normal(n).astype(np.float32)
    v = np.zeros(n, np.float32)
    m = (rng.random(n) < 0.5).astype(np.int64)
    bits = O.pack_mask(m) if masked else None
    dp, dv = _dev(p), _dev(v)
    dbits = _dev(bits.view(np.int32)) if masked else None
    p0 = p.copy()
    for it in range(3):
        g = rng.standard_normal(n).astype(np.float32)
        salun_ctx.masked_sgd_step(dp, _dev(g), dv, dbits, 0.013, 0.9, 5e-4)
        O.masked_sgd_step(p, g, v, bits, 0.013, 0.9, 5e-4)
    assert np.array_equal(dp.cpu().numpy(), p)
    assert np.array_equal(dv.cpu().numpy(), v)
    if masked:
        assert np.array_equal(p[m == 0], p0[m == 0])  # theta0 restored exactly (RL.py:17-34)
        assert not np.any(v[m == 0])


def test_masked_sgd_matches_torch_optim(salun_ctx):
    """End-to-end against the reference statements themselves (RL.py:134-140 with torch.optim.SGD), 1e-6 rel."""
    rng = np.random.default_rng(11)
    n = 50021
    p = rng.standard_normal(n).astype(np.float32)
    m = (rng.random(n) < 0.5).astype(np.int64)
    dp, dv = _dev(p), torch.zeros(n, device="cuda")
    dbits = salun_ctx.pack_mask(_dev(m))
    tp = torch.nn.Parameter(torch.tensor(p.copy()))
    opt = torch.optim.SGD([tp], lr=0.013, momentum=0.9, weight_decay=5e-4)
    th0, tm = tp.detach().clone(), torch.tensor(m)
    for _ in range(4):
        g = rng.standard_normal(n).astype(np.float32)
        salun_ctx.masked_sgd_step(dp, _dev(g), dv, dbits, 0.013, 0.9, 5e-4)
        tp.grad = torch.tensor(g.copy())
        tp.grad *= tm
        opt.step()
        with torch.no_grad():
            mt = tm.float()
            tp.data.mul_(mt).add_(th0 * (1 - mt))
            opt.state[tp]["momentum_buffer"].mul_(mt)
    torch.testing.assert_close(dp.cpu(), tp.detach(), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("n", [7, 4097, 3000000])
def test_clip_masked_adam_bitexact(salun_ctx, n):
    rng = np.random.default_rng(n)
    p = rng.standard_normal(n).astype(np.float32)
    m1 = np.zeros(n, np.float32)
    m2 = np.zeros(n, np.float32)
    m = (rng.random(n) < 0.5).astype(np.int64)
    bits = O.pack_mask(m)
    dp, d1, d2 = _dev(p), _dev(m1), _dev(m2)
    dbits = _dev(bits.view(np.int32))
    p0 = p.copy()
    for step in range(1, 4):
        g = (rng.standard_normal(n) * 2).astype(np.float32)
        dg = _dev(g)
        ss = salun_ctx.grad_sumsq(dg)
        coef = salun_ctx.clip_coef(ss, 1.0)
        salun_ctx.masked_adam_step(dp, dg, d1, d2, dbits, 1e-4, 0.9, 0.999, 1e-8, 0.0, step, coef)
        tn = O.grad_norm(g)
        assert abs(float(ss.item()) ** 0.5 - tn) <= 1e-9 * tn
        c = O.clip_coef(tn, 1.0)
        assert np.float32(coef.item()) == np.float32(c)
        O.masked_adam_step(p, g, m1, m2, bits, 1e-4, 0.9, 0.999, 1e-8, 0.0, step, c)
    assert np.array_equal(dp.cpu().numpy(), p)
    assert np.array_equal(d1.cpu().numpy(), m1)
    assert np.array_equal(d2.cpu().numpy(), m2)
    assert np.array_equal(p[m == 0], p0[m == 0])


def test_apply_mask(salun_ctx):
    rng = np.random.default_rng(5)
    n = 100003
    g = rng.standard_normal(n).astype(np.float32)
    m = (rng.random(n) < 0.5).astype(np.int64)
    bits = O.pack_mask(m)
    dg = _dev(g)
    salun_ctx.apply_mask(dg, _dev(bits.view(np.int32)))
    O.apply_mask(g, bits)
    assert np.array_equal(dg.cpu().numpy(), g)


def test_empty_inputs_are_noops(salun_ctx):
    e = torch.zeros(0, device="cuda")
    salun_ctx.saliency_accumulate_flat(e, e)
    salun_ctx.abs_(e)
    salun_ctx.masked_sgd_step(e, e, e, None, 0.1, 0.9, 0.0)
    m64, bits, info = salun_ctx.topk_mask(e, 0, want_info=True)
    assert m64.numel() == 0 and bits.numel() == 0
    assert float(salun_ctx.grad_sumsq(e).item()) == 0.0


def test_topk_at_sd_unet_scale_properties(salun_ctx):
    """N = 859 520 964 (SD v1.4 U-Net): too large for the scalar oracle, so size-independent properties:
    exactly k ones, every selected |g| >= every unselected |g|, packed bits == int64 mask, nested in the ratio."""
    n = 859520964
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(n, device="cuda", generator=g)
    a[::97] = 0.0
    prev = None
    for ratio in (0.3, 0.5):
        k = int(n * ratio)
        m64, bits, info = salun_ctx.topk_mask(a, k, want_info=True)
        assert int(m64.sum()) == k
        sel = m64.bool()
        absa = a.abs()
        assert float(absa[sel].min()) >= float(absa[~sel].max())
        assert float(absa[sel].min()) == info.thr_value
        assert torch.equal(salun_ctx.pack_mask(m64), bits)
        if prev is not None:
            assert bool((m64 >= prev).all())
        prev = m64
        del sel, absa
    # fused clip + masked Adam at the same scale: masked-out coordinates bit-identical, others moved
    p = torch.randn(n, device="cuda", generator=g)
    p0 = p.clone()
    m1, m2 = torch.zeros_like(p), torch.zeros_like(p)
    ss = salun_ctx.grad_sumsq(a)
    assert abs(float(ss.item()) - float(a.double().square().sum().item())) <= 1e-9 * float(ss.item())
    coef = salun_ctx.clip_coef(ss, 1.0)
    salun_ctx.masked_adam_step(p, a, m1, m2, bits, 1e-5, 0.9, 0.999, 1e-8, 0.0, 1, coef)
    sel = prev.bool()
    assert torch.equal(p[~sel], p0[~sel])
    assert float((p[sel] != p0[sel]).float().mean()) > 0.99
