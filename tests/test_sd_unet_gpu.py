"""SD U-Net forward on the engine (sd/engine.py over the op-level C ABI, csrc/salun_ops.cu):
  * whole-network eps against the UNMODIFIED reference UNetModel (tests/golden/sd_unet.npz), both builds, graph and eager;
  * the SD-specific ops one by one against their torch statements (LayerNorm, GEGLU, multi-head self / cross attention with
    the 77-token context and 40-wide heads, cos|sin timestep embedding, wide GroupNorm, stride-2 convolution)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.golden.make_golden_sd import CONFIGS, sd_inputs, sd_synth_weights

pytestmark = pytest.mark.gpu
Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "sd_unet.npz"))
TOL = {"bf16": 4e-2, "split": 2e-3}


def _rel(a, r):
    a, r = torch.as_tensor(a).float().cpu().flatten(), torch.as_tensor(r).float().cpu().flatten()
    return float((a - r).norm() / r.norm())


class Ops:
    """act-typed scratch helpers for the op tests"""

    def __init__(self, ctx, precision):
        from unlearn_saliency_b200 import _lib
        from unlearn_saliency_b200.tail import _ptr, _stream
        self.L, self.ctx, self.p, self.s = _lib.lib(precision), ctx, _ptr, lambda: _stream(ctx.device)
        self.ab = self.L.salun_act_bytes()

    def act(self, x2d):   # fp32 [rows][cols] -> act buffer
        x2d = x2d.float().cuda().contiguous()
        buf = torch.zeros(x2d.numel() * self.ab, dtype=torch.uint8, device="cuda")
        assert self.L.salun_op_f32_to_act(self.ctx.handle, self.p(x2d), x2d.shape[1], self.p(buf), x2d.shape[1], x2d.shape[0],
                                          x2d.shape[1], self.s()) == 0
        return buf

    def f32(self, buf, rows, cols):
        out = torch.empty(rows, cols, device="cuda")
        assert self.L.salun_op_act_to_f32(self.ctx.handle, self.p(buf), cols, self.p(out), cols, rows, cols, self.s()) == 0
        return out

    def empty(self, elems):
        return torch.zeros(elems * self.ab, dtype=torch.uint8, device="cuda")


@pytest.mark.parametrize("precision", ["split", "bf16"])
@pytest.mark.parametrize("tag", ["a", "b"])
def test_unet_forward_matches_reference_golden(salun_ctx, tag, precision):
    from unlearn_saliency_b200 import _lib
    from unlearn_saliency_b200.sd.engine import SDUNetEngine
    if precision not in _lib.available_precisions():
        pytest.skip("build missing")
    c = CONFIGS[tag]
    eng = SDUNetEngine(c["cfg"], latent_size=c["latent"], max_batch=c["n"], context_len=c["ctx_len"], ctx=salun_ctx,
                       precision=precision, use_graph=True)
    eng.load_state_dict(sd_synth_weights(eng.table, seed=7))
    x, t, ctx = sd_inputs(tag)
    eps = eng.forward(x.cuda(), t.cuda(), ctx.cuda())
    torch.cuda.synchronize()
    r = _rel(eps, Z[f"{tag}_eps"])
    print(f"SD U-Net config {tag} [{precision}]: eps relative error vs the reference {r:.3e} ({eng.launches_per_forward(c['n'])} op calls)")
    assert r < TOL[precision], r
    eps2 = eng.forward(x.cuda(), t.cuda(), ctx.cuda())          # graph replay is deterministic
    assert torch.equal(eps, eps2)
    if tag == "a":                                              # eager program == captured graph, and a smaller batch
        e2 = SDUNetEngine(c["cfg"], latent_size=c["latent"], max_batch=c["n"], context_len=c["ctx_len"], ctx=salun_ctx,
                          precision=precision, use_graph=False)
        e2.load_state_dict(sd_synth_weights(eng.table, seed=7))
        assert torch.equal(e2.forward(x.cuda(), t.cuda(), ctx.cuda()), eps)
        one = e2.forward(x[:1].cuda(), t[:1].cuda(), ctx[:1].cuda())
        assert _rel(one, Z[f"{tag}_eps"][:1]) < TOL[precision]


@pytest.mark.parametrize("precision", ["split", "bf16"])
def test_sd_ops_vs_torch(salun_ctx, precision):
    from unlearn_saliency_b200 import _lib
    if precision not in _lib.available_precisions():
        pytest.skip("build missing")
    o = Ops(salun_ctx, precision)
    L, h, P = o.L, salun_ctx.handle, o.p
    tol = 3e-2 if precision == "bf16" else 3e-4
    g = torch.Generator().manual_seed(0)
    # ---- LayerNorm (C = 320) and GEGLU
    rows, C = 200, 320
    x = torch.randn(rows, C, generator=g) * 2 + 0.5
    gam, bet = torch.randn(C, generator=g).cuda(), torch.randn(C, generator=g).cuda()
    out = o.empty(rows * C)
    assert L.salun_sd_layernorm(h, P(o.act(x)), P(gam), P(bet), P(out), rows, C, 1e-5, o.s()) == 0
    assert _rel(o.f32(out, rows, C), F.layer_norm(x.cuda(), (C,), gam, bet, 1e-5)) < tol
    pr = torch.randn(rows, 2 * 1280, generator=g)
    out = o.empty(rows * 1280)
    assert L.salun_sd_geglu(h, P(o.act(pr)), P(out), rows, 1280, o.s()) == 0
    a, gate = pr.cuda().chunk(2, dim=-1)
    assert _rel(o.f32(out, rows, 1280), a * F.gelu(gate)) < tol
    # ---- timestep embedding (util.py:173-197)
    t = torch.tensor([0.0, 1.0, 37.0, 999.0]).cuda()
    emb = torch.empty(4, 320, device="cuda")
    assert L.salun_sd_timestep_embedding(h, P(t), P(emb), 4, 320, 10000.0, o.s()) == 0
    half = 160
    freqs = torch.exp(-np.log(10000.0) * torch.arange(half, dtype=torch.float32) / half).cuda()
    args = t[:, None] * freqs[None]
    np.testing.assert_allclose(emb.cpu().numpy(), torch.cat([args.cos(), args.sin()], -1).cpu().numpy(), rtol=2e-4, atol=2e-4)
    # ---- multi-head attention: self (T = 64, 8 heads of 40) and cross (77 keys)
    for (n, Tq, Tk, heads, d) in ((2, 64, 64, 8, 40), (2, 256, 77, 8, 40), (1, 16, 77, 2, 64), (3, 128, 128, 4, 160)):
        Cc = heads * d
        q, k, v = (torch.randn(n * T_, Cc, generator=g) for T_ in (Tq, Tk, Tk))
        nb = int(L.salun_sd_attention_ws_bytes(n, Tq, Tk, heads, d))
        ws = torch.zeros(nb, dtype=torch.uint8, device="cuda")
        out = o.empty(n * Tq * Cc)
        assert L.salun_sd_attention(h, P(ws), nb, P(o.act(q)), P(o.act(k)), P(o.act(v)), P(out), n, Tq, Tk, heads, d, o.s()) == 0
        sp = lambda z, T_: z.cuda().view(n, T_, heads, d).permute(0, 2, 1, 3)
        ref = torch.softmax(sp(q, Tq) @ sp(k, Tk).transpose(-1, -2) * d ** -0.5, dim=-1) @ sp(v, Tk)
        ref = ref.permute(0, 2, 1, 3).reshape(n * Tq, Cc)
        assert _rel(o.f32(out, n * Tq, Cc), ref) < tol, (n, Tq, Tk, heads, d)
    # ---- GroupNorm at 1920 channels (+ SiLU), conv 3x3 with all epilogue terms, stride-2 conv
    n, H, C = 2, 8, 1920
    x = torch.randn(n, C, H, H, generator=g)
    xp = o.empty(n * (H + 2) * (H + 2) * C)
    assert L.salun_op_nchw_to_padded(h, P(x.cuda()), P(xp), n, C, C, H, H, o.s()) == 0
    gam, bet = torch.randn(C, generator=g).cuda(), torch.randn(C, generator=g).cuda()
    st = torch.zeros(int(L.salun_op_groupnorm_ws_floats(n)), device="cuda")
    yp = o.empty(n * (H + 2) * (H + 2) * C)
    assert L.salun_op_groupnorm(h, P(xp), P(gam), P(bet), P(st), P(yp), 0, n, H, H, C, 1e-5, 1, o.s()) == 0
    back = torch.empty(n, C, H, H, device="cuda")
    assert L.salun_op_padded_to_nchw(h, P(yp), P(back), n, C, H, H, o.s()) == 0
    assert _rel(back, F.silu(F.group_norm(x.cuda(), 32, gam, bet, 1e-5))) < tol
    cin, cout = 128, 192
    x = torch.randn(n, cin, H, H, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5
    b, rb = torch.randn(cout, generator=g).cuda(), torch.randn(n, cout, generator=g).cuda()
    res = torch.randn(n, cout, H, H, generator=g)
    wk = salun_wop(L, h, P, w, cout, cin, 3, o)
    xp, rp = o.empty(n * (H + 2) ** 2 * cin), o.empty(n * (H + 2) ** 2 * cout)
    L.salun_op_nchw_to_padded(h, P(x.cuda()), P(xp), n, cin, cin, H, H, o.s())
    L.salun_op_nchw_to_padded(h, P(res.cuda()), P(rp), n, cout, cout, H, H, o.s())
    yp = o.empty(n * (H + 2) ** 2 * cout)
    assert L.salun_op_conv(h, P(xp), 0, P(wk), P(b), P(rb), cout, P(rp), P(yp), 1, None, n, H, H, cin, cout, 3, o.s()) == 0
    back = torch.empty(n, cout, H, H, device="cuda")
    L.salun_op_padded_to_nchw(h, P(yp), P(back), n, cout, H, H, o.s())
    ref = F.conv2d(x.cuda(), w.cuda(), b, padding=1) + rb[:, :, None, None] + res.cuda()
    assert _rel(back, ref) < tol
    col = o.empty(n * (H // 2) ** 2 * 9 * cin)
    yp = o.empty(n * (H // 2 + 2) ** 2 * cout)
    assert L.salun_op_conv_s2(h, P(xp), P(col), P(wk), P(b), P(yp), n, H, H, cin, cout, o.s()) == 0
    back = torch.empty(n, cout, H // 2, H // 2, device="cuda")
    L.salun_op_padded_to_nchw(h, P(yp), P(back), n, cout, H // 2, H // 2, o.s())
    assert _rel(back, F.conv2d(x.cuda(), w.cuda(), b, stride=2, padding=1)) < tol


def salun_wop(L, h, P, w, cout, cin, ks, o):
    wk = torch.empty(cout * ks * ks * cin * L.salun_wop_k(), dtype=torch.bfloat16, device="cuda")
    assert L.salun_op_prep_weight(h, P(w.cuda().contiguous()), P(wk), cout, cin, ks, cout, cin, o.s()) == 0
    return wk


@pytest.mark.parametrize("precision", ["split", "bf16"])
def test_splitk_convolution_equals_unsplit(salun_ctx, precision):
    """Small-M / deep-K convolutions (the 8x8 level of the U-Net at batch 2: 10 output tiles for 148 SMs) split the k loop
    over several CTAs per tile when the context has a scratch; same result as the unsplit launch and as torch."""
    import ctypes as C
    from unlearn_saliency_b200 import _lib
    if precision not in _lib.available_precisions():
        pytest.skip("build missing")
    o = Ops(salun_ctx, precision)
    L, h, P = o.L, salun_ctx.handle, o.p
    g = torch.Generator().manual_seed(3)
    for (n, H, cin, cout, ks, flat) in [(2, 8, 1280, 1280, 3, False), (2, 8, 2560, 1280, 1, True), (1, 16, 640, 320, 3, False)]:
        x = torch.randn(n, cin, H, H, generator=g)
        w = torch.randn(cout, cin, ks, ks, generator=g) / (ks * ks * cin) ** 0.5
        b, rb = torch.randn(cout, generator=g).cuda(), torch.randn(n, cout, generator=g).cuda()
        res = torch.randn(n, cout, H, H, generator=g)
        wk = salun_wop(L, h, P, w, cout, cin, ks, o)
        rp = o.empty(n * (H + 2) ** 2 * cout)
        L.salun_op_nchw_to_padded(h, P(res.cuda()), P(rp), n, cout, cout, H, H, o.s())
        if flat:
            xin = o.act(x.permute(0, 2, 3, 1).reshape(n * H * H, cin))
        else:
            xin = o.empty(n * (H + 2) ** 2 * cin)
            L.salun_op_nchw_to_padded(h, P(x.cuda()), P(xin), n, cin, cin, H, H, o.s())
        outs, launches = [], []
        for scratch in (False, True):
            if scratch:
                salun_ctx._op_scratch = None
                salun_ctx.ensure_op_scratch()
            else:
                assert L.salun_op_set_scratch(h, None, 0) == 0
            yp = o.empty(n * (H + 2) ** 2 * cout)
            yf = torch.zeros(n * H * H, cout, device="cuda")
            l0 = L.salun_launch_count()
            assert L.salun_op_conv(h, P(xin), int(flat), P(wk), P(b), P(rb), cout, P(rp), P(yp), 1, None, n, H, H, cin, cout, ks,
                                   o.s()) == 0, L.salun_last_error()
            assert L.salun_op_conv(h, P(xin), int(flat), P(wk), P(b), P(rb), cout, None, None, 0, P(yf), n, H, H, cin, cout, ks,
                                   o.s()) == 0, L.salun_last_error()
            launches.append(L.salun_launch_count() - l0)
            back = torch.empty(n, cout, H, H, device="cuda")
            L.salun_op_padded_to_nchw(h, P(yp), P(back), n, cout, H, H, o.s())
            outs.append((back, yf.clone()))
        ref32 = F.conv2d(x.cuda(), w.cuda(), b, padding=ks // 2) + rb[:, :, None, None]
        ref = ref32 + res.cuda()
        assert launches == [2, 4], launches                       # GEMM | GEMM + split-K epilogue, twice
        for back, yf in outs:
            assert _rel(back, ref) < TOL[precision] / 4
            assert _rel(yf.view(n, H, H, cout).permute(0, 3, 1, 2), ref32) < (5e-3 if precision == "bf16" else 5e-4)  # split: fp32 accumulator truncation over 2880 MMAs
        assert _rel(outs[1][1], outs[0][1]) < 1e-4                 # fp32 sums in a different order only


@pytest.mark.parametrize("precision", ["split", "bf16"])
def test_flash_attention_many_key_blocks(salun_ctx, precision):
    """The fused attention kernel (head width <= 64) over several key blocks, with ragged token counts, padded keys in the
    last block, and scores large enough that the running maximum moves between blocks; also against the unfused chain."""
    import subprocess, sys
    from unlearn_saliency_b200 import _lib
    if precision not in _lib.available_precisions():
        pytest.skip("build missing")
    o = Ops(salun_ctx, precision)
    L, h, P = o.L, salun_ctx.handle, o.p
    g = torch.Generator().manual_seed(11)
    tol = {"bf16": 2e-2, "split": 2e-4}[precision]
    for (n, Tq, Tk, heads, d, qs) in ((1, 1024, 1024, 8, 40, 1.0), (2, 300, 200, 3, 24, 4.0), (1, 4096, 4096, 2, 40, 2.0),
                                      (2, 129, 77, 8, 64, 1.0), (1, 128, 65, 1, 8, 6.0)):
        Cc = heads * d
        q, k, v = (torch.randn(n * T_, Cc, generator=g) for T_ in (Tq, Tk, Tk))
        q = q * qs
        nb = int(L.salun_sd_attention_ws_bytes(n, Tq, Tk, heads, d))
        ws = torch.zeros(nb, dtype=torch.uint8, device="cuda")
        out = o.empty(n * Tq * Cc)
        assert L.salun_sd_attention(h, P(ws), nb, P(o.act(q)), P(o.act(k)), P(o.act(v)), P(out), n, Tq, Tk, heads, d, o.s()) == 0
        torch.cuda.synchronize()
        sp = lambda z, T_: z.cuda().double().view(n, T_, heads, d).permute(0, 2, 1, 3)
        ref = torch.softmax(sp(q, Tq) @ sp(k, Tk).transpose(-1, -2) * d ** -0.5, dim=-1) @ sp(v, Tk)
        ref = ref.permute(0, 2, 1, 3).reshape(n * Tq, Cc).float()
        got = o.f32(out, n * Tq, Cc)
        err = _rel(got, ref)
        print(f"flash attention [{precision}] n={n} Tq={Tq} Tk={Tk} heads={heads} d={d} qscale={qs}: rel err {err:.3e}")
        assert torch.isfinite(got).all() and err < tol, (n, Tq, Tk, heads, d, err)
