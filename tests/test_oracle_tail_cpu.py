"""CPU checks of oracle/salun_oracle.c against the reference formulas (stable argsort, torch.optim)."""
import numpy as np
import torch

from oracle import tail as O


def test_topk_matches_reference_formula_with_ties_nan():
    rng = np.random.default_rng(0)
    for n, k in [(1, 0), (1, 1), (37, 5), (1000, 100), (1000, 500), (4096, 4095), (5000, 5000), (5000, 2500)]:
        a = np.abs(rng.standard_normal(n)).astype(np.float32)
        if n > 10:
            a[rng.integers(0, n, n // 5)] = 0.0
            a[rng.integers(0, n, n // 10)] = a[3]
        if n > 100:
            a[7] = np.nan
        m, b, thr, ngt, neq = O.topk_mask(a, k)
        assert np.array_equal(m, O.topk_mask_argsort(a, k)), (n, k)
        assert m.sum() == k
        assert np.array_equal(O.pack_mask(m), b)
    # empty input
    m, b, *_ = O.topk_mask(np.zeros(0, np.float32), 0)
    assert m.size == 0 and b.size == 0


def test_masked_sgd_matches_torch_optim_and_restore():
    rng = np.random.default_rng(1)
    n = 10007
    p = rng.standard_normal(n).astype(np.float32)
    v = np.zeros(n, np.float32)
    m = (rng.random(n) < 0.5).astype(np.int64)
    bits = O.pack_mask(m)
    tp = torch.nn.Parameter(torch.tensor(p.copy()))
    opt = torch.optim.SGD([tp], lr=0.013, momentum=0.9, weight_decay=5e-4)
    th0, tm = tp.detach().clone(), torch.tensor(m)
    for _ in range(5):
        g = rng.standard_normal(n).astype(np.float32)
        O.masked_sgd_step(p, g, v, bits, 0.013, 0.9, 5e-4)
        tp.grad = torch.tensor(g.copy()); tp.grad *= tm; opt.step()  # RL.py:11-14, impl.py:68-73
        with torch.no_grad():  # RL.py:17-34
            mt = tm.float(); tp.data.mul_(mt).add_(th0 * (1 - mt)); opt.state[tp]["momentum_buffer"].mul_(mt)
    np.testing.assert_allclose(p, tp.detach().numpy(), rtol=1e-6, atol=1e-6)
    assert np.array_equal(p[m == 0], th0.numpy()[m == 0])
    np.testing.assert_allclose(v, opt.state[tp]["momentum_buffer"].numpy(), rtol=1e-6, atol=1e-6)


def test_clip_masked_adam_matches_torch():
    rng = np.random.default_rng(2)
    n = 9001
    p = rng.standard_normal(n).astype(np.float32); m1 = np.zeros(n, np.float32); m2 = np.zeros(n, np.float32)
    m = (rng.random(n) < 0.5).astype(np.int64); bits = O.pack_mask(m)
    tp = torch.nn.Parameter(torch.tensor(p.copy()))
    opt = torch.optim.Adam([tp], lr=1e-4, betas=(0.9, 0.999), eps=1e-8)
    p0, tm = p.copy(), torch.tensor(m)
    for step in range(1, 6):
        g = (rng.standard_normal(n) * 3).astype(np.float32)
        tn = O.grad_norm(g); c = O.clip_coef(tn, 1.0)
        O.masked_adam_step(p, g, m1, m2, bits, 1e-4, 0.9, 0.999, 1e-8, 0.0, step, c)
        tp.grad = torch.tensor(g.copy())
        tnt = torch.nn.utils.clip_grad_norm_([tp], 1.0)  # DDPM/runners/diffusion.py:582-587
        tp.grad *= tm; opt.step()                         # :589-593
        assert abs(float(tnt) - tn) < 1e-4 * tn
    np.testing.assert_allclose(p, tp.detach().numpy(), rtol=1e-6, atol=1e-7)
    assert np.array_equal(p[m == 0], p0[m == 0])
