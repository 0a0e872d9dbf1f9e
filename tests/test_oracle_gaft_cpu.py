"""oracle/classification.py restatement of the GA / FT / FT_l1 loop bodies against the UNMODIFIED reference
(unlearn.GA GA.py:107-150, unlearn.FT / FT_l1 FT.py:116-180; fixtures: tests/golden/make_golden_gaft.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import classification as OC

G = os.path.join(os.path.dirname(__file__), "golden")


def sample_idx(n, k=64):
    g = np.random.default_rng(n)
    return np.sort(g.choice(n, size=min(k, n), replace=False))


def gaft_inputs():
    g = torch.Generator().manual_seed(11)
    x = torch.rand(32, 3, 32, 32, generator=g)
    y = torch.randint(0, 10, (32,), generator=g)
    g = torch.Generator().manual_seed(13)
    xr = torch.rand(32, 3, 32, 32, generator=g)
    yr = torch.randint(0, 10, (32,), generator=g)
    return x, y, xr, yr


def golden_mask():
    n = 11173962
    zm = np.load(os.path.join(G, "resnet18_mask.npz"))
    flat = torch.from_numpy(np.unpackbits(zm["bits_0.5"], bitorder="little")[:n].astype(np.int64))
    return OC.split_mask(flat, OC.resnet18_param_shapes(10))


def schedule(name, n_batches):
    """[(use_forget_set, loss_sign, l1_alpha)] per step, as the reference loops run them (FT_l1: 2 epochs, alpha decays)."""
    if name == "GA":
        return [(True, -1.0, 0.0)] * n_batches
    if name == "FT":
        return [(False, 1.0, 0.0)] * n_batches
    return [(False, 1.0, 5e-4 * (1 - e / 2)) for e in range(2) for _ in range(n_batches)]


@pytest.mark.parametrize("name", ["GA", "FT", "FT_l1"])
def test_oracle_method_equals_reference(name):
    torch.set_num_threads(8)
    z = np.load(os.path.join(G, "resnet18_gaft.npz"))
    mask = golden_mask()
    x, y, xr, yr = gaft_inputs()
    for tag, nb, rtol, atol in (("1", 1, 1e-5, 1e-6), ("2", 2, 1e-2, 5e-4)):
        params, buffers = OC.synth_state(10, seed=0)
        p0 = {k: v.clone() for k, v in params.items()}
        opt = OC.MaskedSGD(params, mask, lr=0.013, momentum=0.9, wd=5e-4)
        steps = schedule(name, nb)
        for i, (forget, sign, alpha) in enumerate(steps):
            j = i % nb
            xs, ys = (x, y) if forget else (xr, yr)
            OC.unlearn_step(params, buffers, opt, xs[16 * j:16 * j + 16], ys[16 * j:16 * j + 16], sign=sign, l1_alpha=alpha)
        if name == "FT_l1" and tag == "1":
            rtol, atol = 1e-4, 2e-6      # two chained steps
        if name == "FT_l1" and tag == "2":
            rtol, atol = 2e-2, 2e-3      # four chained steps at random init: 1-ulp differences are amplified (see RL test)
        ps = np.concatenate([t.flatten()[sample_idx(t.numel())].numpy() for t in params.values()])
        np.testing.assert_allclose(ps, z[f"{name}_{tag}_psample"], rtol=rtol, atol=atol)
        np.testing.assert_allclose(buffers["bn1.running_mean"].numpy(), z[f"{name}_{tag}_rm_bn1"], rtol=1e-3, atol=5e-5)
        for k in params:
            assert torch.equal(params[k][mask[k] == 0], p0[k][mask[k] == 0])
