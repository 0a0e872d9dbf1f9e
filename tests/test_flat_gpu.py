"""Fused DDPM/SD tail (unlearn_saliency_b200/flat.py) against the reference's own statements executed with PyTorch:
clip_grad_norm_ -> grad *= mask -> Adam.step()  (DDPM/runners/diffusion.py:582-593)  and
gradients[name] += clip(grad) ; abs ; argsort-argsort mask  (runners/diffusion.py:985-1039)."""
import copy

import numpy as np
import pytest
import torch

from oracle import tail as OT

pytestmark = pytest.mark.gpu


class Net(torch.nn.Module):
    """GroupNorm + conv + linear toy with odd-sized tensors (exercises the arena alignment padding)"""

    def __init__(self):
        super().__init__()
        self.conv_in = torch.nn.Conv2d(3, 14, 3, padding=1)
        self.norm = torch.nn.GroupNorm(2, 14, eps=1e-6)
        self.conv_out = torch.nn.Conv2d(14, 3, 3, padding=1)
        self.emb = torch.nn.Linear(5, 14)

    def forward(self, x, t):
        h = self.conv_in(x) + self.emb(t)[:, :, None, None]
        h = self.norm(h)
        return self.conv_out(h * torch.sigmoid(h))


def _batch(seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(6, 3, 8, 8, generator=g).cuda(), torch.randn(6, 5, generator=g).cuda(), torch.randn(6, 3, 8, 8, generator=g).cuda()


def test_masked_clipped_adam_matches_reference_statements(salun_ctx):
    from unlearn_saliency_b200.flat import FlatMaskedAdam, FlatParams
    torch.manual_seed(0)
    ref = Net().cuda()
    mine = copy.deepcopy(ref)
    g = torch.Generator().manual_seed(1)
    mask = {n: (torch.rand(p.shape, generator=g) < 0.5).to(torch.int64) for n, p in ref.named_parameters()}  # CPU int64, as torch.load gives
    opt_ref = torch.optim.Adam(ref.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    flat = FlatParams(mine, salun_ctx)
    opt = FlatMaskedAdam(flat, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, mask=mask, max_norm=1.0)
    p0 = {n: p.detach().clone() for n, p in ref.named_parameters()}
    for s in range(5):
        x, t, e = _batch(10 + s)
        # reference statements
        loss = (e - ref(x, t)).square().sum(dim=(1, 2, 3)).mean(dim=0)  # losses.py:21-37 shape of the eps-loss
        opt_ref.zero_grad()
        loss.backward()
        tn = torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)       # runners/diffusion.py:582-587
        for n, p in ref.named_parameters():
            p.grad *= mask[n].to(p.device)                               # :589-592
        opt_ref.step()                                                   # :593
        # fused tail
        loss2 = (e - mine(x, t)).square().sum(dim=(1, 2, 3)).mean(dim=0)
        opt.zero_grad()
        loss2.backward()
        opt.step()
        assert abs(float(opt.grad_norm()) - float(tn)) <= 1e-5 * float(tn)
    for (n, p), (_, q) in zip(ref.named_parameters(), mine.named_parameters()):
        torch.testing.assert_close(q, p, rtol=2e-5, atol=2e-6)
        m = mask[n].cuda()
        assert torch.equal(q[m == 0], p0[n][m == 0])  # masked-out coordinates never move


def test_flat_saliency_mask_matches_reference_formula(salun_ctx, tmp_path):
    from unlearn_saliency_b200.flat import FlatParams, FlatSaliency
    torch.manual_seed(0)
    ref = Net().cuda()
    mine = copy.deepcopy(ref)
    flat = FlatParams(mine, salun_ctx)
    sal = FlatSaliency(flat, max_norm=1.0)
    grads = {n: torch.zeros_like(p, device="cpu") for n, p in ref.named_parameters()}
    for s in range(3):
        x, t, e = _batch(20 + s)
        ref.zero_grad()
        (e - ref(x, t)).square().sum(dim=(1, 2, 3)).mean(dim=0).backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)            # runners/diffusion.py:985-990
        for n, p in ref.named_parameters():
            grads[n] += p.grad.data.cpu()                                # :992-996
        flat.zero_grad()
        (e - mine(x, t)).square().sum(dim=(1, 2, 3)).mean(dim=0).backward()
        sal.accumulate()
    allg = torch.cat([g.abs().flatten() for g in grads.values()]).numpy()
    k = int(allg.size * 0.5)
    info = sal.save(str(tmp_path / "mask" / "with_0.5.pt"), 0.5, key_prefix="module.")
    m = torch.load(str(tmp_path / "mask" / "with_0.5.pt"))
    assert list(m.keys()) == ["module." + n for n in grads]              # DataParallel key layout
    assert all(v.dtype == torch.int64 and not v.is_cuda for v in m.values())
    mine_flat = torch.cat([v.flatten() for v in m.values()]).numpy()
    assert mine_flat.sum() == k
    # accumulated gradients agree to fp32 rounding (separately rounded scale*g vs in-place g*=c), so compare index sets
    ref_mask = OT.topk_mask_argsort(allg, k)
    assert (mine_flat != ref_mask).sum() <= 2
    # and bit-exactly on the engine's own accumulator
    dense = torch.cat([sal.acc[flat.offsets[n]: flat.offsets[n] + p.numel()] for n, p in mine.named_parameters()])
    om, *_ = OT.topk_mask(np.abs(dense.cpu().numpy()), k)
    assert np.array_equal(mine_flat, om)
