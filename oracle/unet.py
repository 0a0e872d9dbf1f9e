"""TEST INFRASTRUCTURE (checker network; only tests/, __graft_entry__.smoke() and bench.py's CPU / torch baseline legs import
it -- the product path runs the U-Net in libsalun, csrc/salun_unet.cu).

Class-conditional DDPM U-Net with the parameter names / registration order of the reference's
``Conditional_Model`` (DDPM/models/diffusion.py:195-413), so reference checkpoints (``states[0]``, with or without the
DataParallel ``module.`` prefix) load unchanged and masks keyed by ``named_parameters()`` line up (SURVEY.md A.3:
334 tensors, 38 632 323 parameters for the cifar10 config, ``null_classes_emb`` first).

Restated from the architecture description, not copied: sinusoidal timestep embedding (sin || cos,
log(10000)/(half-1), diffusion.py:17-35) -> 2 Linear; class embedding (replaced by a learned null embedding with
probability cond_drop_prob, :372-376) -> 2 Linear; conv_in; per level `num_res_blocks` residual blocks
(GroupNorm32(eps 1e-6) -> swish -> conv3x3 -> + Linear(swish([temb, cemb])) -> GroupNorm -> swish -> dropout ->
conv3x3, 1x1 shortcut when channels change, :124-145) with single-head spatial self-attention at the configured
resolutions (:167-192), strided-conv downsample with (0,1,0,1) padding (:75-79), nearest x2 upsample + conv, skip
concatenation on the way up, GroupNorm -> swish -> conv_out.
"""
from __future__ import annotations

import math
import torch
import torch.nn.functional as F
from torch import nn


from unlearn_saliency_b200.diffusion.config import cifar10_config  # noqa: E402,F401  (re-exported for the tests)


def swish(x):
    return x * torch.sigmoid(x)


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    half = dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32, device=t.device) * -(math.log(10000) / (half - 1)))
    ang = t.float()[:, None] * freq[None, :]
    emb = torch.cat([ang.sin(), ang.cos()], dim=1)
    return F.pad(emb, (0, 1)) if dim % 2 else emb


def _gn(c):
    return nn.GroupNorm(32, c, eps=1e-6, affine=True)


class _Dense(nn.Module):
    def __init__(self, d_in, d_hidden):
        super().__init__()
        self.dense = nn.ModuleList([nn.Linear(d_in, d_hidden), nn.Linear(d_hidden, d_hidden)])

    def forward(self, x):
        return self.dense[1](swish(self.dense[0](x)))


class ResBlock(nn.Module):
    def __init__(self, c_in, c_out, emb_ch, dropout):
        super().__init__()
        self.norm1 = _gn(c_in)
        self.conv1 = nn.Conv2d(c_in, c_out, 3, padding=1)
        # the reference leaves cemb_channels at its default 512 (diffusion.py:93,106-108): only ch = 128 is consistent
        self.temb_cemb_proj = nn.Linear(emb_ch + 512, c_out)
        self.norm2 = _gn(c_out)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(c_out, c_out, 3, padding=1)
        if c_in != c_out:
            self.nin_shortcut = nn.Conv2d(c_in, c_out, 1)

    def forward(self, x, emb_act):
        h = self.conv1(swish(self.norm1(x)))
        h = h + self.temb_cemb_proj(emb_act)[:, :, None, None]
        h = self.conv2(self.dropout(swish(self.norm2(h))))
        if hasattr(self, "nin_shortcut"):
            x = self.nin_shortcut(x)
        return x + h


class SelfAttention(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.norm = _gn(c)
        self.q, self.k, self.v, self.proj_out = (nn.Conv2d(c, c, 1) for _ in range(4))

    def forward(self, x):
        b, c, hh, ww = x.shape
        h = self.norm(x)
        q = self.q(h).flatten(2).transpose(1, 2)          # b, hw, c
        k = self.k(h).flatten(2)                          # b, c, hw
        v = self.v(h).flatten(2)                          # b, c, hw
        w = torch.softmax(torch.bmm(q, k) * (int(c) ** -0.5), dim=2)   # b, hw(q), hw(k)
        out = torch.bmm(v, w.transpose(1, 2)).reshape(b, c, hh, ww)
        return x + self.proj_out(out)


class _Down(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1)))


class _Up(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class ConditionalUNet(nn.Module):
    def __init__(self, config):
        super().__init__()
        m = config.model
        if not m.resamp_with_conv:
            raise NotImplementedError("resamp_with_conv=False is not used by the SalUn configs")
        self.config = config
        self.ch, self.resolution = m.ch, config.data.image_size
        self.cond_drop_prob = m.cond_drop_prob
        ch, mult, nrb = m.ch, tuple(m.ch_mult), m.num_res_blocks
        emb = 4 * ch
        self.null_classes_emb = nn.Parameter(torch.randn(ch))   # a direct parameter: first in named_parameters()
        self.temb = _Dense(ch, emb)
        self.classes_emb = nn.Embedding(config.data.n_classes, ch)
        self.cemb = _Dense(ch, emb)
        self.conv_in = nn.Conv2d(m.in_channels, ch, 3, padding=1)
        res, in_mult = self.resolution, (1,) + mult
        self.down = nn.ModuleList()
        c = None
        for lvl in range(len(mult)):
            level = nn.Module()
            level.block, level.attn = nn.ModuleList(), nn.ModuleList()
            c, c_out = ch * in_mult[lvl], ch * mult[lvl]
            for _ in range(nrb):
                level.block.append(ResBlock(c, c_out, emb, m.dropout))
                c = c_out
                if res in m.attn_resolutions:
                    level.attn.append(SelfAttention(c))
            if lvl != len(mult) - 1:
                level.downsample = _Down(c)
                res //= 2
            self.down.append(level)
        self.mid = nn.Module()
        self.mid.block_1 = ResBlock(c, c, emb, m.dropout)
        self.mid.attn_1 = SelfAttention(c)
        self.mid.block_2 = ResBlock(c, c, emb, m.dropout)
        ups = []
        for lvl in reversed(range(len(mult))):
            level = nn.Module()
            level.block, level.attn = nn.ModuleList(), nn.ModuleList()
            c_out, skip = ch * mult[lvl], ch * mult[lvl]
            for i in range(nrb + 1):
                if i == nrb:
                    skip = ch * in_mult[lvl]
                level.block.append(ResBlock(c + skip, c_out, emb, m.dropout))
                c = c_out
                if res in m.attn_resolutions:
                    level.attn.append(SelfAttention(c))
            if lvl != 0:
                level.upsample = _Up(c)
                res *= 2
            ups.insert(0, level)
        self.up = nn.ModuleList(ups)
        self.norm_out = _gn(c)
        self.conv_out = nn.Conv2d(c, m.out_ch, 3, padding=1)

    # -- reference call surface: model(x, t, c, cond_scale=s, mode="test") / model(x, t, c, mode="train", cond_drop_prob=p)
    def forward(self, x, t, c, mode, **kw):
        if mode == "train":
            return self._forward(x, t, c, cond_drop_prob=kw.get("cond_drop_prob"), drop_mask=kw.get("drop_mask"))
        if mode == "test":
            s = kw.get("cond_scale")
            cond = self._forward(x, t, c, cond_drop_prob=0.0)
            if s == 0:
                return cond
            null = self._forward(x, t, c, cond_drop_prob=1.0)
            return (1 + s) * cond - s * null          # diffusion.py:340-355
        raise AssertionError("mode must be 'train' or 'test'")

    def _forward(self, x, t, c, cond_drop_prob=None, drop_mask=None):
        assert x.shape[2] == x.shape[3] == self.resolution
        p = self.cond_drop_prob if cond_drop_prob is None else cond_drop_prob
        temb = self.temb(timestep_embedding(t, self.ch))
        ce = self.classes_emb(c)
        if drop_mask is not None:      # externally supplied class-dropout decisions (parity runs)
            keep = ~drop_mask
            ce = torch.where(keep[:, None], ce, self.null_classes_emb[None, :].expand_as(ce))
        elif p > 0:
            if p == 1:
                keep = torch.zeros(x.shape[0], dtype=torch.bool, device=x.device)
            else:
                keep = torch.zeros(x.shape[0], device=x.device).float().uniform_(0, 1) < (1 - p)
            ce = torch.where(keep[:, None], ce, self.null_classes_emb[None, :].expand_as(ce))
        cemb = self.cemb(ce)
        emb_act = swish(torch.cat([temb, cemb], dim=-1))   # every block applies swish to [temb, cemb] before its Linear
        hs = [self.conv_in(x)]
        for lvl, level in enumerate(self.down):
            for i, blk in enumerate(level.block):
                h = blk(hs[-1], emb_act)
                if len(level.attn):
                    h = level.attn[i](h)
                hs.append(h)
            if lvl != len(self.down) - 1:
                hs.append(level.downsample(hs[-1]))
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(hs[-1], emb_act)), emb_act)
        for lvl in reversed(range(len(self.up))):
            level = self.up[lvl]
            for i, blk in enumerate(level.block):
                h = blk(torch.cat([h, hs.pop()], dim=1), emb_act)
                if len(level.attn):
                    h = level.attn[i](h)
            if lvl != 0:
                h = level.upsample(h)
        return self.conv_out(swish(self.norm_out(h)))
