"""Configuration and parameter table of the class-conditional DDPM U-Net (DDPM/models/diffusion.py:195-338) -- plain
Python, no torch modules: the network itself lives in libsalun (csrc/salun_unet.cu); the PyTorch restatement that the
tests check it against is test infrastructure (oracle/unet.py)."""
from __future__ import annotations

from collections import OrderedDict
from types import SimpleNamespace


def cifar10_config(n_classes: int = 10, dropout: float = 0.1, cond_drop_prob: float = 0.1) -> SimpleNamespace:
    """DDPM/configs/cifar10_saliency_unlearn.yml:1-57 (model / data / diffusion keys the network reads)"""
    return SimpleNamespace(
        model=SimpleNamespace(type="conditional", in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 2, 2], num_res_blocks=2,
                              attn_resolutions=[16], dropout=dropout, resamp_with_conv=True, cond_drop_prob=cond_drop_prob),
        data=SimpleNamespace(image_size=32, channels=3, n_classes=n_classes),
        diffusion=SimpleNamespace(beta_schedule="linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000),
    )


def unet_param_table(config) -> "OrderedDict[str, tuple]":
    """named_parameters() order and PyTorch shapes of the reference's Conditional_Model: ``null_classes_emb`` first
    (a direct nn.Parameter), then temb / classes_emb / cemb / conv_in, the down levels (blocks, attentions, downsample),
    mid, the up levels in INDEX order (up.0 first although it runs last), norm_out, conv_out.  334 tensors and
    38 632 323 parameters for the cifar10 config."""
    m, d = config.model, config.data
    ch, mult, nrb = m.ch, tuple(m.ch_mult), m.num_res_blocks
    emb = 4 * ch
    t: "OrderedDict[str, tuple]" = OrderedDict()

    def dense(pre, d_in):
        t[pre + ".dense.0.weight"], t[pre + ".dense.0.bias"] = (emb, d_in), (emb,)
        t[pre + ".dense.1.weight"], t[pre + ".dense.1.bias"] = (emb, emb), (emb,)

    def conv(pre, c_in, c_out, k):
        t[pre + ".weight"], t[pre + ".bias"] = (c_out, c_in, k, k), (c_out,)

    def norm(pre, c):
        t[pre + ".weight"], t[pre + ".bias"] = (c,), (c,)

    def resblock(pre, c_in, c_out):
        norm(pre + ".norm1", c_in)
        conv(pre + ".conv1", c_in, c_out, 3)
        # the reference leaves cemb_channels at its default 512 (diffusion.py:93,106-108)
        t[pre + ".temb_cemb_proj.weight"], t[pre + ".temb_cemb_proj.bias"] = (c_out, emb + 512), (c_out,)
        norm(pre + ".norm2", c_out)
        conv(pre + ".conv2", c_out, c_out, 3)
        if c_in != c_out:
            conv(pre + ".nin_shortcut", c_in, c_out, 1)

    def attn(pre, c):
        norm(pre + ".norm", c)
        for nm in ("q", "k", "v", "proj_out"):
            conv(pre + "." + nm, c, c, 1)

    t["null_classes_emb"] = (ch,)
    dense("temb", ch)
    t["classes_emb.weight"] = (d.n_classes, ch)
    dense("cemb", ch)
    conv("conv_in", m.in_channels, ch, 3)
    res, in_mult = d.image_size, (1,) + mult
    c = ch
    for lvl in range(len(mult)):
        c, c_out = ch * in_mult[lvl], ch * mult[lvl]
        has_attn = res in m.attn_resolutions
        for i in range(nrb):
            resblock(f"down.{lvl}.block.{i}", c, c_out)
            c = c_out
        if has_attn:
            for i in range(nrb):
                attn(f"down.{lvl}.attn.{i}", c)
        if lvl != len(mult) - 1:
            conv(f"down.{lvl}.downsample.conv", c, c, 3)
            res //= 2
    resblock("mid.block_1", c, c)
    attn("mid.attn_1", c)
    resblock("mid.block_2", c, c)
    ups = {}
    for lvl in reversed(range(len(mult))):
        entries = []
        c_out, skip = ch * mult[lvl], ch * mult[lvl]
        has_attn = res in m.attn_resolutions
        for i in range(nrb + 1):
            if i == nrb:
                skip = ch * in_mult[lvl]
            entries.append(("block", i, c + skip, c_out))
            c = c_out
        if has_attn:
            for i in range(nrb + 1):
                entries.append(("attn", i, c, c))
        if lvl != 0:
            entries.append(("upsample", 0, c, c))
            res *= 2
        ups[lvl] = entries
    for lvl in range(len(mult)):
        for kind, i, a, b in ups[lvl]:
            if kind == "block":
                resblock(f"up.{lvl}.block.{i}", a, b)
            elif kind == "attn":
                attn(f"up.{lvl}.attn.{i}", a)
            else:
                conv(f"up.{lvl}.upsample.conv", a, a, 3)
    norm("norm_out", c)
    conv("conv_out", c, m.out_ch, 3)
    return t
