"""Forward pass of the Stable-Diffusion (LDM) U-Net on the sm_100a kernels -- SURVEY.md section 8 rows a16 / f1.

``SDUNetEngine`` runs ``UNetModel.forward(x, timesteps, context)`` (SD/ldm/modules/diffusionmodules/openaimodel.py:814-846;
ResBlock :268-288, Downsample / Upsample :87-160, SpatialTransformer / BasicTransformerBlock / CrossAttention / GEGLU
SD/ldm/modules/attention.py:37-66,168-303, timestep_embedding util.py:173-197) as a fixed program of op-level C-ABI calls
(include/salun.h ``salun_op_*`` / ``salun_sd_*``, csrc/salun_ops.cu): tcgen05 implicit-GEMM convolutions and Linears with
bias / time-embedding / residual fused into their epilogues, GroupNorm + SiLU, LayerNorm, GEGLU and multi-head self / cross
attention kernels.  The program is built once per batch size and replayed from a CUDA graph, so the ~700 calls of one forward
cost one launch on the host.  No autograd: this is the no-grad path of the ESD loop -- the DDIM partial sampling with
classifier-free guidance and the frozen-model passes, about 90 % of its FLOPs (SD/train-scripts/train-esd.py:287-300); the one
trainable forward + backward per iteration stays with the reference's module.

Parameters keep the reference's ``named_parameters()`` names and PyTorch layouts (``load_state_dict`` takes the
``model.model.diffusion_model`` state dict as is); the tensor-core weight operands are prepared from them at load time.
PyTorch owns the device memory and the stream; there is no fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence

import torch

from .. import _lib
from .._lib import check
from ..tail import SalunContext, _ptr, _stream


def sd_v1_config() -> dict:
    """SD/configs/stable-diffusion/v1-inference.yaml:29-44 (unet_config.params)"""
    return dict(in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2,
                channel_mult=[1, 2, 4, 4], num_heads=8, transformer_depth=1, context_dim=768)


def _pad64(c: int) -> int:
    return (c + 63) // 64 * 64


class _Arch:
    """walks UNetModel.__init__ (openaimodel.py:466-800; legacy=False, use_spatial_transformer=True, conv_resample=True) and
    records (a) the parameter table in named_parameters() order, (b) the block structure the program builder consumes"""

    def __init__(self, cfg: dict):
        self.cfg = cfg
        self.table: "OrderedDict[str, tuple]" = OrderedDict()
        mc, heads, dc = cfg["model_channels"], cfg["num_heads"], cfg["context_dim"]
        ted = 4 * mc
        self.time_embed_dim = ted
        t = self.table
        t["time_embed.0.weight"], t["time_embed.0.bias"] = (ted, mc), (ted,)
        t["time_embed.2.weight"], t["time_embed.2.bias"] = (ted, ted), (ted,)

        def conv(pre, cin, cout, k):
            t[pre + ".weight"], t[pre + ".bias"] = (cout, cin, k, k), (cout,)

        def norm(pre, c):
            t[pre + ".weight"], t[pre + ".bias"] = (c,), (c,)

        def lin(pre, cin, cout, bias=True):
            t[pre + ".weight"] = (cout, cin)
            if bias:
                t[pre + ".bias"] = (cout,)

        def resblock(pre, cin, cout):
            norm(pre + ".in_layers.0", cin)
            conv(pre + ".in_layers.2", cin, cout, 3)
            lin(pre + ".emb_layers.1", ted, cout)
            norm(pre + ".out_layers.0", cout)
            conv(pre + ".out_layers.3", cout, cout, 3)
            if cin != cout:
                conv(pre + ".skip_connection", cin, cout, 1)
            return ("res", pre, cin, cout)

        def transformer(pre, c):
            d = c // heads
            norm(pre + ".norm", c)
            conv(pre + ".proj_in", c, c, 1)
            for i in range(cfg.get("transformer_depth", 1)):
                b = f"{pre}.transformer_blocks.{i}"
                for a, kd in (("attn1", c), ("ff", None), ("attn2", dc)):
                    if a == "ff":
                        lin(b + ".ff.net.0.proj", c, 8 * c)
                        lin(b + ".ff.net.2", 4 * c, c)
                    else:
                        lin(f"{b}.{a}.to_q", c, c, bias=False)
                        lin(f"{b}.{a}.to_k", kd, c, bias=False)
                        lin(f"{b}.{a}.to_v", kd, c, bias=False)
                        lin(f"{b}.{a}.to_out.0", c, c)
                for k in (1, 2, 3):
                    norm(f"{b}.norm{k}", c)
            conv(pre + ".proj_out", c, c, 1)
            return ("attn", pre, c, d)

        self.input_blocks: List[list] = []
        conv("input_blocks.0.0", cfg["in_channels"], mc, 3)
        self.input_blocks.append([("conv_in", "input_blocks.0.0", cfg["in_channels"], mc)])
        chans, ch, ds = [mc], mc, 1
        nrb = cfg["num_res_blocks"]
        for level, mult in enumerate(cfg["channel_mult"]):
            for _ in range(nrb):
                i = len(self.input_blocks)
                layers = [resblock(f"input_blocks.{i}.0", ch, mult * mc)]
                ch = mult * mc
                if ds in cfg["attention_resolutions"]:
                    layers.append(transformer(f"input_blocks.{i}.1", ch))
                self.input_blocks.append(layers)
                chans.append(ch)
            if level != len(cfg["channel_mult"]) - 1:
                i = len(self.input_blocks)
                conv(f"input_blocks.{i}.0.op", ch, ch, 3)
                self.input_blocks.append([("down", f"input_blocks.{i}.0.op", ch, ch)])
                chans.append(ch)
                ds *= 2
        self.middle = [resblock("middle_block.0", ch, ch), transformer("middle_block.1", ch), resblock("middle_block.2", ch, ch)]
        self.output_blocks: List[list] = []
        for level, mult in list(enumerate(cfg["channel_mult"]))[::-1]:
            for i in range(nrb + 1):
                ich = chans.pop()
                j = len(self.output_blocks)
                layers = [resblock(f"output_blocks.{j}.0", ch + ich, mc * mult)]
                ch = mc * mult
                if ds in cfg["attention_resolutions"]:
                    layers.append(transformer(f"output_blocks.{j}.{len(layers)}", ch))
                if level and i == nrb:
                    pre = f"output_blocks.{j}.{len(layers)}.conv"
                    conv(pre, ch, ch, 3)
                    layers.append(("up", pre, ch, ch))
                    ds //= 2
                self.output_blocks.append(layers)
        norm("out.0", ch)
        conv("out.2", mc, cfg["out_channels"], 3)
        self.final_ch = ch


def sd_unet_param_table(cfg: dict) -> "OrderedDict[str, tuple]":
    """named_parameters() order and shapes of the reference's UNetModel (686 tensors / 859 520 964 parameters for SD v1.4)"""
    return _Arch(cfg).table


def config_from_state_dict(sd: Dict[str, "torch.Tensor"], num_heads: int = 8) -> dict:
    """unet_config.params of a UNetModel from its parameter names and shapes (keys may carry the LatentDiffusion prefix
    `model.diffusion_model.`).  The head count is not visible in the shapes (to_q is C x C for any split): pass it, or use
    config_from_module.  The result is checked against the full parameter table before it is returned."""
    shp = {k.split("model.diffusion_model.")[-1]: tuple(v.shape) for k, v in sd.items()}
    mc, cin = shp["input_blocks.0.0.weight"][:2]
    mults, attn_res, nrb, ds, level_blocks, i = [], [], None, 1, 0, 1
    depth, ctx_dim = 1, None
    while f"input_blocks.{i}.0.in_layers.2.weight" in shp or f"input_blocks.{i}.0.op.weight" in shp:
        if f"input_blocks.{i}.0.op.weight" in shp:
            nrb = level_blocks if nrb is None else nrb
            ds, level_blocks = ds * 2, 0
        else:
            mult = shp[f"input_blocks.{i}.0.in_layers.2.weight"][0] // mc
            if level_blocks == 0:
                mults.append(mult)
            level_blocks += 1
            if f"input_blocks.{i}.1.norm.weight" in shp:
                if ds not in attn_res:
                    attn_res.append(ds)
                pre = f"input_blocks.{i}.1.transformer_blocks."
                depth = max(depth, 1 + max(int(k[len(pre):].split(".")[0]) for k in shp if k.startswith(pre)))
                ctx_dim = shp[pre + "0.attn2.to_k.weight"][1]
        i += 1
    nrb = level_blocks if nrb is None else nrb
    if ctx_dim is None:
        ctx_dim = shp["middle_block.1.transformer_blocks.0.attn2.to_k.weight"][1]
    cfg = dict(in_channels=cin, out_channels=shp["out.2.weight"][0], model_channels=mc, attention_resolutions=attn_res,
               num_res_blocks=nrb, channel_mult=mults, num_heads=num_heads, transformer_depth=depth, context_dim=ctx_dim)
    table = sd_unet_param_table(cfg)
    if set(table) != set(shp) or any(tuple(table[k]) != shp[k] for k in table):
        bad = [k for k in table if shp.get(k) != tuple(table[k])][:3] + [k for k in shp if k not in table][:3]
        raise ValueError(f"state dict is not a UNetModel this engine covers (use_spatial_transformer, conv_resample): {bad}")
    return cfg


def config_from_module(unet) -> dict:
    """the same from a live UNetModel (openaimodel.py:466-560 keeps num_heads as an attribute)"""
    return config_from_state_dict(dict(unet.named_parameters()), num_heads=int(getattr(unet, "num_heads", 8)))


class SDUNetEngine:
    """eps = UNetModel(x, timesteps, context) without autograd, on libsalun's op-level entry points."""

    def __init__(self, cfg: Optional[dict] = None, latent_size: int = 64, max_batch: int = 2, context_len: int = 77,
                 device=None, ctx: Optional[SalunContext] = None, precision: str = "bf16", use_graph: bool = True):
        self.cfg = dict(sd_v1_config() if cfg is None else cfg)
        self.arch = _Arch(self.cfg)
        self.table = self.arch.table
        self.S, self.max_batch, self.L = int(latent_size), int(max_batch), int(context_len)
        self.ctx = ctx if ctx is not None else SalunContext(device)
        self.device = self.ctx.device
        self.precision = precision
        self._lib = _lib.lib(precision)
        self.act_bytes = int(self._lib.salun_act_bytes())
        self.wop_k = int(self._lib.salun_wop_k())
        self.use_graph = use_graph
        self.ctx.ensure_op_scratch()
        mc, heads = self.cfg["model_channels"], self.cfg["num_heads"]
        if mc % 64 or self.cfg["context_dim"] % 64 or (mc // heads) % 8:
            raise ValueError("model_channels and context_dim must be multiples of 64, the head width a multiple of 8")
        nlev = len(self.cfg["channel_mult"])
        if self.S & (self.S - 1) or (self.S >> (nlev - 1)) < 4:
            raise ValueError("latent_size must be a power of two with at least 4x4 at the deepest level")
        self.params: Dict[str, torch.Tensor] = {}
        self._wops: Dict[str, torch.Tensor] = {}
        self._programs: Dict[int, dict] = {}
        self.n_params = sum(math.prod(s) for s in self.table.values())

    # ---- parameters ------------------------------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        sd = {k.split("model.diffusion_model.")[-1]: v for k, v in sd.items()}
        missing = [k for k in self.table if k not in sd]
        if missing and strict:
            raise KeyError(f"missing keys: {missing[:4]} ... ({len(missing)})")
        for k, shp in self.table.items():
            if k in sd:
                if tuple(sd[k].shape) != tuple(shp):
                    raise ValueError(f"{k}: shape {tuple(sd[k].shape)} != {shp}")
                self.params[k] = sd[k].detach().to(self.device, torch.float32).contiguous()
            elif k not in self.params:
                self.params[k] = torch.zeros(shp, device=self.device)
        self._prepare_weights()
        self._programs.clear()
        return self

    @torch.no_grad()
    def update_parameters(self, sd: Dict[str, torch.Tensor]):
        """In-place refresh after an optimizer step on the caller's module (ESD samples from the model it trains,
        train-esd.py:262-283): same storage, so captured graphs stay valid; only the given keys are re-prepared."""
        if not self.params:
            return self.load_state_dict(sd)
        touched = []
        for k, v in sd.items():
            k = k.split("model.diffusion_model.")[-1]
            if k not in self.table:
                raise KeyError(k)
            self.params[k].copy_(v.detach().reshape(self.table[k]), non_blocking=True)
            touched.append(k)
        for k in touched:
            if k in self._wops:
                shp = self.table[k]
                self._prep(k, shp[0], shp[1], shp[2] if len(shp) == 4 else 1)
            elif ".emb_layers.1." in k:
                self._stack_emb(k.split(".emb_layers.1.")[0])
        return self

    def _prep(self, name: str, cout: int, cin: int, ks: int):
        """tensor-core operand of one Conv2d / Linear weight (zero padded to multiples of 64)"""
        cp, kp = _pad64(cout), _pad64(cin)
        w = self.params[name]
        buf = self._wops.get(name)
        if buf is None:
            buf = torch.empty(cp * ks * ks * kp * self.wop_k, dtype=torch.bfloat16, device=self.device)
        check(self._lib.salun_op_prep_weight(self.ctx.handle, _ptr(w), _ptr(buf), cout, cin, ks, cp, kp, _stream(self.device)),
              "salun_op_prep_weight")
        self._wops[name] = buf

    def _prepare_weights(self):
        # to_q | to_k | to_v of a self-attention (to_k | to_v of a cross-attention) read the same rows: their operands are
        # slices of ONE stacked buffer, so the projections run as one GEMM with N = 3C (2C); _prep fills the slices in place
        if not hasattr(self, "_stack"):
            self._stack: Dict[str, tuple] = {}
        for k, shp in self.table.items():
            if k.endswith(".to_q.weight") and k[:-len(".to_q.weight")] not in self._stack:
                pre = k[:-len(".to_q.weight")]
                C_, kd = shp[0], self.table[pre + ".to_k.weight"][1]
                names = [pre + ".to_q.weight", pre + ".to_k.weight", pre + ".to_v.weight"] if kd == C_ else \
                        [pre + ".to_k.weight", pre + ".to_v.weight"]
                per = C_ * _pad64(kd) * self.wop_k
                stack = torch.empty(per * len(names), dtype=torch.bfloat16, device=self.device)
                for i, nm in enumerate(names):
                    self._wops[nm] = stack[i * per:(i + 1) * per]
                self._stack[pre] = ("qkv" if kd == C_ else "kv", stack, per)
        for k, shp in self.table.items():
            if not k.endswith(".weight") or len(shp) < 2 or k.startswith("time_embed") or ".emb_layers." in k:
                continue      # norms and the fp32 embedding MLPs keep their fp32 weights
            self._prep(k, shp[0], shp[1], shp[2] if len(shp) == 4 else 1)
        # every ResBlock's emb_layers Linear reads the same SiLU(emb): one stacked [sum cout][time_embed_dim] weight, one launch
        self._emb_off, off = {}, 0
        for k, shp in self.table.items():
            if k.endswith(".emb_layers.1.weight"):
                self._emb_off[k[:-len(".emb_layers.1.weight")]] = (off, shp[0])
                off += shp[0]
        self._emb_total = off
        if getattr(self, "_emb_w", None) is None:
            self._emb_w = torch.empty(off, self.arch.time_embed_dim, device=self.device)
            self._emb_b = torch.empty(off, device=self.device)
        for pre, (o, c) in self._emb_off.items():
            self._stack_emb(pre)

    def _stack_emb(self, pre: str):
        o, c = self._emb_off[pre]
        self._emb_w[o:o + c].copy_(self.params[pre + ".emb_layers.1.weight"])
        self._emb_b[o:o + c].copy_(self.params[pre + ".emb_layers.1.bias"])

    # ---- buffers -----------------------------------------------------------------------------------------------------
    def _act(self, elems: int) -> torch.Tensor:
        return torch.zeros(elems * self.act_bytes, dtype=torch.uint8, device=self.device)   # zero: halos, channel padding

    def _padded(self, n, H, C):
        return self._act(n * (H + 2) * (H + 2) * C)

    def _flat(self, rows, C):
        return self._act(rows * C)

    # ---- program -----------------------------------------------------------------------------------------------------
    def _build(self, n: int) -> dict:
        L, lib, h, dev = self._lib, self._lib, self.ctx.handle, self.device
        P, W = self.params, self._wops
        ops: list = []
        st = lambda: _stream(dev)
        keep: list = []    # every buffer of the program stays alive with it

        def run(fn, *a, what=""):
            ops.append((fn, a, what))

        def buf(t):
            keep.append(t)
            return t

        cfg, S = self.cfg, self.S
        mc, ted, heads, dc, Lc = cfg["model_channels"], self.arch.time_embed_dim, cfg["num_heads"], cfg["context_dim"], self.L
        x_in = buf(torch.zeros(n, cfg["in_channels"], S, S, device=dev))
        t_in = buf(torch.zeros(n, device=dev))
        c_in = buf(torch.zeros(n, Lc, dc, device=dev))
        eps_out = buf(torch.zeros(n, cfg["out_channels"], S, S, device=dev))
        # time embedding: timestep_embedding -> Linear -> SiLU -> Linear   (openaimodel.py:833-834)
        temb0 = buf(torch.zeros(n, mc, device=dev))
        temb1 = buf(torch.zeros(n, ted, device=dev))
        emb = buf(torch.zeros(n, ted, device=dev))
        tmp = buf(torch.zeros(n, ted, device=dev))
        run(L.salun_sd_timestep_embedding, h, _ptr(t_in), _ptr(temb0), n, mc, 10000.0, what="timestep_embedding")
        run(L.salun_op_linear_f32, h, _ptr(temb0), _ptr(P["time_embed.0.weight"]), _ptr(P["time_embed.0.bias"]), _ptr(temb1), None, n,
            mc, ted, 0, what="time_embed.0")
        run(L.salun_op_linear_f32, h, _ptr(temb1), _ptr(P["time_embed.2.weight"]), _ptr(P["time_embed.2.bias"]), _ptr(emb), _ptr(tmp),
            n, ted, ted, 1, what="time_embed.2")
        emb_all = buf(torch.zeros(n, self._emb_total, device=dev))
        run(L.salun_op_linear_f32, h, _ptr(emb), _ptr(self._emb_w), _ptr(self._emb_b), _ptr(emb_all), _ptr(tmp), n, ted,
            self._emb_total, 1, what="emb_layers (all ResBlocks)")
        ctx_act = buf(self._flat(n * Lc, dc))
        run(L.salun_op_f32_to_act, h, _ptr(c_in), dc, _ptr(ctx_act), dc, n * Lc, dc, what="context")
        stats = buf(torch.zeros(int(L.salun_op_groupnorm_ws_floats(n)), device=dev))
        attn_ws: Dict[tuple, torch.Tensor] = {}   # one attention workspace per (Tq, Tk, d) shape, shared by the layers
        out_ch = lambda layer: layer[2] if layer[0] == "attn" else layer[3]

        def conv(x, in_flat, name, cin, cout, ks, H, bias=True, rowbias=None, addend=None, out_pad=True, out_f32=None):
            cp, kp = _pad64(cout), _pad64(cin)
            out = None
            if out_f32 is None:
                out = buf(self._padded(n, H, cp) if out_pad else self._flat(n * H * H, cp))
            b = _ptr(P[name[:-len(".weight")] + ".bias"]) if bias and cout == cp else None
            rb, rb_ld = (None, 0) if rowbias is None else rowbias      # (pointer, row stride in floats)
            run(L.salun_op_conv, h, _ptr(x), 1 if in_flat else 0, _ptr(W[name]), b, rb, rb_ld,
                _ptr(addend), _ptr(out), 1 if (out_pad and out_f32 is None) else 0, _ptr(out_f32), n, H, H, kp, cp, ks, what=name)
            return out

        def linear_rows(x, rows, name, cin, cout, bias=True, addend=None, wop=None):
            out = buf(self._flat(rows, cout))
            b = _ptr(P[name[:-len(".weight")] + ".bias"]) if bias else None
            run(L.salun_op_conv, h, _ptr(x), 1, _ptr(W[name] if wop is None else wop), b, None, 0, _ptr(addend), _ptr(out), 0, None,
                rows, 1, 1, cin, cout, 1, what=name)
            return out

        def col_slice(t, col):     # pointer to column `col` of a flat activation matrix
            return C.c_void_p(t.data_ptr() + col * self.act_bytes)

        def groupnorm(x, pre, C, H, eps, swish, out_flat=False):
            out = buf(self._flat(n * H * H, C) if out_flat else self._padded(n, H, C))
            run(L.salun_op_groupnorm, h, _ptr(x), _ptr(P[pre + ".weight"]), _ptr(P[pre + ".bias"]), _ptr(stats), _ptr(out),
                1 if out_flat else 0, n, H, H, C, eps, 1 if swish else 0, what=pre)
            return out

        def resblock(x, pre, cin, cout, H):
            """ResBlock._forward (openaimodel.py:268-288), use_scale_shift_norm=False"""
            eo = (C.c_void_p(emb_all.data_ptr() + 4 * self._emb_off[pre][0]), self._emb_total)   # this block's emb_layers columns
            a1 = groupnorm(x, pre + ".in_layers.0", cin, H, 1e-5, True)
            h1 = conv(a1, False, pre + ".in_layers.2.weight", cin, cout, 3, H, rowbias=eo)
            a2 = groupnorm(h1, pre + ".out_layers.0", cout, H, 1e-5, True)
            skip = x if cin == cout else conv(x, False, pre + ".skip_connection.weight", cin, cout, 1, H)
            return conv(a2, False, pre + ".out_layers.3.weight", cout, cout, 3, H, addend=skip)

        def attention(xq, rows_q, Tq, kv, Tk, pre, kd, C, d, addend):
            """CrossAttention.forward (attention.py:168-192): projections, per-head softmax(QK^T/sqrt(d)) V, to_out + residual"""
            kind, stack, per = self._stack[pre]
            if kind == "qkv" and kv is xq:      # self-attention: one GEMM for q | k | v
                qkv = linear_rows(xq, rows_q, pre + ".to_q|k|v", C, 3 * C, bias=False, wop=stack)
                qp, kp_, vp, ldq, ldkv = _ptr(qkv), col_slice(qkv, C), col_slice(qkv, 2 * C), 3 * C, 3 * C
            else:                               # cross-attention: q from the tokens, k | v from the context in one GEMM
                q = linear_rows(xq, rows_q, pre + ".to_q.weight", C, C, bias=False)
                kvb = linear_rows(kv, n * Tk, pre + ".to_k|v", kd, 2 * C, bias=False, wop=stack if kind == "kv" else stack[per:])
                qp, kp_, vp, ldq, ldkv = _ptr(q), _ptr(kvb), col_slice(kvb, C), C, 2 * C
            o = buf(self._flat(rows_q, C))
            nbytes = int(L.salun_sd_attention_ws_bytes(n, Tq, Tk, heads, d))
            if (Tq, Tk, d) not in attn_ws:
                attn_ws[(Tq, Tk, d)] = buf(torch.zeros(nbytes, dtype=torch.uint8, device=dev))
            ws = attn_ws[(Tq, Tk, d)]
            run(L.salun_sd_attention_ld, h, _ptr(ws), nbytes, qp, ldq, kp_, ldkv, vp, ldkv, _ptr(o), n, Tq, Tk, heads, d, what=pre)
            return linear_rows(o, rows_q, pre + ".to_out.0.weight", C, C, addend=addend)

        def transformer(x, pre, C, d, H):
            """SpatialTransformer.forward + BasicTransformerBlock._forward (attention.py:234-303)"""
            T, rows = H * H, n * H * H
            xn = groupnorm(x, pre + ".norm", C, H, 1e-6, False, out_flat=True)
            tok = conv(xn, True, pre + ".proj_in.weight", C, C, 1, H, out_pad=False)
            for i in range(cfg.get("transformer_depth", 1)):
                b = f"{pre}.transformer_blocks.{i}"

                def ln(src, k):
                    out = buf(self._flat(rows, C))
                    run(L.salun_sd_layernorm, h, _ptr(src), _ptr(P[f"{b}.norm{k}.weight"]), _ptr(P[f"{b}.norm{k}.bias"]), _ptr(out),
                        rows, C, 1e-5, what=f"{b}.norm{k}")
                    return out

                n1 = ln(tok, 1)
                tok = attention(n1, rows, T, n1, T, b + ".attn1", C, C, d, tok)
                tok = attention(ln(tok, 2), rows, T, ctx_act, Lc, b + ".attn2", dc, C, d, tok)
                pr = linear_rows(ln(tok, 3), rows, b + ".ff.net.0.proj.weight", C, 8 * C)
                gg = buf(self._flat(rows, 4 * C))
                run(L.salun_sd_geglu, h, _ptr(pr), _ptr(gg), rows, 4 * C, what=b + ".ff.geglu")
                tok = linear_rows(gg, rows, b + ".ff.net.2.weight", 4 * C, C, addend=tok)
            return conv(tok, True, pre + ".proj_out.weight", C, C, 1, H, addend=x)

        def run_layers(x, layers, H):
            for kind, pre, a, b in layers:
                if kind == "res":
                    x = resblock(x, pre, a, b, H)
                elif kind == "attn":
                    x = transformer(x, pre, a, b, H)
                elif kind == "down":       # Downsample: conv 3x3 stride 2 padding 1
                    col = buf(self._flat(n * (H // 2) ** 2, 9 * a))
                    out = buf(self._padded(n, H // 2, b))
                    run(L.salun_op_conv_s2, h, _ptr(x), _ptr(col), _ptr(W[pre + ".weight"]), _ptr(P[pre + ".bias"]), _ptr(out), n, H, H,
                        a, b, what=pre)
                    x, H = out, H // 2
                elif kind == "up":         # Upsample: nearest x2 then conv 3x3
                    up = buf(self._padded(n, 2 * H, a))
                    run(L.salun_op_upsample2, h, _ptr(x), _ptr(up), n, H, a, what=pre + ".nearest")
                    H *= 2
                    x = conv(up, False, pre + ".weight", a, b, 3, H)
            return x, H

        cin_p = _pad64(cfg["in_channels"])
        xp = buf(self._padded(n, S, cin_p))
        run(L.salun_op_nchw_to_padded, h, _ptr(x_in), _ptr(xp), n, cfg["in_channels"], cin_p, S, S, what="input")
        hcur = conv(xp, False, "input_blocks.0.0.weight", cfg["in_channels"], mc, 3, S)
        H, hs = S, [(hcur, mc, S)]
        for layers in self.arch.input_blocks[1:]:
            hcur, H = run_layers(hcur, layers, H)
            hs.append((hcur, out_ch(layers[-1]), H))
        hcur, H = run_layers(hcur, self.arch.middle, H)
        ch = self.arch.middle[-1][3]
        for layers in self.arch.output_blocks:
            skip, cs, Hs = hs.pop()
            assert Hs == H
            cat = buf(self._padded(n, H, ch + cs))
            run(L.salun_op_concat, h, _ptr(hcur), ch, _ptr(skip), cs, _ptr(cat), n, H, what="concat")   # th.cat([h, hs.pop()], dim=1)
            hcur, H = run_layers(cat, layers, H)
            ch = layers[0][3]
        a = groupnorm(hcur, "out.0", ch, H, 1e-5, True)
        co = cfg["out_channels"]
        yf = buf(torch.zeros(n * S * S, _pad64(co), device=dev))
        conv(a, False, "out.2.weight", ch, co, 3, S, bias=False, out_f32=yf)
        run(L.salun_op_rows_to_nchw, h, _ptr(yf), _pad64(co), _ptr(P["out.2.bias"]), _ptr(eps_out), n, co, S, S, what="out.2 -> NCHW")
        return dict(ops=ops, keep=keep, x=x_in, t=t_in, c=c_in, eps=eps_out, graph=None)

    def _run_ops(self, prog):
        s = _stream(self.device)
        for fn, a, what in prog["ops"]:
            check(fn(*a, s), what, self._lib)

    def _program(self, n: int) -> dict:
        prog = self._programs.get(n)
        if prog is None:
            if not self.params:
                raise RuntimeError("load_state_dict first")
            prog = self._build(n)
            self._programs[n] = prog
            if self.use_graph:
                cur = torch.cuda.current_stream(self.device)
                side = torch.cuda.Stream(self.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):       # warm-up outside the capture (function attributes, lazy module loads)
                    self._run_ops(prog)
                cur.wait_stream(side)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._run_ops(prog)
                prog["graph"] = g
        return prog

    # ---- the reference's call: model.apply_model(x, t, c) -> diffusion_model(x, timesteps=t, context=c) -----------------
    @torch.no_grad()
    def forward(self, x: torch.Tensor, timesteps: torch.Tensor, context: torch.Tensor) -> torch.Tensor:
        n = int(x.shape[0])
        if not 0 < n <= self.max_batch:
            raise ValueError(f"batch {n} out of range (max_batch {self.max_batch})")
        if tuple(x.shape[1:]) != (self.cfg["in_channels"], self.S, self.S):
            raise ValueError(f"x must be [n,{self.cfg['in_channels']},{self.S},{self.S}]")
        if tuple(context.shape) != (n, self.L, self.cfg["context_dim"]):
            raise ValueError(f"context must be [n,{self.L},{self.cfg['context_dim']}]")
        prog = self._program(n)
        prog["x"].copy_(x, non_blocking=True)
        prog["t"].copy_(timesteps.to(torch.float32), non_blocking=True)
        prog["c"].copy_(context, non_blocking=True)
        if prog["graph"] is not None:
            prog["graph"].replay()
        else:
            self._run_ops(prog)
        return prog["eps"].clone()

    __call__ = forward

    def launches_per_forward(self, n: int = 1) -> int:
        return len(self._program(n)["ops"])
