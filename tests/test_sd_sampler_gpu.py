"""Row f1 on the GPU: DDIM partial sampling (t_start / till_T / classifier-free guidance) on SDUNetEngine against the torch
restatement (oracle/sd_unet.py, pinned to the unmodified reference on CPU), the in-place parameter refresh ESD needs, and a
whole ESD iteration whose no-grad passes (quick_sample_till_t and the frozen e_0 / e_p) run on the engine."""
import copy
import zlib

import pytest
import torch
import torch.nn as nn

from oracle import sd_unet as OS
from tests.golden.make_golden_sd import CONFIGS, sd_synth_weights

pytestmark = pytest.mark.gpu
TOL = {"bf16": 6e-2, "split": 1e-3}


def _rel(a, r):
    return float((a.float() - r.float()).norm() / r.float().norm())


def _engine(ctx, precision, max_batch=2, seed=7):
    from unlearn_saliency_b200 import _lib
    from unlearn_saliency_b200.sd.engine import SDUNetEngine
    if precision not in _lib.available_precisions():
        pytest.skip("build missing")
    c = CONFIGS["a"]
    eng = SDUNetEngine(c["cfg"], latent_size=c["latent"], max_batch=max_batch, context_len=c["ctx_len"], ctx=ctx,
                       precision=precision)
    P = {k: v.cuda() for k, v in sd_synth_weights(eng.table, seed=seed).items()}
    eng.load_state_dict(P)
    return eng, P, c


@pytest.fixture(autouse=True)
def _fp32_reference():
    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("precision", ["split", "bf16"])
@pytest.mark.parametrize("eta,scale,till_T", [(0.0, 3.0, 4), (0.5, 7.5, None), (0.0, 1.0, 6)])
def test_ddim_partial_sampling(salun_ctx, precision, eta, scale, till_T):
    from unlearn_saliency_b200.sd.sampler import EngineDDIMSampler
    eng, P, c = _engine(salun_ctx, precision)
    g = torch.Generator().manual_seed(5)
    S, L, D = c["latent"], c["ctx_len"], c["cfg"]["context_dim"]
    x_T = torch.randn(1, 4, S, S, generator=g).cuda()
    cond, uncond = torch.randn(1, L, D, generator=g).cuda(), torch.randn(1, L, D, generator=g).cuda()
    noises = [torch.randn(1, 4, S, S, generator=g).cuda() for _ in range(10)]
    sampler = EngineDDIMSampler(eng)
    launches0 = salun_ctx.launch_count() if hasattr(salun_ctx, "launch_count") else None
    z, inter = sampler.sample(S=10, batch_size=1, shape=[4, S, S], conditioning=cond, eta=eta, x_T=x_T,
                              unconditional_guidance_scale=scale, unconditional_conditioning=uncond, till_T=till_T,
                              noise=noises)
    with torch.no_grad():
        ref = OS.ddim_sample(lambda x, t, cc: OS.unet_forward(P, c["cfg"], x, t.float(), cc), cond, uncond, x_T, 10, scale,
                             eta=eta, till_T=till_T, noises=noises)
    err = _rel(z, ref)
    print(f"DDIM partial sampling [{precision}] eta={eta} scale={scale} till_T={till_T}: relative error {err:.3e}")
    assert torch.isfinite(z).all() and err < TOL[precision], err
    assert not torch.equal(z, x_T) and len(inter["x_inter"]) >= 2


def test_schedule_tables_equal_restatement(salun_ctx):
    from unlearn_saliency_b200.sd import sampler as SP
    import numpy as np
    ac = np.cumprod(1.0 - SP.make_beta_schedule())
    np.testing.assert_array_equal(ac, OS.sd_alphas_cumprod())
    for n, eta in ((50, 0.0), (10, 0.7), (25, 1.0)):
        steps = SP.make_ddim_timesteps("uniform", n, 1000)
        sg, a, ap = SP.make_ddim_sampling_parameters(ac, steps, eta)
        rs, ra, rap, rsg = OS.ddim_schedule(ac, n, eta)
        for mine, ref in ((steps, rs), (a, ra), (ap, rap), (sg, rsg)):
            np.testing.assert_array_equal(mine, ref)


def test_update_parameters_in_place_keeps_graph(salun_ctx):
    eng, P, c = _engine(salun_ctx, "split")
    from tests.golden.make_golden_sd import sd_inputs
    x, t, ctx = (v.cuda() for v in sd_inputs("a"))
    e0 = eng(x, t, ctx)
    graph0 = eng._programs[2]["graph"]
    P2 = {k: v.cuda() for k, v in sd_synth_weights(eng.table, seed=8).items()}
    xattn = {k: P2[k] for k in P2 if "attn2" in k}
    eng.update_parameters(xattn)
    assert eng._programs[2]["graph"] is graph0            # the captured program survives the refresh
    e1 = eng(x, t, ctx)
    Pm = dict(P); Pm.update(xattn)
    with torch.no_grad():
        ref = OS.unet_forward(Pm, c["cfg"], x, t, ctx)
    assert _rel(e1, ref) < TOL["split"] and _rel(e0, ref) > 1e-2
    with pytest.raises(KeyError):
        eng.update_parameters({"nope.weight": torch.zeros(1)})


# ---- a whole ESD iteration ------------------------------------------------------------------------------------------------
class _TorchUNet(nn.Module):
    """UNetModel stand-in with the reference's parameter names (nested containers) and the restated forward"""

    def __init__(self, P, cfg):
        super().__init__()
        self.cfg = cfg
        self.num_heads = cfg["num_heads"]
        for k, v in P.items():
            m, parts = self, k.split(".")
            for p in parts[:-1]:
                if p not in m._modules:
                    m.add_module(p, nn.Module())
                m = m._modules[p]
            m.register_parameter(parts[-1], nn.Parameter(v.clone()))

    def forward(self, x, t, context):
        return OS.unet_forward(dict(self.named_parameters()), self.cfg, x, t.float(), context)


class _LDM(nn.Module):
    def __init__(self, P, cfg, L, D):
        super().__init__()
        self.model = nn.Module()
        self.model.diffusion_model = _TorchUNet(P, cfg)
        self.L, self.D = L, D

    def get_learned_conditioning(self, prompts):
        out = []
        for p in prompts:
            g = torch.Generator().manual_seed(zlib.crc32(p.encode()))
            out.append(torch.randn(self.L, self.D, generator=g))
        return torch.stack(out).cuda()

    def apply_model(self, x, t, cond):
        return self.model.diffusion_model(x, t, cond)


def test_esd_iteration_with_engine_sampler_and_frozen_engine(salun_ctx):
    from unlearn_saliency_b200.sd import SDTail, esd_iteration
    from unlearn_saliency_b200.sd.sampler import EngineApplyModel, EngineDDIMSampler, make_quick_sample_till_t
    eng, P, c = _engine(salun_ctx, "split")
    frozen_eng, _, _ = _engine(salun_ctx, "split")
    S, L, D = c["latent"], c["ctx_len"], c["cfg"]["context_dim"]
    mine = _LDM(P, c["cfg"], L, D).cuda()
    ref, frozen = copy.deepcopy(mine), copy.deepcopy(mine)
    lr, steps = 1e-4, 10
    tail = SDTail(mine, lr=lr, train_method="xattn", ctx=salun_ctx)
    sel = [n for n, _ in ref.model.diffusion_model.named_parameters() if "attn2" in n]
    opt = torch.optim.Adam([p for n, p in ref.model.diffusion_model.named_parameters() if n in sel], lr=lr)
    uncond = mine.get_learned_conditioning([""])
    sample_fn = make_quick_sample_till_t(EngineDDIMSampler(eng), uncond, image_size=8 * S, ddim_steps=steps,
                                         module=mine.model.diffusion_model, train_keys=sel)
    for it in range(2):
        g = torch.Generator().manual_seed(70 + it)
        rng = dict(t_enc=torch.randint(1, steps, (1,), generator=g).cuda(), start_code=torch.randn(1, 4, S, S, generator=g).cuda())
        rng["t_enc_ddpm"] = torch.randint(0, 1000, (1,), generator=g).cuda()
        loss_m = esd_iteration(mine, EngineApplyModel(frozen_eng), sample_fn, tail, "Van Gogh", 3.0, 1.0, image_size=8 * S,
                               ddim_steps=steps, rng=rng)
        emb_0, emb_p = ref.get_learned_conditioning([""]), ref.get_learned_conditioning(["Van Gogh"])
        opt.zero_grad()
        with torch.no_grad():
            z = OS.ddim_sample(ref.apply_model, emb_p, emb_0, rng["start_code"], steps, 3.0, till_T=int(rng["t_enc"]))
            e_0 = frozen.apply_model(z, rng["t_enc_ddpm"], emb_0)
            e_p = frozen.apply_model(z, rng["t_enc_ddpm"], emb_p)
        e_n = ref.apply_model(z, rng["t_enc_ddpm"], emb_p)
        loss = torch.nn.functional.mse_loss(e_n, e_0 - (1.0 * (e_p - e_0)))
        loss.backward()
        opt.step()
        print(f"ESD iteration {it}: loss engine-path {float(loss_m):.6f} | torch {float(loss):.6f}")
        assert abs(float(loss_m) - float(loss)) <= 2e-3 * abs(float(loss))
    num = den = 0.0
    for (n, p), (_, q) in zip(ref.model.diffusion_model.named_parameters(), mine.model.diffusion_model.named_parameters()):
        if n not in sel:
            assert torch.equal(q, P[n]), n
    # the second iteration sampled from the UPDATED weights on both sides: the engine refresh is on the path
    assert any(not torch.equal(q, P[n]) for n, q in mine.model.diffusion_model.named_parameters() if n in sel)


def test_train_esd_on_engines_matches_torch_sampling(salun_ctx):
    """train_esd(engine="split"): engines are built from the modules (config inferred from the parameter shapes), sampling
    follows the trained weights; losses equal a run whose no-grad passes are the torch restatement."""
    from unlearn_saliency_b200.sd import train_esd
    c = CONFIGS["a"]
    S, L, D = c["latent"], c["ctx_len"], c["cfg"]["context_dim"]
    P = {k: v.cuda() for k, v in sd_synth_weights(__import__("unlearn_saliency_b200.sd.engine", fromlist=["x"]).sd_unet_param_table(c["cfg"]), seed=7).items()}
    runs = []
    for engine in ("split", None):
        model, frozen = _LDM(P, c["cfg"], L, D).cuda(), _LDM(P, c["cfg"], L, D).cuda()
        uncond = model.get_learned_conditioning([""])
        torch_sample = lambda emb, s, code, t: OS.ddim_sample(model.apply_model, emb, uncond, code, 10, s, till_T=t)
        torch.manual_seed(123)
        with torch.random.fork_rng(devices=[0]):
            torch.manual_seed(123)
            runs.append(train_esd("Van Gogh", "xattn", 3.0, 1.0, 3, 1e-4, None, None, None, None, None, image_size=8 * S,
                                  ddim_steps=10, models=(frozen, None, model, None), ctx=salun_ctx, engine=engine,
                                  sample_fn=None if engine else torch_sample))
    print("train_esd losses: engines", runs[0], "| torch sampling", runs[1])
    assert all(abs(a - b) <= 3e-3 * abs(b) for a, b in zip(*runs)), runs
