"""sm_100a DDPM U-Net engine (salun_unet_*, through the C ABI) against
  * the golden outputs of the UNMODIFIED reference Conditional_Model (tests/golden/ddpm_tiny.npz), and
  * the PyTorch fp32 restatement (oracle/unet.py, pinned to the reference on CPU by
    tests/test_ddpm_cpu.py) run on the same GPU with TF32 off,
and the DDPM SalUn loop bodies on the engine against the reference's statements (DDPM/runners/diffusion.py:519-593,
959-1039) with stock PyTorch.

Tolerance model: the engine stores activations / GEMM operands in bf16 (rel. rounding 2^-9 = 0.2% per tensor) with fp32
accumulation; through the ~60 layers of the network the relative L2 error grows smoothly to ~1% on eps and ~2-4% on the
deepest gradients (profiles/r1_unet_probe_full_n8.log).  Tests bound the relative L2 error per tensor, not elementwise.
"""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import tail as OT
from tests.golden.make_golden_ddpm import inputs, synth_weights, tiny_config

pytestmark = pytest.mark.gpu

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

G = os.path.join(os.path.dirname(__file__), "golden", "ddpm_tiny.npz")


def small_config(dropout=0.0):
    """two levels (16x16 -> 8x8), channel change 128 -> 256 (nin_shortcut), 384-wide skip concat, attention with 64 tokens"""
    return SimpleNamespace(
        model=SimpleNamespace(type="conditional", in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2], num_res_blocks=1,
                              attn_resolutions=[8], dropout=dropout, resamp_with_conv=True, cond_drop_prob=0.1),
        data=SimpleNamespace(image_size=16, channels=3, n_classes=10),
        diffusion=SimpleNamespace(beta_schedule="linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000))


def attn16_config():
    """attention with 256 tokens (16x16), the shape of the cifar10 config's attention blocks"""
    c = small_config()
    c.model.attn_resolutions = [16]
    return c


def rel(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm() / (b.norm() + 1e-20))


def _torch_model(cfg, seed=0):
    from oracle.unet import ConditionalUNet
    torch.manual_seed(seed)
    m = ConditionalUNet(cfg).cuda()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for k, p in m.named_parameters():
            if "norm" in k and k.endswith("weight"):
                p.copy_((1.0 + 0.1 * torch.randn(p.shape, generator=g)).cuda())
            elif k.endswith("bias"):
                p.copy_((0.05 * torch.randn(p.shape, generator=g)).cuda())
    return m


def _engine(cfg, model, salun_ctx, max_batch=16):
    from unlearn_saliency_b200.diffusion.engine import UNetEngine
    eng = UNetEngine(cfg, max_batch=max_batch, ctx=salun_ctx)
    eng.load_state_dict(model.state_dict())
    return eng


def _batch(cfg, n, seed):
    g = torch.Generator().manual_seed(seed)
    S = cfg.data.image_size
    return (torch.randn(n, 3, S, S, generator=g).cuda(), torch.randint(0, 1000, (n,), generator=g).cuda(),
            torch.randint(0, 10, (n,), generator=g).cuda(), (torch.rand(n, generator=g) < 0.3).cuda(),
            (torch.randn(n, 3, S, S, generator=g) / n).cuda())


def _assert_grads_close(gd, gref, whole_tol=0.04, per_tol=0.10):
    tot = torch.cat([v.reshape(-1) for v in gd.values()])
    totr = torch.cat([gref[k].reshape(-1) for k in gd])
    assert torch.isfinite(tot).all()
    assert rel(tot, totr) < whole_tol, rel(tot, totr)
    # per tensor: error norm within per_tol of the tensor's own norm, plus a floor for gradients that are zero
    # analytically (softmax is invariant to the key bias: d loss / d k.bias == 0 up to rounding)
    floor = 1e-4 * max(float(v.norm()) for v in gref.values())
    for k in gd:
        err = float((gd[k].float() - gref[k].float()).norm())
        assert err <= per_tol * float(gref[k].norm()) + floor, (k, err, float(gref[k].norm()))


@pytest.mark.parametrize("which,n", [("tiny", 6), ("small", 8), ("attn16", 5)])
def test_engine_forward_backward_matches_torch(salun_ctx, which, n):
    cfg = {"tiny": tiny_config, "small": small_config, "attn16": attn16_config}[which]()
    model = _torch_model(cfg)
    model.eval()
    eng = _engine(cfg, model, salun_ctx).eval()
    x, t, c, drop, d_eps = _batch(cfg, n, 5)
    model.zero_grad()
    eps_ref = model(x, t.float(), c, mode="train", drop_mask=drop)
    (eps_ref * d_eps).sum().backward()
    gref = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in model.named_parameters()}
    eps = eng.forward(x, t.float(), c, drop=drop, save=True)
    assert rel(eps, eps_ref) < 0.02, rel(eps, eps_ref)
    eng.backward(d_eps)
    _assert_grads_close(eng.grad_dict(), gref)
    # a partial last batch (n not a multiple of the 8-image TMA boxes) and a second call reuse the cached plan
    eps2 = eng.forward(x[: n - 1].contiguous(), t[: n - 1].float().contiguous(), c[: n - 1].contiguous(),
                       drop=drop[: n - 1].contiguous(), save=False)
    assert rel(eps2, eps_ref[: n - 1]) < 0.02
    with pytest.raises(RuntimeError):
        eng.backward(d_eps)  # the last forward did not save its activations
    eng.close()


def test_engine_matches_reference_golden(salun_ctx):
    """outputs of the unmodified reference model (make_golden_ddpm.py) on its seeded weights / inputs"""
    from unlearn_saliency_b200.diffusion.engine import UNetEngine
    from unlearn_saliency_b200.diffusion.runner import q_sample
    from oracle.unet import ConditionalUNet
    z = np.load(G)
    cfg = tiny_config()
    eng = UNetEngine(cfg, max_batch=16, ctx=salun_ctx).eval()
    eng.load_state_dict(synth_weights(ConditionalUNet(cfg)))
    x0, e, t, c = inputs()
    betas = torch.from_numpy(z["betas"])
    xt = q_sample(x0, t, e, betas).cuda().contiguous()
    tf, cc, n = t.float().cuda(), c.cuda(), x0.shape[0]
    zeros, ones = torch.zeros(n, dtype=torch.uint8, device="cuda"), torch.ones(n, dtype=torch.uint8, device="cuda")
    eps_train = eng.forward(xt, tf, cc, drop=zeros, save=True)
    ref_train = torch.from_numpy(z["eps_train"]).cuda()
    assert rel(eps_train, ref_train) < 0.03, rel(eps_train, ref_train)
    # loss = sum_CHW (e - eps)^2, mean over the batch (functions/losses.py:21-37), and its gradient
    ed = e.cuda()
    loss = (ed - eps_train).square().sum(dim=(1, 2, 3)).mean()
    assert abs(float(loss) - float(z["loss"])) < 0.03 * float(z["loss"])
    eng.backward(((-2.0 / n) * (ed - eps_train)).contiguous())
    gn = np.array([float(g.norm()) for g in eng.grad_dict().values()])
    ref_gn = z["gnorm"]
    big = ref_gn > 1e-3 * ref_gn.max()
    assert np.all(np.abs(gn[big] - ref_gn[big]) <= 0.08 * ref_gn[big]), np.abs(gn[big] / ref_gn[big] - 1).max()
    # classifier-free guidance: (1 + s) * eps(c) - s * eps(null)  (models/diffusion.py:340-355), as one batch of 2n
    eps2 = eng.forward(torch.cat([xt, xt]), torch.cat([tf, tf]), torch.cat([cc, cc]), drop=torch.cat([zeros, ones]))
    eps_test = 3.0 * eps2[:n] - 2.0 * eps2[n:]
    ref_test = torch.from_numpy(z["eps_test"]).cuda()
    assert rel(eps_test, ref_test) < 0.05, rel(eps_test, ref_test)
    eng.close()


def test_engine_accumulate_and_layout_round_trip(salun_ctx):
    cfg = tiny_config()
    model = _torch_model(cfg)
    eng = _engine(cfg, model, salun_ctx).eval()
    sd = eng.state_dict()
    for k, v in model.state_dict().items():
        assert torch.equal(sd[k], v), k   # OIHW -> OHWI arena -> OIHW is exact
    x, t, c, drop, d_eps = _batch(cfg, 4, 9)
    eng.forward(x, t.float(), c, drop=drop, save=True)
    eng.backward(d_eps)
    g1 = eng.grads.clone()
    eng.backward(d_eps, accumulate=True)
    torch.testing.assert_close(eng.grads, 2 * g1, rtol=1e-6, atol=1e-9)
    eng.backward(d_eps)                      # run-to-run identical (no atomics anywhere)
    assert torch.equal(eng.grads, g1)
    eng.close()


class _FixedDropout(torch.nn.Module):
    def __init__(self, scale):
        super().__init__()
        self.scale = scale

    def forward(self, x):
        return x * self.scale


def test_engine_dropout_forward_backward_consistent(salun_ctx):
    """train-mode dropout uses the engine's own counter-based generator: recover its masks from the exported
    post-dropout activations, plant them in the PyTorch restatement, and compare outputs and gradients."""
    p = 0.25
    cfg = tiny_config()
    cfg.model.dropout = p
    model = _torch_model(cfg)
    model.eval()
    eng = _engine(cfg, model, salun_ctx)
    x, t, c, drop, d_eps = _batch(cfg, 6, 11)
    n = x.shape[0]
    e_eval = eng.forward(x, t.float(), c, drop=drop, train=False)
    e_tr = eng.forward(x, t.float(), c, drop=drop, train=True, seed=7, save=True)
    e_tr_b = eng.forward(x, t.float(), c, drop=drop, train=True, seed=7)
    e_tr_c = eng.forward(x, t.float(), c, drop=drop, train=True, seed=8)
    assert torch.equal(e_tr, e_tr_b) and not torch.equal(e_tr, e_tr_c) and not torch.equal(e_tr, e_eval)
    eng.forward(x, t.float(), c, drop=drop, train=True, seed=7, save=True)
    blocks = {k[:-3]: None for k in eng.tensor_names() if k.endswith(".a2")}
    kept = tot = 0
    for prefix in blocks:
        a2 = eng.export(prefix + ".a2", n)
        mask = (a2 != 0).float()
        kept += float(mask.sum())
        tot += mask.numel()
        blk = model
        for part in prefix.split("."):
            blk = blk[int(part)] if part.isdigit() else getattr(blk, part)
        blk.dropout = _FixedDropout(mask / (1 - p))
    assert abs(kept / tot - (1 - p)) < 0.01, kept / tot
    model.zero_grad()
    eps_ref = model(x, t.float(), c, mode="train", drop_mask=drop)
    (eps_ref * d_eps).sum().backward()
    gref = {k: (q.grad if q.grad is not None else torch.zeros_like(q)) for k, q in model.named_parameters()}
    assert rel(e_tr, eps_ref) < 0.02, rel(e_tr, eps_ref)
    eng.backward(d_eps)
    _assert_grads_close(eng.grad_dict(), gref)
    eng.close()


# ---- loop bodies on the engine vs the reference's statements with stock PyTorch -----------------------------------

def _draw(seed, n, size):
    g = torch.Generator().manual_seed(seed)
    return dict(x_r=torch.rand(n, 3, size, size, generator=g), c_r=torch.randint(1, 10, (n,), generator=g),
                x_f=torch.rand(n, 3, size, size, generator=g), c_f=torch.zeros(n, dtype=torch.long),
                t_r=torch.randint(0, 1000, (n,), generator=g), e_r=torch.randn(n, 3, size, size, generator=g),
                t_f=torch.randint(0, 1000, (n,), generator=g), e_f=torch.randn(n, 3, size, size, generator=g),
                drop_r=torch.rand(n, generator=g) < 0.1, drop_f=torch.rand(n, generator=g) < 0.1,
                drop_p=torch.rand(n, generator=g) < 0.1)


@pytest.mark.parametrize("method", ["rl", "ga"])
def test_engine_saliency_unlearn_step_matches_reference_statements(salun_ctx, method):
    """DDPMEngineUnlearner.saliency_unlearn_step vs oracle/ddpm.py (the reference's statements, stock PyTorch fp32)"""
    from oracle import ddpm as OD
    from unlearn_saliency_b200.diffusion.runner import DDPMEngineUnlearner, get_beta_schedule
    cfg = small_config()
    ref = _torch_model(cfg)
    eng = _engine(cfg, ref, salun_ctx)
    betas = torch.from_numpy(get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)).float()
    g = torch.Generator().manual_seed(3)
    mask = {k: (torch.rand(p.shape, generator=g) < 0.5).to(torch.int64) for k, p in ref.named_parameters()}
    un = DDPMEngineUnlearner(eng, betas, lr=1e-4, grad_clip=1.0, mask={"module." + k: v for k, v in mask.items()})
    opt = torch.optim.Adam(ref.parameters(), lr=1e-4, weight_decay=0.0, betas=(0.9, 0.999), amsgrad=False, eps=1e-8)
    p0 = {k: p.detach().clone() for k, p in ref.named_parameters()}
    r = _draw(10, 4, cfg.data.image_size)
    loss_mine = un.saliency_unlearn_step(r["x_r"], r["c_r"], r["x_f"], r["c_f"], alpha=1e-3, method=method,
                                         rng={k: r[k] for k in ("t_r", "e_r", "t_f", "e_f", "drop_r", "drop_f", "drop_p")})
    loss, norm_ref, gref = OD.saliency_unlearn_step(ref, opt, mask, r, betas.cuda(), alpha=1e-3, method=method)
    assert abs(float(loss_mine) - float(loss)) <= 0.02 * abs(float(loss)) + 1e-5, (float(loss_mine), float(loss))
    assert abs(float(un.opt.grad_norm()) - float(norm_ref)) <= 0.03 * float(norm_ref)   # the clip used the same norm
    _assert_grads_close(eng.grad_dict(), gref)
    # first Adam step = lr * sign(g) on masked-in coordinates (|g| >> eps); masked-out coordinates must not move at all
    mine = eng.state_dict()
    agree = tot = 0
    for k, p in ref.named_parameters():
        m = mask[k].cuda().bool()
        assert torch.equal(mine[k][~m], p0[k][~m]), k
        big = m & (gref[k].abs() > 0.1 * gref[k].abs().mean())     # skip near-zero gradients (their sign is rounding noise)
        agree += int((torch.sign(mine[k] - p0[k])[big] == torch.sign(p.detach() - p0[k])[big]).sum())
        tot += int(big.sum())
    assert agree / tot > 0.97, agree / tot
    eng.close()


def test_engine_generate_mask_matches_reference_statements(salun_ctx, tmp_path):
    """DDPMEngineUnlearner.generate_mask_batch / finish_mask vs oracle/ddpm.py + the argsort definition of the mask"""
    from oracle import ddpm as OD
    from unlearn_saliency_b200.diffusion.runner import DDPMEngineUnlearner, get_beta_schedule
    cfg = small_config()
    ref = _torch_model(cfg)
    eng = _engine(cfg, ref, salun_ctx, max_batch=8)
    betas = torch.from_numpy(get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)).float()
    un = DDPMEngineUnlearner(eng, betas)
    grads = {}
    for b, n in enumerate((4, 6)):   # 2n = 8 fits max_batch (one batched call), 2n = 12 does not (three-pass path)
        r = _draw(30 + b, n, cfg.data.image_size)
        loss_mine = un.generate_mask_batch(r["x_f"], r["c_f"], cond_scale=2.0, t=r["t_f"], e=r["e_f"])
        loss = OD.generate_mask_batch(ref, r["x_f"].cuda(), r["c_f"].cuda(), r["t_f"].cuda(), r["e_f"].cuda(), betas.cuda(),
                                      grads, cond_scale=2.0)
        assert abs(float(loss_mine) - float(loss)) <= 0.03 * float(loss)
    grads = {k: grads[k] for k, _ in ref.named_parameters()}
    path = str(tmp_path / "mask" / "0" / "with_0.5.pt")
    un.finish_mask(path, 0.5)
    m = torch.load(path)
    assert list(m.keys()) == ["module." + k for k in grads]
    assert all(v.dtype == torch.int64 and v.device.type == "cpu" and v.shape == ref.state_dict()[k[7:]].shape
               for k, v in m.items())
    # accumulated saliency: relative L2 error of the whole vector, then the masks
    acc = eng.from_native(un.saliency.acc)
    flat_ref = torch.cat([torch.as_tensor(g).flatten() for g in grads.values()])
    flat_mine = torch.cat([acc[k].flatten().cpu() for k in grads])
    assert rel(flat_mine, flat_ref) < 0.04, rel(flat_mine, flat_ref)
    mine_mask = torch.cat([v.flatten() for v in m.values()]).numpy()
    k = int(mine_mask.size * 0.5)
    assert mine_mask.sum() == k
    # bit-exact selection on the engine's own accumulator (the index-set claim is on identical inputs) ...
    own = OT.topk_mask_argsort(flat_mine.abs().numpy(), k)
    assert np.array_equal(mine_mask, own)
    # ... and against the fp32 reference statements only elements within the bf16 error band of the threshold differ
    ref_mask = OT.topk_mask_argsort(flat_ref.abs().numpy(), k)
    assert (mine_mask != ref_mask).mean() < 0.05, (mine_mask != ref_mask).mean()
    eng.close()


def test_diffusion_runner_mirror_end_to_end(salun_ctx, tmp_path, monkeypatch):
    """Diffusion(args, config).generate_mask() / .saliency_unlearn() (DDPM/train.py:150-155) on synthetic loaders:
    reads <ckpt_folder>/ckpts/ckpt.pth, writes results/cifar10/mask/<label>/with_0.5.pt and <ckpt_dir>/ckpt.pth in the
    reference's formats; masked-out weights stay at their checkpoint values."""
    from torch.utils.data import DataLoader, TensorDataset
    from unlearn_saliency_b200.diffusion.runner import Diffusion
    from oracle.unet import ConditionalUNet
    cfg = tiny_config()
    cfg.training = SimpleNamespace(batch_size=4, n_iters=3, snapshot_freq=3, log_freq=1)
    cfg.optim = SimpleNamespace(weight_decay=0.0, optimizer="Adam", lr=1e-4, beta1=0.9, amsgrad=False, eps=1e-8, grad_clip=1.0)
    cfg.ckpt_dir = str(tmp_path / "out")
    model = ConditionalUNet(cfg)
    sd0 = {"module." + k: v.clone() for k, v in model.state_dict().items()}
    (tmp_path / "ck" / "ckpts").mkdir(parents=True)
    torch.save([sd0, {}, 0], str(tmp_path / "ck" / "ckpts" / "ckpt.pth"))
    g = torch.Generator().manual_seed(0)
    S = cfg.data.image_size
    remain = DataLoader(TensorDataset(torch.rand(12, 3, S, S, generator=g), torch.randint(1, 10, (12,), generator=g)), batch_size=4)
    forget = DataLoader(TensorDataset(torch.rand(8, 3, S, S, generator=g), torch.zeros(8, dtype=torch.long)), batch_size=4)
    monkeypatch.chdir(tmp_path)
    args = SimpleNamespace(ckpt_folder=str(tmp_path / "ck"), label_to_forget=0, cond_scale=2.0, mask_path=None, alpha=1e-3,
                           method="rl", seed=1234)
    runner = Diffusion(args, cfg, loaders=(remain, forget))
    runner.generate_mask()
    mpath = tmp_path / "results" / "cifar10" / "mask" / "0" / "with_0.5.pt"
    mask = torch.load(str(mpath))
    assert list(mask.keys()) == list(sd0.keys())
    n_tot = sum(v.numel() for v in mask.values())
    assert sum(int(v.sum()) for v in mask.values()) == int(n_tot * 0.5)
    assert all(v.dtype == torch.int64 and v.device.type == "cpu" for v in mask.values())
    args.mask_path = str(mpath)
    loss = runner.saliency_unlearn()
    assert loss is not None and np.isfinite(loss)
    states = torch.load(str(tmp_path / "out" / "ckpt.pth"))
    assert len(states) == 3 and states[2] == 2 and list(states[0].keys()) == list(sd0.keys())
    moved = 0
    for k, v in states[0].items():
        m = mask[k].bool()
        assert torch.equal(v.cpu()[~m], sd0[k][~m]), k
        moved += int((v.cpu()[m] != sd0[k][m]).sum())
    assert moved > 0.9 * int(n_tot * 0.5)


def mid_config():
    """three levels (32x32 -> 16x16 -> 8x8), two ResnetBlocks per level, attention with 256 tokens at 16x16: every kernel
    variant of the cifar10 config (1024-pixel images, stride-2 down / nearest up at two scales, 384-wide concat at 32x32)"""
    c = small_config()
    c.model.ch_mult, c.model.attn_resolutions, c.model.num_res_blocks = [1, 2, 2], [16], 2
    c.data.image_size = 32
    return c


@pytest.mark.parametrize("n", [1, 3, 16])
def test_engine_batch_edges(salun_ctx, n):
    """one image, an odd batch and the full allocated batch through the same engine instance's plan cache"""
    cfg = mid_config()
    model = _torch_model(cfg)
    model.eval()
    eng = _engine(cfg, model, salun_ctx, max_batch=16).eval()
    for nn_ in (n, max(1, n - 1), n):     # growing / shrinking batches must not see stale rows of earlier ones
        x, t, c, drop, d_eps = _batch(cfg, nn_, 20 + nn_)
        model.zero_grad()
        eps_ref = model(x, t.float(), c, mode="train", drop_mask=drop)
        (eps_ref * d_eps).sum().backward()
        gref = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in model.named_parameters()}
        eps = eng.forward(x, t.float(), c, drop=drop, save=True)
        assert rel(eps, eps_ref) < 0.02, (nn_, rel(eps, eps_ref))
        eng.backward(d_eps)
        _assert_grads_close(eng.grad_dict(), gref, whole_tol=0.05, per_tol=0.12)
    eng.close()


def test_engine_rejects_unserved_configs(salun_ctx):
    from unlearn_saliency_b200.diffusion.engine import UNetEngine
    for mutate in (lambda c: setattr(c.model, "ch", 64), lambda c: setattr(c.model, "attn_resolutions", [32]),
                   lambda c: setattr(c.data, "image_size", 24)):
        cfg = mid_config()
        mutate(cfg)
        with pytest.raises(RuntimeError):
            UNetEngine(cfg, max_batch=4, ctx=salun_ctx)
    eng = UNetEngine(tiny_config(), max_batch=4, ctx=salun_ctx)
    x = torch.zeros(5, 3, 8, 8, device="cuda")
    with pytest.raises(ValueError):
        eng.forward(x, torch.zeros(5, device="cuda"), torch.zeros(5, dtype=torch.long, device="cuda"))   # batch > max_batch
    with pytest.raises(ValueError):
        eng.forward(x[:2].double(), torch.zeros(2, device="cuda"), torch.zeros(2, dtype=torch.long, device="cuda"))
    eng.close()


def test_cli_mirror_of_train_py(salun_ctx, tmp_path, monkeypatch):
    """python -m unlearn_saliency_b200.diffusion.cli: the flags / YAML keys / directories of DDPM/train.py on synthetic data"""
    import yaml
    from unlearn_saliency_b200.diffusion import cli
    from oracle.unet import ConditionalUNet
    monkeypatch.chdir(tmp_path)
    cfg = dict(data=dict(dataset="CIFAR10", image_size=8, channels=3, random_flip=True, num_workers=0, n_classes=10, path="./data"),
               model=dict(type="simple", in_channels=3, out_ch=3, ch=128, ch_mult=[1, 1], num_res_blocks=1, attn_resolutions=[4],
                          dropout=0.1, ema=False, resamp_with_conv=True, cond_drop_prob=0.1),
               diffusion=dict(beta_schedule="linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000),
               training=dict(batch_size=4, n_iters=2, snapshot_freq=2, log_freq=1),
               optim=dict(weight_decay=0.0, optimizer="Adam", lr=1e-4, beta1=0.9, amsgrad=False, eps=1e-8, grad_clip=1.0))
    (tmp_path / "configs").mkdir()
    (tmp_path / "configs" / "tiny.yml").write_text(yaml.safe_dump(cfg))
    model = ConditionalUNet(cli.dict2namespace(cfg))
    (tmp_path / "ck" / "ckpts").mkdir(parents=True)
    torch.save([{"module." + k: v for k, v in model.state_dict().items()}, {}, 0], str(tmp_path / "ck" / "ckpts" / "ckpt.pth"))
    assert cli.main(["--config", "tiny.yml", "--ckpt_folder", "ck", "--label_to_forget", "3", "--mode", "generate_mask",
                     "--synthetic", "8"]) == 0
    mpath = tmp_path / "results" / "cifar10" / "mask" / "3" / "with_0.5.pt"
    assert mpath.exists()
    assert cli.main(["--config", "tiny.yml", "--ckpt_folder", "ck", "--label_to_forget", "3", "--mode", "saliency_unlearn",
                     "--mask_path", str(mpath), "--alpha", "0.001", "--method", "rl", "--synthetic", "8"]) == 0
    runs = list((tmp_path / "results" / "cifar10" / "forget" / "rl").glob("0.001_full/*/ckpts/ckpt.pth"))
    assert len(runs) == 1
    states = torch.load(str(runs[0]))
    assert states[2] == 1 and all(k.startswith("module.") for k in states[0])


def test_q_sample_and_loss_kernels_match_torch(salun_ctx):
    """salun_ddpm_q_sample is bit-identical to the reference statements (2x - 1, x0 sqrt(abar) + e sqrt(1 - abar));
    salun_ddpm_eps_loss_grad equals autograd on forget_loss + alpha * remain_loss (runners/diffusion.py:533-580)."""
    from oracle import ddpm as OD
    from unlearn_saliency_b200.diffusion.engine import DDPMLoss
    from unlearn_saliency_b200.diffusion.runner import get_beta_schedule
    betas = torch.from_numpy(get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)).float()
    L = DDPMLoss(betas, salun_ctx)
    g = torch.Generator().manual_seed(2)
    for n in (1, 7, 128):
        x01 = torch.rand(n, 3, 32, 32, generator=g).cuda()
        e = torch.randn(n, 3, 32, 32, generator=g).cuda()
        t = torch.randint(0, 1000, (n,), generator=g).cuda()
        ref = OD.q_sample(2 * x01 - 1.0, t, e, betas.cuda())
        assert torch.equal(L.q_sample(x01, e, t, rescale=True), ref)
        assert torch.equal(L.q_sample(x01, e, t, rescale=False), OD.q_sample(x01, t, e, betas.cuda()))
    nr, nf, alpha = 5, 3, 1e-3
    for method in ("rl", "ga"):
        eps = torch.randn(nr + nf, 3, 16, 16, generator=g).cuda().requires_grad_(True)
        e = torch.randn(nr + nf, 3, 16, 16, generator=g).cuda()
        pseudo = torch.randn(nf, 3, 16, 16, generator=g).cuda()
        remain = (e[:nr] - eps[:nr]).square().sum(dim=(1, 2, 3)).mean(dim=0)
        if method == "rl":
            forget = torch.nn.functional.mse_loss(eps[nr:], pseudo)
            target = torch.cat([e[:nr], pseudo])
            wf = 1.0 / (nf * 3 * 16 * 16)
        else:
            forget = -(e[nr:] - eps[nr:]).square().sum(dim=(1, 2, 3)).mean(dim=0)
            target = e
            wf = -1.0 / nf
        loss_ref = forget + alpha * remain
        loss_ref.backward()
        w = torch.cat([torch.full((nr,), alpha / nr), torch.full((nf,), wf)]).cuda()
        loss, d, ss = L.loss_grad(eps.detach(), target, w)
        torch.testing.assert_close(loss[0], loss_ref.detach(), rtol=2e-5, atol=1e-7)
        torch.testing.assert_close(d, eps.grad, rtol=1e-5, atol=1e-9)
        torch.testing.assert_close(ss, (eps.detach() - target).square().sum(dim=(1, 2, 3)), rtol=1e-5, atol=1e-6)


def test_engine_matches_reference_golden_channel_changing_config(salun_ctx):
    """tests/golden/ddpm_small.npz -- outputs of the UNMODIFIED reference on the config with nin_shortcut, the 384-wide
    skip concat, down / up sampling and 64-token attention -- against the engine directly"""
    from tests.golden.make_golden_ddpm import default_init_weights, small_config as golden_small
    from unlearn_saliency_b200.diffusion.engine import DDPMLoss, UNetEngine
    from unlearn_saliency_b200.diffusion.runner import get_beta_schedule
    from oracle.unet import ConditionalUNet
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ddpm_small.npz"))
    cfg = golden_small()
    eng = UNetEngine(cfg, max_batch=8, ctx=salun_ctx).eval()
    assert eng.names == list(z["keys"])
    eng.load_state_dict(default_init_weights(ConditionalUNet(cfg)))
    x0, e, t, c = inputs(seed=2, n=4, size=16)
    betas = torch.from_numpy(get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)).float()
    L = DDPMLoss(betas, salun_ctx)
    xt = L.q_sample(x0.cuda(), e.cuda(), t.cuda(), rescale=False)
    n = 4
    tf, cc = t.float().cuda(), c.cuda()
    zeros, ones = torch.zeros(n, dtype=torch.uint8, device="cuda"), torch.ones(n, dtype=torch.uint8, device="cuda")
    eps2 = eng.forward(torch.cat([xt, xt]), torch.cat([tf, tf]), torch.cat([cc, cc]), drop=torch.cat([zeros, ones]))
    assert rel(eps2[:n], torch.from_numpy(z["eps_cond"]).cuda()) < 0.02
    assert rel(eps2[n:], torch.from_numpy(z["eps_null"]).cuda()) < 0.02
    eps = eng.forward(xt, tf, cc, drop=zeros, save=True)
    loss, d, _ = L.loss_grad(eps, e.cuda(), torch.full((n,), 1.0 / n, device="cuda"))
    assert abs(float(loss) - float(z["loss"])) < 0.02 * float(z["loss"])
    eng.backward(d)
    gn = np.array([float(g.norm()) for g in eng.grad_dict().values()])
    ref_gn = z["gnorm"]
    big = ref_gn > 1e-3 * ref_gn.max()
    assert np.all(np.abs(gn[big] - ref_gn[big]) <= 0.08 * ref_gn[big]), np.abs(gn[big] / ref_gn[big] - 1).max()
    eng.close()
