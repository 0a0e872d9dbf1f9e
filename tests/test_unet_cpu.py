"""Host-side pieces of the DDPM / SD mirrors that need no GPU: the C++ architecture walk of the U-Net engine against the
reference's parameter count, the train.py-style CLI parsing / directory layout, and the train_method parameter selection
of the SD scripts."""
import ctypes as C
import os

import numpy as np
import yaml

from tests.golden.make_golden_ddpm import tiny_config
from unlearn_saliency_b200 import _lib
from unlearn_saliency_b200.diffusion.engine import salun_unet_cfg, unet_param_table
from unlearn_saliency_b200.diffusion.config import cifar10_config

G = os.path.join(os.path.dirname(__file__), "golden", "ddpm_tiny.npz")


def _count(cfg, max_batch=8):
    m, d = cfg.model, cfg.data
    mult, attn = list(m.ch_mult), list(m.attn_resolutions)
    c = salun_unet_cfg(m.ch, len(mult), (C.c_int * 8)(*mult), m.num_res_blocks, len(attn), (C.c_int * 8)(*attn),
                       d.image_size, m.in_channels, m.out_ch, d.n_classes, max_batch, 0.0)
    return int(_lib.lib().salun_unet_param_count(C.byref(c)))


def test_engine_parameter_count_equals_reference():
    """salun_unet_param_count walks the architecture in C++ (named_parameters() order); the total must be the reference
    model's (golden: 38 632 323 for the cifar10 config) and the host table's for other shapes"""
    z = np.load(G)
    assert _count(cifar10_config()) == int(z["numel_full"]) == 38632323
    for cfg in (tiny_config(), cifar10_config()):
        table = unet_param_table(cfg)
        assert _count(cfg) == sum(int(np.prod(s)) for s in table.values())
    assert list(unet_param_table(cifar10_config())) == list(z["keys_full"])


def test_engine_rejects_unserved_architectures_on_the_host():
    cfg = cifar10_config()
    cfg.model.ch = 64          # the reference's ResnetBlock hard-codes cemb_channels = 512 = 4 * 128
    assert _count(cfg) == -1
    cfg = cifar10_config()
    cfg.model.attn_resolutions = [32]
    assert _count(cfg) == -1
    cfg = cifar10_config()
    cfg.data.image_size = 48
    assert _count(cfg) == -1


def test_cli_parses_like_train_py(tmp_path, monkeypatch):
    from unlearn_saliency_b200.diffusion import cli
    monkeypatch.chdir(tmp_path)
    cfg = dict(data=dict(dataset="CIFAR10", image_size=32, n_classes=10), model=dict(ch=128, ema=False),
               training=dict(batch_size=128, n_iters=1000, snapshot_freq=100, log_freq=100),
               optim=dict(lr=1e-4, grad_clip=1.0))
    (tmp_path / "configs").mkdir()
    (tmp_path / "configs" / "c.yml").write_text(yaml.safe_dump(cfg))
    args, config = cli.parse_args_and_config(["--config", "c.yml", "--ckpt_folder", "ck", "--label_to_forget", "7",
                                              "--mode", "saliency_unlearn", "--mask_path", "results/cifar10/mask/7/with_0.5.pt",
                                              "--alpha", "0.001", "--method", "rl"])
    assert args.label_to_forget == 7 and args.cond_scale == 2.0 and args.seed == 1234 and args.method == "rl"
    assert config.training.n_iters == 1000 and config.optim.grad_clip == 1.0 and config.model.ch == 128
    # results/<dataset>/forget/<method>/<alpha>_<mask kind>/<timestamp>/{logs,ckpts}  (functions/__init__.py:51-87)
    parts = os.path.normpath(config.exp_root_dir).split(os.sep)
    assert parts[:5] == ["results", "cifar10", "forget", "rl", "0.001_full"]
    assert os.path.isdir(config.log_dir) and os.path.isdir(config.ckpt_dir)
    assert os.path.exists(os.path.join(config.log_dir, "config.yaml"))


def test_sd_train_method_selection():
    """train-esd.py:192-224 / random_label.py:45-54"""
    from unlearn_saliency_b200.sd import select_parameters
    names = ["time_embed.0.weight", "input_blocks.1.0.in_layers.2.weight",
             "input_blocks.4.1.transformer_blocks.0.attn1.to_q.weight", "input_blocks.4.1.transformer_blocks.0.attn2.to_k.weight",
             "input_blocks.7.1.transformer_blocks.0.attn1.to_v.weight", "output_blocks.6.1.transformer_blocks.0.attn2.to_v.weight",
             "output_blocks.8.1.transformer_blocks.0.attn2.to_out.0.weight", "output_blocks.9.1.transformer_blocks.0.attn1.to_q.weight",
             "out.2.weight"]
    sel = lambda m: select_parameters(names, m)
    assert sel("full") == names
    assert sel("xattn") == [n for n in names if "attn2" in n]
    assert sel("selfattn") == [n for n in names if "attn1" in n]
    assert sel("noxattn") == [n for n in names if not (n.startswith("out.") or "attn2" in n or "time_embed" in n)]
    assert sel("notime") == [n for n in names if not (n.startswith("out.") or "time_embed" in n)]
    assert sel("xlayer") == ["output_blocks.6.1.transformer_blocks.0.attn2.to_v.weight",
                             "output_blocks.8.1.transformer_blocks.0.attn2.to_out.0.weight"]
    assert sel("selflayer") == ["input_blocks.4.1.transformer_blocks.0.attn1.to_q.weight",
                                "input_blocks.7.1.transformer_blocks.0.attn1.to_v.weight"]
    assert sel("nonsense") == []


class _FakeSaliency:
    def __init__(self):
        self.reduced, self.saved = 0, []

    def all_reduce(self):
        self.reduced += 1

    def save(self, path, ratio, key_prefix=""):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        open(path, "wb").write(b"mask")
        self.saved.append((path, ratio, key_prefix))
        return "info"

    def mask(self, ratio, key_prefix=""):
        return {}, "info"


class _FakeUnlearner:
    """stands in for DDPMEngineUnlearner: records the control flow of the Diffusion mirror (no GPU)"""
    instances = []

    def __init__(self, eng, betas, **kw):
        self.kw, self.batches, self.steps, self.ckpts = kw, [], [], []
        self.saliency, self.fused_dp = _FakeSaliency(), False
        _FakeUnlearner.instances.append(self)

    def generate_mask_batch(self, x, c, cond_scale=2.0):
        self.batches.append(int(x[0, 0, 0, 0] * 1000))

    def saliency_unlearn_step(self, rx, rc, fx, fc, alpha, method, n_classes, global_counts=None):
        import torch
        self.steps.append((tuple(rx.shape), tuple(fx.shape), alpha, method, n_classes))
        self.counts = getattr(self, "counts", []) + [global_counts]
        self.first = getattr(self, "first", []) + [int(rx[0, 0, 0, 0] * 1000)]
        return torch.tensor(0.5)

    def save_checkpoint(self, path, step, write=True):
        self.ckpts.append((path, step, write))


def _fake_runner(monkeypatch, tmp_path, rank_world):
    import torch
    from types import SimpleNamespace
    from torch.utils.data import DataLoader, TensorDataset
    from unlearn_saliency_b200.diffusion import runner
    _FakeUnlearner.instances.clear()
    monkeypatch.setattr(runner, "DDPMEngineUnlearner", _FakeUnlearner)
    monkeypatch.setattr(runner, "_rank_world", lambda: rank_world)
    monkeypatch.setattr(runner.Diffusion, "_engine", lambda self, mb, precision="bf16": SimpleNamespace(
        close=lambda: None, state_dict=lambda prefix="": {}, precision=precision))
    monkeypatch.chdir(tmp_path)
    cfg = tiny_config()
    cfg.training = SimpleNamespace(batch_size=2, n_iters=5, snapshot_freq=2, log_freq=1)
    cfg.optim = SimpleNamespace(weight_decay=0.0, lr=1e-4, beta1=0.9, eps=1e-8, grad_clip=1.0)
    cfg.ckpt_dir = str(tmp_path / "ckpts")
    x = torch.arange(6).float().view(6, 1, 1, 1).expand(6, 3, 8, 8) / 1000.0     # batch i starts with image 2i
    forget = DataLoader(TensorDataset(x, torch.zeros(6, dtype=torch.long)), batch_size=2)
    remain = DataLoader(TensorDataset(x, torch.ones(6, dtype=torch.long)), batch_size=2)
    args = SimpleNamespace(ckpt_folder="ck", label_to_forget=4, cond_scale=2.0, mask_path=None, alpha=1e-3, method="rl", seed=7)
    return runner.Diffusion(args, cfg, device="cpu", loaders=(remain, forget)), cfg


def test_diffusion_mirror_control_flow_single_process(monkeypatch, tmp_path):
    r, cfg = _fake_runner(monkeypatch, tmp_path, (0, 1))
    r.generate_mask()
    un = _FakeUnlearner.instances[-1]
    assert un.batches == [0, 2, 4] and un.saliency.reduced == 1          # every forget batch, one all-reduce
    assert un.saliency.saved == [(os.path.join("results/cifar10/mask", "4", "with_0.5.pt"), 0.5, "module.")]
    assert r.saliency_unlearn() == 0.5
    un = _FakeUnlearner.instances[-1]
    assert len(un.steps) == 5 and un.steps[0] == ((2, 3, 8, 8), (2, 3, 8, 8), 1e-3, "rl", 10)
    assert un.ckpts == [(os.path.join(cfg.ckpt_dir, "ckpt.pth"), 1, True), (os.path.join(cfg.ckpt_dir, "ckpt.pth"), 3, True)]
    assert un.kw["grad_clip"] == 1.0 and un.kw["lr"] == 1e-4 and un.kw["mask"] is None


def test_diffusion_mirror_control_flow_rank1_of_2(monkeypatch, tmp_path):
    """data parallel: whole forget batches are dealt round-robin, every rank joins the all-reduce and the checkpoint
    gather, only rank 0 writes files"""
    import torch.distributed as dist
    monkeypatch.setattr(dist, "barrier", lambda *a, **k: None)
    r, cfg = _fake_runner(monkeypatch, tmp_path, (1, 2))
    r.generate_mask()
    un = _FakeUnlearner.instances[-1]
    assert un.batches == [2] and un.saliency.reduced == 1 and un.saliency.saved == []
    r.saliency_unlearn()
    un = _FakeUnlearner.instances[-1]
    assert [c[2] for c in un.ckpts] == [False, False] and len(un.steps) == 5
    # ONE global mini-batch of 2 + 2 per iteration, scattered: rank 1 gets the second sample of each (nn.DataParallel
    # semantics, runners/diffusion.py:505), and the loss means are taken over the global counts
    assert un.steps[0][:2] == ((1, 3, 8, 8), (1, 3, 8, 8)) and un.counts[0] == (2, 2)
    assert un.first[:3] == [1, 3, 5]
