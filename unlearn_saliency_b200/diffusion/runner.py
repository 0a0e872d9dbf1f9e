"""Loop bodies of the reference's DDPM SalUn runner (DDPM/runners/diffusion.py) on the fused sm_100a tail.

  get_beta_schedule           runners/diffusion.py:36-66 (linear / quad / const / jsd / sigmoid)
  eps_loss                    functions/losses.py:21-37  (noise_estimation_loss_conditional)
  antithetic_t                runners/diffusion.py:527-531, 966-970
  DDPMEngineUnlearner.generate_mask_batch / .finish_mask      :959-1039
  DDPMEngineUnlearner.saliency_unlearn_step                   :519-593
DDPMEngineUnlearner runs the U-Net forward / backward on the sm_100a engine (diffusion/engine.py, salun_unet_*): the
model calls of one iteration are concatenated into one batch (GroupNorm does not couple samples).  There is no PyTorch
fallback: a U-Net the engine does not serve is refused by salun_unet_create.  clip_grad_norm_, the mask
multiply (the reference re-uploads the 309 MB int64 mask every step, :589-592), Adam, the saliency accumulation (the
reference copies every gradient to the CPU every batch, :992-996) and the top-k (two CPU argsorts of 38.6 M keys,
:1006-1037) run in libsalun.so on flat arenas.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from ..flat import FlatMaskedAdam, FlatParams, FlatSaliency


def get_beta_schedule(beta_schedule: str, *, beta_start: float, beta_end: float, num_diffusion_timesteps: int):
    n = num_diffusion_timesteps
    if beta_schedule == "quad":
        betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=np.float64) ** 2
    elif beta_schedule == "linear":
        betas = np.linspace(beta_start, beta_end, n, dtype=np.float64)
    elif beta_schedule == "const":
        betas = beta_end * np.ones(n, dtype=np.float64)
    elif beta_schedule == "jsd":
        betas = 1.0 / np.linspace(n, 1, n, dtype=np.float64)
    elif beta_schedule == "sigmoid":
        x = np.linspace(-6, 6, n)
        betas = 1.0 / (np.exp(-x) + 1) * (beta_end - beta_start) + beta_start
    else:
        raise NotImplementedError(beta_schedule)
    assert betas.shape == (n,)
    return betas


def q_sample(x0, t, e, betas):
    """x_t = x0 sqrt(abar_t) + e sqrt(1 - abar_t)   (losses.py:31-32)"""
    a = (1 - betas).cumprod(dim=0).index_select(0, t).view(-1, 1, 1, 1)
    return x0 * a.sqrt() + e * (1.0 - a).sqrt()


def eps_loss(model, x0, t, c, e, betas, cond_drop_prob: float = 0.1, mode: str = "train", **kw):
    """sum over CHW of (e - eps_theta(x_t, t, c))^2, mean over the batch   (losses.py:21-37)"""
    x = q_sample(x0, t, e, betas)
    out = model(x, t.float(), c, cond_drop_prob=cond_drop_prob, mode=mode, **kw)
    return (e - out).square().sum(dim=(1, 2, 3)).mean(dim=0)


def antithetic_t(n: int, num_timesteps: int, device, generator=None):
    t = torch.randint(low=0, high=num_timesteps, size=(n // 2 + 1,), device=device, generator=generator)
    return torch.cat([t, num_timesteps - t - 1], dim=0)[:n]


class DDPMEngineUnlearner:
    """UNetEngine + fused tail: the loop bodies of Diffusion.generate_mask / Diffusion.saliency_unlearn with every model
    call on the sm_100a engine.  `mask`: the dict torch.load(mask_path) gives (CPU int64, ``module.`` keys); None = no mask.

    Data parallel (torch.distributed initialised, world W): each rank runs the samples it is handed with dL/d(eps)
    pre-scaled by 1/W, ONE all-reduce(sum) of the flat gradient arena (or the fused peer-memory exchange) gives the
    rank-averaged gradient, and every rank applies the identical clip + mask + Adam (clip uses the global norm,
    SURVEY.md section 7.3).  ``Diffusion.saliency_unlearn`` hands every rank its contiguous SHARD of one global
    mini-batch (what nn.DataParallel's scatter does, runners/diffusion.py:505) and passes ``global_counts`` so that the
    averaged gradient is exactly the gradient of the global-batch loss, also for uneven shards."""

    def __init__(self, engine, betas, lr=1e-4, beta1=0.9, eps=1e-8, weight_decay=0.0, grad_clip=1.0,
                 mask: Optional[Dict[str, torch.Tensor]] = None):
        self.engine = engine
        self.device = engine.device
        self.betas = torch.as_tensor(betas).float().to(self.device)
        self.num_timesteps = self.betas.shape[0]
        self.fused_dp = bool(getattr(engine, "symmetric", False)) and self._world() > 1
        if self.fused_dp:   # reduce-scatter + global clip + mask + Adam + all-gather over NVLink peer memory
            from .engine import DistMaskedAdam
            self.opt = DistMaskedAdam(engine, lr=lr, betas=(beta1, 0.999), eps=eps, weight_decay=weight_decay, mask=mask,
                                      max_norm=grad_clip)
        else:
            self.opt = FlatMaskedAdam(engine, lr=lr, betas=(beta1, 0.999), eps=eps, weight_decay=weight_decay, mask=mask,
                                      max_norm=grad_clip)
        self.saliency = FlatSaliency(engine, max_norm=grad_clip)
        self._step = 0
        from .engine import DDPMLoss
        self.loss_k = DDPMLoss(self.betas, engine.ctx)   # q-sample, eps losses and dL/d(eps) as kernels
        self._w_cache = {}
        # The no-grad pseudo-label pass (:561-569) is independent of the main forward until the loss: a forward-only
        # replica of the engine (same parameter arena, own activations) runs it on a second stream so that its HBM-bound
        # kernels overlap the main pass' tensor-core kernels and vice versa.  SALUN_DDPM_OVERLAP=0 keeps one stream.
        self._pseudo_engine, self._pseudo_stream = None, None
        self._overlap = os.environ.get("SALUN_DDPM_OVERLAP", "1") != "0"

    def _pseudo_forward(self, xt_f, tf_f, c_p, drop_p, train, seed):
        eng = self.engine
        if not self._overlap:
            return eng.forward(xt_f, tf_f, c_p, drop=drop_p, save=False, train=train, seed=seed)
        if self._pseudo_engine is None or self._pseudo_engine.max_batch < xt_f.shape[0]:
            from .engine import UNetEngine
            self._pseudo_engine = UNetEngine(eng.config, max_batch=max(xt_f.shape[0], 8), share_with=eng)
            self._pseudo_stream = torch.cuda.Stream(self.device)
        cur = torch.cuda.current_stream(self.device)
        self._pseudo_stream.wait_stream(cur)            # x_t, t, labels are ready
        with torch.cuda.stream(self._pseudo_stream):
            pseudo = self._pseudo_engine.forward(xt_f, tf_f, c_p, drop=drop_p, save=False, train=train, seed=seed)
        for tns in (xt_f, tf_f, c_p, drop_p, pseudo):
            if tns is not None:
                tns.record_stream(self._pseudo_stream)
        return pseudo

    def _pseudo_join(self):
        if self._overlap and self._pseudo_stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._pseudo_stream)

    def _weights(self, nr, nf, alpha, method, chw, scale, nr_w=None, nf_w=None):
        """per-sample loss weights: loss = sum_i w_i * sum_chw (eps_i - target_i)^2  (losses.py:33-37, diffusion.py:552-572);
        nr_w / nf_w: the divisor of the two batch means (the local counts unless the samples are a shard)."""
        nr_w, nf_w = float(nr if nr_w is None else nr_w), float(nf if nf_w is None else nf_w)
        key = (nr, nf, float(alpha), method, chw, float(scale), nr_w, nf_w)
        w = self._w_cache.get(key)
        if w is None:
            wf = 1.0 / (nf_w * chw) if method == "rl" else -1.0 / nf_w
            w = torch.cat([torch.full((nr,), alpha / nr_w), torch.full((nf,), wf)]).mul_(scale).to(self.device)
            self._w_cache[key] = w
        return w

    @staticmethod
    def _world():
        import torch.distributed as dist
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def _all_reduce_grads(self):
        if self._world() > 1:
            import torch.distributed as dist
            dist.all_reduce(self.engine.grads)

    def _drop(self, rng, key, n, p):
        if key in rng:
            return rng[key].to(self.device).to(torch.uint8)
        if p <= 0:
            return None
        return (torch.rand(n, device=self.device) < p).to(torch.uint8)   # keep_mask = uniform < 1 - p  (diffusion.py:8-14)

    # ---- runners/diffusion.py:959-996 -------------------------------------------------------------------------
    def generate_mask_batch(self, x, c, cond_scale: float = 2.0, t=None, e=None):
        """x in [0,1]; eval mode; the conditional and the null pass of _forward_with_cond_scale run as one batch of 2n."""
        eng, dev = self.engine, self.device
        x = x.to(dev).float()
        c = c.to(dev)
        n = x.shape[0]
        e = torch.randn_like(x) if e is None else e.to(dev)
        t = antithetic_t(n, self.num_timesteps, dev) if t is None else t.to(dev)
        xt = self.loss_k.q_sample(x, e, t, rescale=True)     # data_transform (2x - 1) + q-sample, one kernel
        tf = t.float()
        s = float(cond_scale)
        if s == 0:
            out = eng.forward(xt, tf, c, drop=None, save=True, train=False)
            d_out = (-2.0 / n) * (e - out)
            loss = (e - out).square().sum(dim=(1, 2, 3)).mean(dim=0)
            eng.backward(d_out.contiguous())
        elif 2 * n <= eng.max_batch:
            drop = torch.cat([torch.zeros(n, dtype=torch.uint8, device=dev), torch.ones(n, dtype=torch.uint8, device=dev)])
            eps2 = eng.forward(torch.cat([xt, xt]), torch.cat([tf, tf]), torch.cat([c, c]), drop=drop, save=True, train=False)
            out = (1 + s) * eps2[:n] - s * eps2[n:]
            loss = (e - out).square().sum(dim=(1, 2, 3)).mean(dim=0)
            d_out = (-2.0 / n) * (e - out)
            eng.backward(torch.cat([(1 + s) * d_out, -s * d_out]).contiguous())
        else:
            ones, zeros = torch.ones(n, dtype=torch.uint8, device=dev), torch.zeros(n, dtype=torch.uint8, device=dev)
            null = eng.forward(xt, tf, c, drop=ones, save=False, train=False)
            cond = eng.forward(xt, tf, c, drop=zeros, save=True, train=False)
            out = (1 + s) * cond - s * null
            loss = (e - out).square().sum(dim=(1, 2, 3)).mean(dim=0)
            d_out = (-2.0 / n) * (e - out)
            eng.backward(((1 + s) * d_out).contiguous())
            eng.forward(xt, tf, c, drop=ones, save=True, train=False)
            eng.backward((-s * d_out).contiguous(), accumulate=True)
        self.saliency.accumulate()  # clip to norm 1 (per batch, :985-990) then gradients += grad (:992-996)
        return loss.detach()

    def finish_mask(self, path: Optional[str] = None, ratio: float = 0.5, key_prefix: str = "module."):
        self.saliency.all_reduce()
        if path is None:
            return self.saliency.mask(ratio, key_prefix=key_prefix)
        return self.saliency.save(path, ratio, key_prefix=key_prefix)

    # ---- runners/diffusion.py:519-593 -------------------------------------------------------------------------
    def saliency_unlearn_step(self, remain_x, remain_c, forget_x, forget_c, alpha: float = 1e-3, method: str = "rl",
                              n_classes: int = 10, rng: Optional[dict] = None, train: bool = True,
                              global_counts: Optional[Tuple[int, int]] = None):
        """One iteration.  `rng` may carry externally drawn (t_r, e_r, t_f, e_f, drop_r, drop_f, drop_p) for parity runs;
        dropout inside the network uses the engine's counter-based generator (seeded per step).
        `global_counts` = (remain, forget) sizes of the GLOBAL mini-batch this rank's samples are a shard of: the loss
        means are then taken over the global batch (default: every rank's samples form its own mean)."""
        rng = rng or {}
        eng, dev = self.engine, self.device
        W = self._world()
        p_drop = eng.cond_drop_prob
        x01 = torch.cat([remain_x.to(dev, non_blocking=True).float(), forget_x.to(dev, non_blocking=True).float()])
        cr, cf = remain_c.to(dev, non_blocking=True), forget_c.to(dev, non_blocking=True)
        nr, nf = remain_x.shape[0], forget_x.shape[0]
        e = torch.cat([rng["e_r"].to(dev), rng["e_f"].to(dev)]) if "e_r" in rng else torch.randn_like(x01)
        t = (torch.cat([rng["t_r"].to(dev), rng["t_f"].to(dev)]) if "t_r" in rng else
             torch.cat([antithetic_t(nr, self.num_timesteps, dev), antithetic_t(nf, self.num_timesteps, dev)]))
        xt = self.loss_k.q_sample(x01, e, t, rescale=True)          # data_transform + q-sample (losses.py:31-32, :558-559)
        tf = t.float()
        drop_r, drop_f = self._drop(rng, "drop_r", nr, p_drop), self._drop(rng, "drop_f", nf, p_drop)
        self._step += 1
        seed = int(rng.get("seed", self._step)) * 2
        pseudo = None
        if method == "rl":
            drop_p = self._drop(rng, "drop_p", nf, p_drop)
            pseudo = self._pseudo_forward(xt[nr:], tf[nr:], (cf + 1) % n_classes, drop_p, train, seed + 1)   # :561-569 (no grad)
        elif method != "ga":
            raise NotImplementedError(method)
        if drop_r is None and drop_f is None:
            drop = None
        else:
            z = lambda k: torch.zeros(k, dtype=torch.uint8, device=dev)
            drop = torch.cat([drop_r if drop_r is not None else z(nr), drop_f if drop_f is not None else z(nf)])
        eps = eng.forward(xt, tf, torch.cat([cr, cf]), drop=drop, save=True, train=train, seed=seed)
        if pseudo is not None:
            self._pseudo_join()
            target = torch.cat([e[:nr], pseudo])
        else:
            target = e
        # loss = forget_loss + alpha * remain_loss (:533-572) and dL/d(eps), one kernel; the NCCL path averages the
        # gradient by pre-scaling dL/d(eps) with 1/W (the fused DP step averages inside its reduce kernel)
        scale = 1.0 / W if (W > 1 and not self.fused_dp) else 1.0
        # mean over the global batch: after the 1/W rank average a sample must weigh 1/n_global, i.e. W/n_global here
        nr_w, nf_w = (nr, nf) if global_counts is None else (global_counts[0] / W, global_counts[1] / W)
        loss, d, _ = self.loss_k.loss_grad(eps, target, self._weights(nr, nf, alpha, method, eps[0].numel(), scale, nr_w, nf_w))
        if scale != 1.0:
            loss = loss * W
        eng.backward(d)                                                                            # :579-580
        if not self.fused_dp:
            self._all_reduce_grads()
        self.opt.step()   # clip_grad_norm_(1.0) BEFORE the mask, grad *= mask, Adam -- one fused pass (:582-593)
        return loss[0]

    def save_checkpoint(self, path: str, step: int, write: bool = True):
        """states = [model_sd, optim_sd, step] like :598-610.  Data parallel: call on every rank (the fused DP step keeps
        only a shard of the Adam moments per rank; they are all-gathered here) with write=True on one of them."""
        m1, m2 = self.opt.gather_state() if self.fused_dp else (self.opt.exp_avg, self.opt.exp_avg_sq)
        if not write:
            return
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        sd = self.engine.state_dict(prefix="module.")
        optim = {"exp_avg": self.engine.dict_from_flat(m1), "exp_avg_sq": self.engine.dict_from_flat(m2),
                 "step": self.opt.step_count}
        torch.save([sd, optim, step], path)


def get_forget_dataset(args, config, label_to_drop):
    """(remain_loader, forget_loader) like DDPM/datasets/__init__.py:120-177: CIFAR-10 train split, Resize +
    RandomHorizontalFlip (config.data.random_flip) + ToTensor, split by label.  The data must already be under
    config.data.path (this mirror never downloads)."""
    from torch.utils.data import DataLoader
    from torchvision import transforms
    from torchvision.datasets import CIFAR10
    tf = [transforms.Resize(config.data.image_size)]
    if getattr(config.data, "random_flip", True):
        tf.append(transforms.RandomHorizontalFlip(p=0.5))
    tf.append(transforms.ToTensor())
    if config.data.dataset != "CIFAR10":
        raise NotImplementedError(f"dataset {config.data.dataset!r}: the SalUn DDPM configs served here use CIFAR10")
    dataset = CIFAR10(config.data.path, train=True, download=False, transform=transforms.Compose(tf))
    remain = [d for d in dataset if d[1] != label_to_drop]
    forget = [d for d in dataset if d[1] == label_to_drop]
    bs, nw = config.training.batch_size, getattr(config.data, "num_workers", 0)
    return (DataLoader(remain, batch_size=bs, shuffle=True, num_workers=nw),
            DataLoader(forget, batch_size=bs, shuffle=True, num_workers=nw))


def _rank_world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_batch(x, c, rank, world):
    """This rank's contiguous chunk of a global mini-batch (nn.DataParallel scatters contiguous chunks too; here the
    chunk sizes differ by at most one, so no rank is empty while n >= W)."""
    n = x.shape[0]
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    return x[lo:hi], c[lo:hi]


def _cycle(loader):
    while True:
        for batch in loader:
            yield batch


class Diffusion:
    """Mirror of the reference runner ``Diffusion(args, config)`` for the two SalUn modes that DDPM/train.py:150-155
    dispatches to: ``generate_mask()`` (runners/diffusion.py:947-1039) and ``saliency_unlearn()`` (:485-620).

    Same args (ckpt_folder, label_to_forget, cond_scale, mask_path, alpha, method, seed), same config keys
    (data.*, model.*, diffusion.*, training.{batch_size,n_iters,snapshot_freq,log_freq}, optim.*), same files: the
    checkpoint ``<ckpt_folder>/ckpts/ckpt.pth`` ([model_sd with ``module.`` keys, ...]) is read, the mask is written to
    ``results/cifar10/mask/<label>/with_0.5.pt`` (CPU int64 dict, ``module.`` keys) and snapshots go to
    ``<config.ckpt_dir>/ckpt.pth`` as [model_sd, optim_sd, step].  The U-Net runs on the sm_100a engine, and so does the
    DDIM sampling of ``sample_visualization`` / ``visualization`` (:598-619, diffusion/sampler.py); the FID network is
    outside the path (``on_snapshot`` is called with (step, state_dict) for callers who score snapshots inline).

    ``loaders=(remain_loader, forget_loader)`` overrides get_forget_dataset (tests, synthetic data)."""

    def __init__(self, args, config, device=None, loaders=None, on_snapshot=None):
        self.args, self.config = args, config
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        d = config.diffusion
        self.betas = torch.from_numpy(get_beta_schedule(d.beta_schedule, beta_start=d.beta_start, beta_end=d.beta_end,
                                                        num_diffusion_timesteps=d.num_diffusion_timesteps)).float()
        self.num_timesteps = self.betas.shape[0]
        self.loaders, self.on_snapshot = loaders, on_snapshot

    def _loaders(self):
        if self.loaders is not None:
            return self.loaders
        return get_forget_dataset(self.args, self.config, self.args.label_to_forget)

    def _engine(self, max_batch, precision="bf16"):
        from .engine import UNetEngine
        precision = getattr(self.args, "precision", None) or precision
        eng = UNetEngine(self.config, max_batch=max_batch, device=self.device, precision=precision)
        path = os.path.join(self.args.ckpt_folder, "ckpts/ckpt.pth")
        states = torch.load(path, map_location="cpu")
        eng.load_state_dict(states[0], strict=True)           # DataParallel keys: the ``module.`` prefix is stripped
        return eng

    def generate_mask(self, threshold_list=(0.5,)):
        args, config = self.args, self.config
        _, forget_loader = self._loaders()
        # the saliency pass decides an index set: it runs on the split-precision build (fp32-class products, DESIGN.md
        # section 4) unless args.precision says otherwise; conditional + null pass as one batch
        eng = self._engine(2 * config.training.batch_size, precision="split")
        o = config.optim
        un = DDPMEngineUnlearner(eng, self.betas, lr=o.lr, beta1=o.beta1, eps=o.eps, weight_decay=o.weight_decay,
                                 grad_clip=o.grad_clip)
        if getattr(args, "seed", None) is not None:
            torch.manual_seed(args.seed)      # the same loader order on every rank
        rank, world = _rank_world()
        for i, (x, forget_c) in enumerate(forget_loader):
            # data parallel: whole batches are dealt round-robin, so the per-batch clip (:985-990)
            # sees exactly the batches of the single-process run; ONE all-reduce of the accumulator follows
            if i % world != rank:
                continue
            un.generate_mask_batch(x, forget_c, cond_scale=args.cond_scale)
        mask_path = os.path.join("results/cifar10/mask", str(args.label_to_forget))     # :1003
        infos = {}
        un.saliency.all_reduce()
        for i in threshold_list:
            path = os.path.join(mask_path, f"with_{str(i)}.pt") if rank == 0 else None   # rank 0 writes the file
            infos[i] = un.saliency.save(path, i, key_prefix="module.") if path else un.saliency.mask(i, key_prefix="module.")[1]
        if world > 1:
            torch.distributed.barrier()
        eng.close()
        return infos

    def sample_visualization(self, eng, name, cond_scale):
        """runners/diffusion.py:877-931: visualization_samples images (the same number per class) through the engine's
        class-conditional sampler, saved as <log_dir>/sample-<name>.png.  Flags as in DDPM/train.py (--sample_type
        generalized, --skip_type, --timesteps, --eta)."""
        from .sampler import EngineSampler
        args, config = self.args, self.config
        total = int(config.training.visualization_samples)
        bs = min(int(config.sampling.batch_size), eng.max_batch // 2)
        was_training = eng.training
        eng.eval()
        sm = EngineSampler(eng, self.betas)
        out_dir = getattr(config, "log_dir", None) or args.ckpt_folder
        imgs = sm.sample_visualization(config.data.n_classes, total, bs, cond_scale, config.data.image_size,
                                       config.data.channels, path=os.path.join(out_dir, f"sample-{name}.png"),
                                       sample_type=getattr(args, "sample_type", "generalized"),
                                       skip_type=getattr(args, "skip_type", "uniform"),
                                       timesteps=int(getattr(args, "timesteps", 1000)), eta=float(getattr(args, "eta", 1.0)))
        eng.train(was_training)
        return imgs

    def visualization(self):
        """DDPM/train.py --mode visualization (runners/diffusion.py:635-668 `sample`): load the checkpoint, sample, save"""
        eng = self._engine(2 * int(self.config.sampling.batch_size))
        imgs = self.sample_visualization(eng, str(self.args.cond_scale), self.args.cond_scale)
        eng.close()
        return imgs

    def saliency_unlearn(self):
        args, config = self.args, self.config
        remain_loader, forget_loader = self._loaders()
        remain_iter, forget_iter = _cycle(remain_loader), _cycle(forget_loader)
        mask = torch.load(args.mask_path) if getattr(args, "mask_path", None) else None   # :493-496
        eng = self._engine(2 * config.training.batch_size)      # remain + forget mini-batch as one batch
        o = config.optim
        un = DDPMEngineUnlearner(eng, self.betas, lr=o.lr, beta1=o.beta1, eps=o.eps, weight_decay=o.weight_decay,
                                 grad_clip=o.grad_clip, mask=mask)
        if getattr(config.model, "ema", False):
            raise NotImplementedError("model.ema=True: the SalUn unlearning config sets ema False (cifar10_saliency_unlearn.yml:23)")
        rank, world = _rank_world()
        if getattr(args, "seed", None) is not None:
            # data parallel = nn.DataParallel semantics (runners/diffusion.py:505): ONE global mini-batch per iteration,
            # scattered across the ranks.  The loaders (CPU generator) therefore run in lock-step on every rank; the
            # noise / timestep / dropout draws (CUDA generator) are per rank.
            torch.manual_seed(args.seed)
            if world > 1 and self.device.type == "cuda":
                torch.cuda.manual_seed(args.seed + rank)
        import logging
        import time
        start = time.time()
        loss = None
        for step in range(config.training.n_iters):
            remain_x, remain_c = next(remain_iter)
            forget_x, forget_c = next(forget_iter)
            counts = None
            if world > 1:
                counts = (remain_x.shape[0], forget_x.shape[0])
                (remain_x, remain_c), (forget_x, forget_c) = (shard_batch(remain_x, remain_c, rank, world),
                                                              shard_batch(forget_x, forget_c, rank, world))
                if min(counts) < world:   # a trailing partial batch too small to give every rank a sample: skipped by all
                    continue
            loss = un.saliency_unlearn_step(remain_x, remain_c, forget_x, forget_c, alpha=args.alpha, method=args.method,
                                            n_classes=config.data.n_classes, global_counts=counts)
            if (step + 1) % config.training.log_freq == 0:
                logging.info(f"step: {step}, loss: {loss.item()}, time: {time.time() - start}")
                start = time.time()
            if (step + 1) % config.training.snapshot_freq == 0:
                # every rank calls (the sharded Adam moments are all-gathered); rank 0 writes the file
                un.save_checkpoint(os.path.join(config.ckpt_dir, "ckpt.pth"), step, write=(rank == 0))
                if self.on_snapshot is not None and rank == 0:
                    self.on_snapshot(step, eng.state_dict(prefix="module."))
                if rank == 0 and getattr(args, "visualize", False):
                    # :611-619 sample_visualization(test_model, step, cond_scale) on the engine's forward kernels
                    self.sample_visualization(eng, step, args.cond_scale)
        eng.close()
        return None if loss is None else float(loss)
