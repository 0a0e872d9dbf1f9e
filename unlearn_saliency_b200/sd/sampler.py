"""DDIM sampling of the Stable-Diffusion U-Net on the engine: the no-grad passes of ESD.

Mirrors, by name and argument meaning,
  make_ddim_timesteps / make_ddim_sampling_parameters / make_beta_schedule("linear")
                                        SD/ldm/modules/diffusionmodules/util.py:20-96
  DDIMSampler.make_schedule / sample / ddim_sampling / p_sample_ddim (t_start, till_T, classifier-free guidance)
                                        SD/ldm/models/diffusion/ddim.py:37-100,103-282,285-362
  sample_model / quick_sample_till_t    SD/train-scripts/train-esd.py:60-127,262-283
Every U-Net evaluation is SDUNetEngine.forward (one CUDA-graph replay; unconditional | conditional halves as one batch of
2n, ddim.py:309-314) and every x_{t-1} update is one salun_ddim_step launch (guidance blend, predicted x0, direction term and
noise in a single pass); nothing is copied to the host between steps.  No PyTorch fallback: without libsalun the import
of the engine fails.
"""
from typing import Optional

import numpy as np
import torch

from .._lib import check
from ..engine import _ptr, _stream
from .engine import SDUNetEngine


def make_beta_schedule(schedule="linear", n_timestep=1000, linear_start=0.00085, linear_end=0.012):
    if schedule != "linear":
        raise ValueError("Stable Diffusion v1 uses the 'linear' (sqrt-space) schedule")
    return (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2).numpy()


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=False):
    if ddim_discr_method == "uniform":
        c = num_ddpm_timesteps // num_ddim_timesteps
        steps = np.asarray(list(range(0, num_ddpm_timesteps, c)))
    elif ddim_discr_method == "quad":
        steps = ((np.linspace(0, np.sqrt(num_ddpm_timesteps * .8), num_ddim_timesteps)) ** 2).astype(int)
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
    return steps + 1      # "add one to get the final alpha values right"


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=False):
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return sigmas, alphas, alphas_prev


class EngineDDIMSampler:
    """DDIMSampler over an SDUNetEngine.  `engine.max_batch` must cover 2x the sampled batch when guidance is on."""

    def __init__(self, engine: SDUNetEngine, alphas_cumprod=None, schedule="linear", **kwargs):
        self.engine = engine
        self.device = engine.device
        self.alphas_cumprod = np.asarray(alphas_cumprod.detach().cpu().numpy() if torch.is_tensor(alphas_cumprod) else
                                         alphas_cumprod if alphas_cumprod is not None else
                                         np.cumprod(1.0 - make_beta_schedule(schedule)), dtype=np.float64)
        self.ddpm_num_timesteps = int(self.alphas_cumprod.shape[0])
        self.ddim_timesteps = None

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=False):
        self.ddim_timesteps = make_ddim_timesteps(ddim_discretize, ddim_num_steps, self.ddpm_num_timesteps)
        sig, a, ap = make_ddim_sampling_parameters(self.alphas_cumprod, self.ddim_timesteps, ddim_eta)
        f32 = lambda v: torch.tensor(np.asarray(v), dtype=torch.float32, device=self.device)
        self.ddim_sigmas, self.ddim_alphas, self.ddim_alphas_prev = f32(sig), f32(a), f32(ap)
        self.ddim_sqrt_one_minus_alphas = f32(np.sqrt(1. - a))
        self.ddim_eta = float(ddim_eta)

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, eta=0., x_T=None, verbose=False, log_every_t=100,
               unconditional_guidance_scale=1., unconditional_conditioning=None, t_start=-1, till_T=None, noise=None,
               **kwargs):
        if conditioning is not None and conditioning.shape[0] != batch_size:
            print(f"Warning: Got {conditioning.shape[0]} conditionings but batch-size is {batch_size}")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta)
        C, H, W = shape
        return self.ddim_sampling(conditioning, (batch_size, C, H, W), x_T=x_T, log_every_t=log_every_t,
                                  unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning, t_start=t_start, till_T=till_T,
                                  noise=noise)

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, x_T=None, log_every_t=100, unconditional_guidance_scale=1.,
                      unconditional_conditioning=None, t_start=-1, till_T=None, noise=None):
        dev, b = self.device, shape[0]
        img = (torch.randn(shape, device=dev) if x_T is None else x_T.to(dev, torch.float32)).contiguous().clone()
        timesteps = self.ddim_timesteps[:t_start]
        total_steps = timesteps.shape[0]
        till = till_T if till_T is not None else 0
        intermediates = {"x_inter": [img.clone()], "pred_x0": [img.clone()]}
        cfg = unconditional_conditioning is not None and unconditional_guidance_scale != 1.
        c_in = torch.cat([unconditional_conditioning, cond]).to(dev, torch.float32) if cfg else cond.to(dev, torch.float32)
        ts_all = torch.as_tensor(np.ascontiguousarray(timesteps), device=dev, dtype=torch.float32)
        nxt, x0 = torch.empty_like(img), torch.empty_like(img)
        lib, h, s = self.engine._lib, self.engine.ctx.handle, _stream(dev)
        chw = img[0].numel()
        for i in range(total_steps):
            index = total_steps - i - 1
            ts = ts_all[index].expand(b)
            if cfg:
                eps = self.engine.forward(torch.cat([img, img]), torch.cat([ts, ts]), c_in)
                e_uncond, e_cond = eps[:b], eps[b:]
            else:
                e_uncond, e_cond = None, self.engine.forward(img, ts, c_in)
            nz = None
            if self.ddim_eta != 0.:
                nz = noise[i].to(dev) if noise is not None else torch.randn_like(img)
            # e_t = e_uncond + scale (e_cond - e_uncond) = (1 + (scale-1)) e_cond - (scale-1) e_uncond   (ddim.py:314)
            a_t = self.ddim_alphas[index].expand(b).contiguous()
            a_prev = self.ddim_alphas_prev[index].expand(b).contiguous()
            check(lib.salun_ddim_step(h, _ptr(e_cond), _ptr(e_uncond) if cfg else None, _ptr(img),
                                      _ptr(nz) if nz is not None else None, _ptr(a_t), _ptr(a_prev),
                                      float(unconditional_guidance_scale) - 1.0, self.ddim_eta, b, chw, _ptr(nxt), _ptr(x0), s),
                  "salun_ddim_step", lib)
            img, nxt = nxt, img
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates["x_inter"].append(img.clone())
                intermediates["pred_x0"].append(x0.clone())
            if index + 1 == till:
                break
        return img.clone(), intermediates


class EngineApplyModel:
    """model_orig of train-esd.py (the frozen copy: e_0 and e_p, :295-297) as an engine: apply_model(x, t, c) -> eps"""

    def __init__(self, engine: SDUNetEngine):
        self.engine = engine

    @torch.no_grad()
    def apply_model(self, x_noisy, t, cond):
        return self.engine.forward(x_noisy.to(self.engine.device, torch.float32), t, cond.to(self.engine.device, torch.float32))

    @torch.no_grad()
    def apply_model_pair(self, x_noisy, t, cond_a, cond_b):
        """eps for two conditionings of the SAME (x, t) as one batch of 2n (e_0 and e_p of train-esd.py:295-297)"""
        n = x_noisy.shape[0]
        dev = self.engine.device
        x2 = torch.cat([x_noisy, x_noisy]).to(dev, torch.float32)
        t2 = torch.cat([t, t]) if t.numel() == n else t.expand(2 * n)
        eps = self.engine.forward(x2, t2, torch.cat([cond_a, cond_b]).to(dev, torch.float32))
        return eps[:n], eps[n:]

    def eval(self):
        return self


def sample_model(sampler: EngineDDIMSampler, h, w, ddim_steps, scale, ddim_eta, c, uc=None, start_code=None, n_samples=1,
                 t_start=-1, log_every_t=None, till_T=None, verbose=True):
    """train-esd.py:60-96 (the learned empty-prompt conditioning `uc` is the caller's: the text encoder is not on this path)"""
    log_t = 100 if log_every_t is None else log_every_t
    shape = [4, h // 8, w // 8]
    samples, inters = sampler.sample(S=ddim_steps, conditioning=c, batch_size=n_samples, shape=shape, verbose=False,
                                     x_T=start_code, unconditional_guidance_scale=scale,
                                     unconditional_conditioning=uc if scale != 1.0 else None, eta=ddim_eta,
                                     log_every_t=log_t, t_start=t_start, till_T=till_T)
    return (samples, inters) if log_every_t is not None else samples


def make_quick_sample_till_t(sampler: EngineDDIMSampler, uncond: torch.Tensor, image_size=512, ddim_steps=50, ddim_eta=0.,
                             module: Optional[torch.nn.Module] = None, train_keys=None):
    """quick_sample_till_t of train-esd.py:281-283 for loops.esd_iteration(sample_fn=...).  When `module` (the trained
    UNetModel) is given its current parameters are pushed into the engine first -- ESD samples from the model it trains."""
    def quick_sample_till_t(x, s, code, t):
        if module is not None:
            sd = dict(module.named_parameters())
            sampler.engine.update_parameters(sd if train_keys is None else {k: sd[k] for k in train_keys})
        return sample_model(sampler, image_size, image_size, ddim_steps, s, ddim_eta, x, uc=uncond, start_code=code,
                            till_T=t, verbose=False)
    return quick_sample_till_t
