#!/usr/bin/env python
"""bench.py -- SalUn masked-unlearning steps/sec (ResNet-18 / CIFAR-10 shape; DDPM U-Net 32x32) on N B200s.

    python bench.py --gpus 1 --steps 200 --warmup 20
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W [--scaling strong]
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores
    python bench.py --impl torch ...          # the reference's loop statements in stock PyTorch eager on one B200

Headline workload (BASELINE.json configs[1]): one step = one mini-batch of the SalUn random-label
unlearning loop (Classification/unlearn/RL.py:123-140): train-mode ResNet-18 forward + backward on 256 synthetic
32x32 images, mask (.) grad, SGD(momentum 0.9, wd 5e-4, lr 0.013), restore -- with a 50% saliency mask.
  --scaling weak   (default) 256 images per GPU, global batch 256*N, one gradient exchange per step
  --scaling strong the survey's split (section 8e): global batch 256 sharded 256/N per GPU (DDPM: (128+128)/N per GPU)
Sub-lines of the same JSON object: "ddpm" (configs[2]: one saliency_unlearn iteration of the cifar10 U-Net), "maskgen"
(hot path (i): saliency accumulation over a 512-image forget set + the ten top-k selects), "modes" (the step in the
split-precision build), "torch_eager" (the reference statements in stock PyTorch on this GPU, TF32 and fp32: the
denominator of the 10x target), "parity" (the acceptance numbers tests/test_acceptance_gpu.py measured on a B200,
profiles/r2_acceptance_*.json).

Prints ONE JSON line (rank 0).  `value` times the step with inputs resident in HBM; `e2e` times the public
Python API with pinned HOST inputs (H2D of images + labels and D2H of the loss inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# Keep stdout to the ONE JSON line.  NCCL (every rank) writes its version banner / INFO lines to stdout at whatever
# NCCL_DEBUG level the environment sets; the level is left alone (the driver's rank check reads those lines) but the
# process-level stdout (fd 1) is pointed at stderr, and the JSON line is written to a private duplicate of the real stdout.
if not os.environ.get("NCCL_DEBUG_FILE"):
    os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
sys.stdout.flush()
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


BATCH = 256
FWD_GFLOP_PER_IMG = 1.1108  # SURVEY.md section 6 (torch.utils.flop_counter on the reference model)
METRIC = "unlearn steps/sec (ResNet-18 CIFAR-10 SalUn RL masked step, batch 256 per GPU)"
METRIC_STRONG = "unlearn steps/sec (ResNet-18 CIFAR-10 SalUn RL masked step, global batch 256 sharded over the GPUs)"
N_PARAMS_RN18 = 11173962
DDPM_BATCH = 128
DDPM_FWD_GFLOP = 12.449


def load_peaks():
    """(hbm GB/s, bf16 TF/s burst, bf16 TF/s sustained, source).  The timed regions here are sub-second at a steady SM
    clock, so the roofline denominators are the BURST figures."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        burst = d.get("bf16_tflops", 1590.0)
        return d.get("hbm_gbs", 6650.0), burst, d.get("bf16_tflops_sustained", burst), "measured (MEASURED_PEAKS.json, burst)"
    return 6650.0, 1590.0, 1590.0, "fallback (B200_PROFILING.md)"


def load_json(rel):
    p = os.path.join(ROOT, rel)
    try:
        return json.load(open(p))
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons DURING the timed region"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(float(self.rows[0][1])) if self.rows else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# =====================================================================================================================
# CPU arm: the reference's own CPU path (oracle ports, torch fp32, all host threads) at the REAL batch sizes
# =====================================================================================================================
def oracle_cpu_resnet(steps=2, warmup=1, batch=BATCH):
    """RL.py:123-140 step at batch 256 through oracle/classification.py.  Returns (seconds per step, threads, sample)."""
    import torch
    from oracle import classification as OC
    torch.set_num_threads(os.cpu_count())
    params, buffers = OC.synth_state(10, seed=0)
    g = torch.Generator().manual_seed(1)
    flat_mask = (torch.rand(N_PARAMS_RN18, generator=g) < 0.5).to(torch.int64)
    mask = OC.split_mask(flat_mask, OC.resnet18_param_shapes(10))
    opt = OC.MaskedSGD(params, mask, lr=0.013, momentum=0.9, wd=5e-4)
    x = torch.rand(batch, 3, 32, 32, generator=g)
    times = []
    for s in range(warmup + steps):
        y = torch.randint(0, 10, (batch,), generator=g)
        t0 = time.perf_counter()
        OC.unlearn_step(params, buffers, opt, x, y)
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    sample = f"{steps} RL steps of {batch} images after {warmup} warm-up (oracle/classification.py, torch fp32 CPU, no scaling)"
    return sum(times) / len(times), torch.get_num_threads(), sample


def oracle_cpu_ddpm(steps=1, warmup=0, batch=DDPM_BATCH):
    """One saliency_unlearn iteration (runners/diffusion.py:519-593) at 128 + 128 images through oracle/ddpm.py."""
    import torch
    from oracle import ddpm as OD
    from oracle.unet import ConditionalUNet
    from unlearn_saliency_b200.diffusion.config import cifar10_config
    from unlearn_saliency_b200.diffusion.runner import get_beta_schedule
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    model = ConditionalUNet(cifar10_config())
    g = torch.Generator().manual_seed(1)
    mask = {k: (torch.rand(p.shape, generator=g) < 0.5).to(torch.int64) for k, p in model.named_parameters()}
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    betas = torch.from_numpy(get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)).float()
    n, S = batch, 32
    times = []
    for s in range(warmup + steps):
        r = dict(x_r=torch.rand(n, 3, S, S, generator=g), c_r=torch.randint(1, 10, (n,), generator=g),
                 x_f=torch.rand(n, 3, S, S, generator=g), c_f=torch.zeros(n, dtype=torch.long),
                 t_r=torch.randint(0, 1000, (n,), generator=g), e_r=torch.randn(n, 3, S, S, generator=g),
                 t_f=torch.randint(0, 1000, (n,), generator=g), e_f=torch.randn(n, 3, S, S, generator=g),
                 drop_r=torch.rand(n, generator=g) < 0.1, drop_f=torch.rand(n, generator=g) < 0.1,
                 drop_p=torch.rand(n, generator=g) < 0.1)
        t0 = time.perf_counter()
        OD.saliency_unlearn_step(model, opt, mask, r, betas, alpha=1e-3, method="rl", keep_raw=False)
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    sample = f"{steps} iteration(s) of {n}+{n} images after {warmup} warm-up (oracle/ddpm.py, torch fp32 CPU, no scaling)"
    return sum(times) / len(times), torch.get_num_threads(), sample


def cpu_baselines(ddpm=True, resnet_steps=2):
    sec, threads, sample = oracle_cpu_resnet(steps=resnet_steps, warmup=1)
    out = {"resnet": {"value": 1.0 / sec, "unit": "steps/s", "cores": threads, "kind": "port", "sample": sample}}
    if ddpm:
        dsec, dthreads, dsample = oracle_cpu_ddpm(steps=1, warmup=0)
        out["ddpm"] = {"value": 1.0 / dsec, "unit": "iterations/s", "cores": dthreads, "kind": "port", "sample": dsample}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baselines(ddpm=not args.no_ddpm, resnet_steps=max(1, min(args.steps, 3)))
    val = cb["resnet"]["value"]
    line = {
        "impl": "reference", "metric": METRIC_STRONG if (args.scaling == "strong" and args.gpus > 1) else METRIC, "value": val, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ResNet-18/CIFAR-10 SalUn RL masked unlearn step, batch 256, mask ratio 0.5", "global_batch": BATCH},
        "cpu_baseline": cb["resnet"],
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if "ddpm" in cb:
        line["ddpm"] = {"metric": "DDPM saliency_unlearn iterations/sec (cifar10 U-Net 32x32, 128 remain + 128 forget images, rl)",
                        "value": cb["ddpm"]["value"], "unit": "iterations/s", "dtype": "f32", "cpu_baseline": cb["ddpm"]}
    emit(line)


# =====================================================================================================================
# torch arm: the reference's statements in stock PyTorch eager on ONE B200 (baseline/torch_eager.py)
# =====================================================================================================================
def torch_eager_numbers(ddpm=True, steps=20, ddpm_steps=5):
    from baseline import torch_eager as TE
    out = {"how": "baseline/torch_eager.py: stock torch.nn / autograd / torch.optim, the loop statements of RL.py:11-34,123-140 and "
                  "runners/diffusion.py:519-593, inputs from pinned host memory, 1 GPU"}
    for mode in ("tf32", "fp32"):
        v, ms = TE.resnet18_rl_steps_per_sec(batch=BATCH, steps=steps, warmup=5, mode=mode)
        out[f"resnet18_{mode}"] = {"value": v, "unit": "steps/s", "ms_per_step": ms}
    if ddpm:
        for mode in ("tf32", "fp32"):
            v, ms = TE.ddpm_unlearn_its_per_sec(batch=DDPM_BATCH, steps=ddpm_steps, warmup=2, mode=mode)
            out[f"ddpm_{mode}"] = {"value": v, "unit": "iterations/s", "ms_per_it": ms}
    out["note"] = "tf32 = torch defaults (cuDNN convolutions in TF32): the arithmetic the reference runs on this GPU"
    return out


def run_torch(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("--impl torch needs a CUDA device")
    t = torch_eager_numbers(ddpm=not args.no_ddpm, steps=max(5, min(args.steps, 50)))
    val = t["resnet18_tf32"]["value"]
    line = {"impl": "torch", "metric": METRIC, "value": val, "unit": "steps/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32", "data": "synthetic",
            "config": {"workload": "ResNet-18/CIFAR-10 SalUn RL masked unlearn step, batch 256, mask ratio 0.5 (stock PyTorch eager)"},
            "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": BATCH * 3 * 32 * 32 * 4 + BATCH * 8, "d2h_bytes_per_step": 4},
            "gpu_launches": 0, "torch_eager": t}
    emit(line)


# =====================================================================================================================
# our arm
# =====================================================================================================================
def rn18_state(eng, torch):
    """random-init weights of the reference architecture (kaiming-normal fan_out convs, unit BN), same on every rank"""
    g = torch.Generator(device="cpu").manual_seed(0)
    sd = {}
    for k, shp in eng.table.items():
        if len(shp) == 4:
            sd[k] = torch.randn(shp, generator=g) * (2.0 / (shp[0] * shp[2] * shp[3])) ** 0.5
        elif k == "fc.weight":
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) / shp[1] ** 0.5
        elif k.endswith(".weight"):
            sd[k] = torch.ones(shp)
        else:
            sd[k] = torch.zeros(shp)
    return sd, g


def bench_maskgen(dev, timed_ms, hbm_peak):
    """Hot path (i) on one GPU (BASELINE configs[0] shape): Classification/generate_mask.py:30-82 on a 512-image forget set,
    batch 256 -- eval-mode forward/backward of -CE + flat accumulate per batch, then |.| and the ten top-k selects with
    their int64 mask outputs.  Both engine builds; the split build is what the generate_mask mirror uses."""
    import torch
    from unlearn_saliency_b200 import _lib
    from unlearn_saliency_b200.engine import ResNetEngine
    from unlearn_saliency_b200.tail import topk_count
    res = {"workload": "ResNet-18 saliency mask, 512 forget images, batch 256, ratios 0.1..1.0 (generate_mask.py:30-82)"}
    g = torch.Generator(device="cpu").manual_seed(3)
    xs = [torch.rand(BATCH, 3, 32, 32, generator=g).to(dev) for _ in range(2)]
    ys = [torch.randint(0, 10, (BATCH,), generator=g).to(dev) for _ in range(2)]
    ratios = [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0]
    for prec in _lib.available_precisions():
        eng = ResNetEngine("resnet18", 10, 32, max_batch=BATCH, device=dev, precision=prec)
        sd, _ = rn18_state(eng, torch)
        eng.load_state_dict(sd)
        eng.eval()
        acc = torch.zeros_like(eng.params)

        def accumulate(i):
            for x, y in zip(xs, ys):
                eng.forward_backward(x, y, loss_sign=-1.0, train=False)
                eng.ctx.saliency_accumulate_flat(eng.grads, acc)

        for i in range(3):
            accumulate(i)
        ms_acc = timed_ms(accumulate, 5) / 5
        flat = eng.from_native_flat(acc).abs_().contiguous()
        n = flat.numel()

        def select_single(i):
            for r in ratios:
                eng.ctx.topk_mask(flat, topk_count(n, r), want_info=False)

        def select_all(i):      # what save_gradient_ratio calls: one sweep for the whole threshold_list
            eng.ctx.topk_mask_multi(flat, [topk_count(n, r) for r in ratios], want_info=False)

        for i in range(2):
            select_single(i)
            select_all(i)
        ms_single = timed_ms(select_single, 5) / 5
        ms_sel = timed_ms(select_all, 5) / 5
        # one sweep: 5 reads of |g| (3 histogram passes, tie count, write pass) + int64 mask and packed bits per ratio
        sweep_bytes = n * (5 * 4.0 + len(ratios) * 8.125)
        res[prec] = {"accumulate_ms_per_512_images": ms_acc, "images_per_s": 512e3 / ms_acc,
                     "tflops": 512 * 3 * FWD_GFLOP_PER_IMG / ms_acc, "select_ms_per_ratio": ms_sel / len(ratios),
                     "select_ms_all_ratios": ms_sel, "select_ms_all_ratios_one_call_per_ratio": ms_single,
                     "select_hbm_gbs": sweep_bytes / (ms_sel * 1e-3) / 1e9,
                     "select_hbm_frac": sweep_bytes / (ms_sel * 1e-3) / 1e9 / hbm_peak,
                     "total_ms": ms_acc + ms_sel}
        eng.close()
        del eng
        torch.cuda.empty_cache()
    return res


def bench_ddpm(args, dev, rank, world, L, timed, tf_peak, peak_src, precision="bf16", per_gpu=DDPM_BATCH, roofline=True):
    """Second headline workload (BASELINE.json configs[2]): one DDPM saliency_unlearn iteration (runners/diffusion.py:519-593)
    on the cifar10 U-Net, `per_gpu` remain + `per_gpu` forget images per GPU, method rl, dropout 0.1, 50% mask, clip 1.0,
    Adam; N > 1: the clip + mask + Adam tail is the fused peer-memory exchange (or one NCCL all-reduce)."""
    import ctypes as C
    import torch
    from unlearn_saliency_b200.diffusion.config import cifar10_config
    from unlearn_saliency_b200.diffusion.engine import UNetEngine
    from unlearn_saliency_b200.diffusion.runner import DDPMEngineUnlearner, get_beta_schedule
    cfg = cifar10_config()
    B = per_gpu
    fused_dp = world > 1 and os.environ.get("SALUN_FUSED_DP", "1") != "0"
    try:
        eng = UNetEngine(cfg, max_batch=2 * B, device=dev, symmetric=fused_dp, precision=precision)
    except Exception as e:  # symmetric memory unavailable: NCCL all-reduce + local clip / mask / Adam (same arithmetic)
        print(f"[bench] symmetric memory unavailable for the DDPM arenas ({e!r}); using NCCL all-reduce", file=sys.stderr)
        fused_dp = False
        eng = UNetEngine(cfg, max_batch=2 * B, device=dev, precision=precision)
    g = torch.Generator(device="cpu").manual_seed(0)  # same random-init weights on every rank
    sd = {}
    for k, shp in eng.shapes.items():
        if "norm" in k:
            sd[k] = torch.ones(shp) if k.endswith("weight") else torch.zeros(shp)
        elif len(shp) >= 2:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) / fan_in ** 0.5
        else:
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * 0.05
    eng.load_state_dict(sd)
    mask_native = (torch.rand(eng.n, generator=g) < 0.5).to(torch.int64).to(dev)
    betas = torch.from_numpy(get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)).float()
    un = DDPMEngineUnlearner(eng, betas, lr=1e-4, grad_clip=1.0)
    un.opt.mask_bits = eng.ctx.pack_mask(mask_native)
    gen = torch.Generator(device="cpu").manual_seed(200 + rank)
    host = [(torch.rand(B, 3, 32, 32, generator=gen).pin_memory(), torch.randint(1, 10, (B,), generator=gen).pin_memory(),
             torch.rand(B, 3, 32, 32, generator=gen).pin_memory(), torch.zeros(B, dtype=torch.long).pin_memory())
            for _ in range(4)]
    resident = [tuple(t.to(dev) for t in h) for h in host]
    counts = (B * world, B * world) if args.scaling == "strong" and world > 1 else None

    def it_resident(i):
        xr, cr, xf, cf = resident[i % 4]
        un.saliency_unlearn_step(xr, cr, xf, cf, alpha=1e-3, method="rl", global_counts=counts)

    last = [0.0]

    def it_e2e(i):
        xr, cr, xf, cf = host[i % 4]  # pinned host tensors: the step copies them to the device (H2D inside the timed region)
        last[0] = float(un.saliency_unlearn_step(xr, cr, xf, cf, alpha=1e-3, method="rl", global_counts=counts).item())  # D2H

    steps = max(1, min(args.steps, args.ddpm_steps))
    for i in range(3):
        it_resident(i)
    l0 = L.salun_launch_count()
    ms = timed(it_resident, steps) / steps
    launches = L.salun_launch_count() - l0
    for i in range(3):
        it_e2e(i)
    ms_e2e = timed(it_e2e, steps) / steps
    roof = None
    if roofline:
        # instrumented replay on EVERY rank (the iteration contains the gradient exchange); rank 0 reports
        # (single stream for the replay: with the pseudo-label pass overlapped on a second stream the per-launch events
        # would also time the other stream's kernels)
        un._overlap = False
        L.salun_profile_begin()
        for i in range(min(steps, 3)):
            it_resident(i)
        pm, pc, pf = (C.c_double * 2)(), (C.c_int64 * 2)(), (C.c_double * 2)()
        L.salun_profile_end(pm, pc, pf)
        un._overlap = True
        if rank == 0:
            k = min(steps, 3)
            ach = [pf[c] / (pm[c] * 1e-3) / 1e12 if pm[c] > 0 else 0.0 for c in range(2)]
            tr = (load_json("profiles/r2_traffic.json") or {}).get("unet_conv", {})
            roof = {"bound": "tensor", "kernel": "k_conv_gemm_p / k_gemm2 (conv forward + dgrad, attention and projection GEMMs)",
                    "achieved": ach[0], "peak": tf_peak, "unit": "TFLOP/s", "frac": ach[0] / tf_peak,
                    "traffic": tr.get("bytes"), "traffic_source": tr.get("source"),
                    "how": "CUDA events around every launch of the kernel category, 3 instrumented single-stream iterations "
                           "replayed after the timed region",
                    "peak_source": peak_src, "avg_launch_us": pm[0] * 1e3 / max(1, pc[0]), "launches_per_it": pc[0] / k,
                    "share_of_it": pm[0] / k / ms,
                    "other": {"kernel": "k_wgrad (side stream)", "achieved": ach[1], "share_of_it": pm[1] / k / ms}}
    strong = args.scaling == "strong"
    it_tflop = 7 * B * DDPM_FWD_GFLOP * 1e-3
    h2d = 2 * (B * 3 * 32 * 32 * 4 + B * 8)
    res = {"metric": "DDPM saliency_unlearn iterations/sec (cifar10 U-Net 32x32, "
                     + (f"global batch {B * world}+{B * world} sharded over the GPUs" if strong else f"{B} remain + {B} forget images per GPU") + ", rl)",
           "value": (1.0 if strong else world) * 1000.0 / ms, "unit": "iterations/s", "n_gpus": world, "steps": steps, "ms_per_it": ms,
           "tflops_per_gpu": it_tflop / ms * 1e3, "dtype": "bf16" if precision == "bf16" else "bf16x2-split (fp32-class)",
           "scaling": args.scaling,
           "config": {"workload": "DDPM U-Net CIFAR-10 32x32 saliency_unlearn iteration (runners/diffusion.py:519-593), "
                                  "method rl, alpha 1e-3, dropout 0.1, cond_drop 0.1, mask ratio 0.5, clip 1.0, Adam 1e-4",
                      "per_gpu_batch": [B, B], "params": eng.n, "precision": precision,
                      "streams": "pseudo-label pass on a second stream (forward-only engine replica), wgrad on a side stream",
                      "collective": ("none" if world == 1 else
                                     "fused reduce-scatter + global-norm clip + mask + Adam + all-gather over NVLink peer memory "
                                     "(two kernels around one barrier, optimizer state sharded)" if un.fused_dp else
                                     "NCCL all-reduce of the flat gradient (before the clip)")},
           "gpu_launches": int(launches), "launches_per_it": launches / steps,
           "e2e": {"value": (1.0 if strong else world) * 1000.0 / ms_e2e, "unit": "iterations/s", "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": 4, "ms_per_it": ms_e2e},
           "roofline": roof, "final_loss": last[0]}
    eng.close()
    del un, eng
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "split"], help="engine build of the timed step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ddpm", action="store_true", help="skip the second workload (DDPM U-Net iteration)")
    ap.add_argument("--no-maskgen", action="store_true")
    ap.add_argument("--no-modes", action="store_true", help="skip the split-precision rerun of both workloads")
    ap.add_argument("--no-torch", action="store_true", help="skip the stock-PyTorch eager denominator")
    ap.add_argument("--no-sd", action="store_true", help="skip the Stable-Diffusion v1.4 U-Net forward / DDIM sub-line")
    ap.add_argument("--ddpm-steps", type=int, default=20)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "torch":
        return run_torch(args)
    args.warmup = max(args.warmup, 3)

    import ctypes as C
    import torch
    import torch.distributed as dist
    from unlearn_saliency_b200 import _lib
    from unlearn_saliency_b200.engine import MaskedSGD, ResNetEngine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        # a mismatched collective must fail within minutes, not after the default 10-minute watchdog
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    L = _lib.lib(args.precision)
    strong = args.scaling == "strong" and world > 1
    per_gpu = BATCH // world if strong else BATCH
    if strong and BATCH % world:
        raise SystemExit("--scaling strong needs 256 % N == 0")
    step_gflop = 3 * FWD_GFLOP_PER_IMG * per_gpu

    fused_dp = world > 1 and os.environ.get("SALUN_FUSED_DP", "1") != "0"
    try:
        eng = ResNetEngine("resnet18", 10, 32, max_batch=per_gpu, device=dev, symmetric=fused_dp, precision=args.precision)
    except Exception as e:  # symmetric memory unavailable on this box: NCCL all-reduce + local step (same arithmetic)
        print(f"[bench] symmetric memory unavailable ({e!r}); using NCCL all-reduce", file=sys.stderr)
        fused_dp = False
        eng = ResNetEngine("resnet18", 10, 32, max_batch=per_gpu, device=dev, precision=args.precision)
    sd, g = rn18_state(eng, torch)
    eng.load_state_dict(sd)
    mask_native = (torch.rand(eng.n_params, generator=g) < 0.5).to(torch.int64).to(dev)
    bits = eng.ctx.pack_mask(mask_native)
    if fused_dp:
        from unlearn_saliency_b200.engine import DistMaskedSGD
        try:
            opt = DistMaskedSGD(eng, 0.013, 0.9, 5e-4, mask_bits=bits)
        except Exception as e:
            print(f"[bench] fused DP step unavailable ({e!r}); using NCCL all-reduce", file=sys.stderr)
            fused_dp = False
    if not fused_dp:
        opt = MaskedSGD(eng, 0.013, 0.9, 5e-4, mask_bits=bits)
    eng.train(True)

    # synthetic CIFAR-shaped inputs: a pool larger than L2 is not needed for the images (3 MB/step); the step itself
    # streams ~2 GB of activations + 45 MB of gradients + optimizer state, far beyond the 126 MB L2.
    gen = torch.Generator(device="cpu").manual_seed(100 + rank)
    n_pool = 8
    host_x = [torch.rand(per_gpu, 3, 32, 32, generator=gen).pin_memory() for _ in range(n_pool)]
    host_y = [torch.randint(0, 10, (per_gpu,), generator=gen).pin_memory() for _ in range(n_pool)]
    dev_x = [t.to(dev) for t in host_x]
    dev_y = [t.to(dev) for t in host_y]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def graphed(engine, optimizer, batch, lib):
        """(GraphedStep | None, launches per replay): the ~175 dependent launches of a step are launch-latency bound"""
        # N > 1: only with the fused DP optimizer (its barriers and exchange kernel are captured); NCCL all-reduce stays eager
        if (world != 1 and not fused_dp) or os.environ.get("SALUN_GRAPH", "1") == "0":
            return None, 0
        try:
            from unlearn_saliency_b200.engine import GraphedStep
            c0 = lib.salun_launch_count()
            gs = GraphedStep(engine, optimizer, batch)
            return gs, (lib.salun_launch_count() - c0) // 3  # two warm-up steps + the captured one
        except Exception as e:
            print(f"[bench] CUDA graph capture unavailable ({e!r}); eager launches", file=sys.stderr)
            return None, 0

    graph, graph_launches = graphed(eng, opt, per_gpu, L)

    def step_resident(i):
        if graph is not None:
            graph(dev_x[i % n_pool], dev_y[i % n_pool])
            return
        eng.forward_backward(dev_x[i % n_pool], dev_y[i % n_pool])
        if world > 1 and not fused_dp:
            dist.all_reduce(eng.grads)
            eng.grads.div_(world)
        opt.step()  # fused_dp: reduce-scatter + masked SGD + all-gather in one kernel over NVLink peer memory

    def step_e2e(i):
        if graph is not None:  # H2D copies into the graph's static buffers, replay, D2H read of the loss
            return float(graph(host_x[i % n_pool], host_y[i % n_pool]).item())
        x = host_x[i % n_pool].to(dev, non_blocking=True)
        y = host_y[i % n_pool].to(dev, non_blocking=True)
        loss, _ = eng.forward_backward(x, y)
        if world > 1 and not fused_dp:
            dist.all_reduce(eng.grads)
            eng.grads.div_(world)
        opt.step()
        return float(loss.item())  # D2H read of the step's result

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = L.salun_launch_count()
    ms_total = timed(step_resident, args.steps)
    launches = L.salun_launch_count() - launches0
    if graph is not None:
        launches = graph_launches * args.steps  # graph replays do not pass through the library's launch counter
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / args.steps
    value = (1.0 if strong else world) * 1000.0 / ms_step

    for i in range(3):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps) / args.steps
    e2e_value = (1.0 if strong else world) * 1000.0 / ms_e2e
    h2d = host_x[0].numel() * 4 + host_y[0].numel() * 8
    final_loss = float(eng._loss.item())

    # roofline of the dominant kernel (instrumented replay right after the timed region, same inputs)
    hbm_peak, tf_peak, tf_sustained, peak_src = load_peaks()
    roof = None
    if rank == 0:
        L.salun_profile_begin()
        prof_steps = min(args.steps, 10)
        for i in range(prof_steps):
            eng.forward_backward(dev_x[i % n_pool], dev_y[i % n_pool])
        ms = (C.c_double * 2)()
        cnt = (C.c_int64 * 2)()
        fl = (C.c_double * 2)()
        L.salun_profile_end(ms, cnt, fl)
        ach = [fl[c] / (ms[c] * 1e-3) / 1e12 if ms[c] > 0 else 0.0 for c in range(2)]
        dom = 0 if ms[0] >= ms[1] else 1
        tr = (load_json("profiles/r2_traffic.json") or {}).get("resnet_conv", {})
        roof = {
            "bound": "tensor", "kernel": ["k_conv_gemm_p / k_conv_rw (conv forward + dgrad)", "k_wgrad"][dom],
            "achieved": ach[dom], "peak": tf_peak, "unit": "TFLOP/s", "frac": ach[dom] / tf_peak,
            "frac_of_sustained_peak": ach[dom] / tf_sustained,
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the category's largest kernel, from the round's
            # ncu --set full capture (tools/ncu_traffic.py -> profiles/r2_traffic.json)
            "traffic": tr.get("bytes"), "traffic_source": tr.get("source"),
            "peak_source": peak_src,
            "how": f"CUDA events around every launch of the kernel, {prof_steps} instrumented steps replayed after the timed region",
            "avg_launch_us": ms[dom] * 1e3 / max(1, cnt[dom]), "launches_per_step": cnt[dom] / prof_steps,
            "share_of_step": ms[dom] / prof_steps / ms_step,
            "other": {"kernel": ["k_conv_gemm_p / k_conv_rw", "k_wgrad"][1 - dom], "achieved": ach[1 - dom],
                      "share_of_step": ms[1 - dom] / prof_steps / ms_step},
        }
    eng.close()
    del opt, graph
    torch.cuda.empty_cache()

    single = rank == 0 and world == 1
    modes = None
    if single and not args.no_modes and args.precision == "bf16" and "split" in _lib.available_precisions():
        # the same step in the split-precision build (fp32-class products: at least the reference's arithmetic class)
        try:
            Ls = _lib.lib("split")
            e2 = ResNetEngine("resnet18", 10, 32, max_batch=BATCH, device=dev, precision="split")
            e2.load_state_dict(sd)
            o2 = MaskedSGD(e2, 0.013, 0.9, 5e-4, mask_bits=bits)
            e2.train(True)
            g2, _ = graphed(e2, o2, BATCH, Ls)

            def step2(i):
                if g2 is not None:
                    g2(dev_x[i % n_pool], dev_y[i % n_pool])
                else:
                    e2.forward_backward(dev_x[i % n_pool], dev_y[i % n_pool])
                    o2.step()

            def step2_e2e(i):
                if g2 is not None:
                    return float(g2(host_x[i % n_pool], host_y[i % n_pool]).item())
                loss, _ = e2.forward_backward(host_x[i % n_pool].to(dev, non_blocking=True), host_y[i % n_pool].to(dev, non_blocking=True))
                o2.step()
                return float(loss.item())

            k2 = max(5, min(args.steps, 50))
            for i in range(5):
                step2(i)
            ms2 = timed(step2, k2) / k2
            for i in range(3):
                step2_e2e(i)
            ms2e = timed(step2_e2e, k2) / k2
            modes = {"split": {"metric": METRIC, "value": 1000.0 / ms2, "unit": "steps/s", "ms_per_step": ms2,
                               "e2e": {"value": 1000.0 / ms2e, "unit": "steps/s", "ms_per_step": ms2e},
                               "dtype": "bf16x2-split (fp32-class): activations as bf16 hi/lo pairs, 4 tensor-core partial "
                                        "products per multiply, fp32 accumulate", "steps": k2}}
            e2.close()
            del e2, o2, g2
            torch.cuda.empty_cache()
        except Exception as e:
            modes = {"split": {"error": repr(e)}}
            print(f"[bench] split-precision step failed: {e!r}", file=sys.stderr)

    maskgen = None
    if single and not args.no_maskgen:
        try:
            maskgen = bench_maskgen(dev, timed, hbm_peak)
        except Exception as e:
            maskgen = {"error": repr(e)}
            print(f"[bench] maskgen workload failed: {e!r}", file=sys.stderr)

    ddpm = None
    if not args.no_ddpm:
        try:
            dper = DDPM_BATCH // world if strong else DDPM_BATCH
            ddpm = bench_ddpm(args, dev, rank, world, L, timed, tf_peak, peak_src, precision=args.precision, per_gpu=dper)
            if single and modes is not None and args.precision == "bf16" and "split" in _lib.available_precisions():
                try:
                    dm = bench_ddpm(args, dev, rank, world, _lib.lib("split"), timed, tf_peak, peak_src, precision="split",
                                    per_gpu=dper, roofline=False)
                    modes["split"]["ddpm"] = {k: dm[k] for k in ("metric", "value", "unit", "ms_per_it", "dtype", "e2e")}
                except Exception as e:
                    modes["split"]["ddpm"] = {"error": repr(e)}
        except Exception as e:  # the headline line must still be printed
            ddpm = {"error": repr(e)}
            print(f"[bench] DDPM workload failed: {e!r}", file=sys.stderr)

    torch_eager = None
    if single and not args.no_torch:
        try:
            torch_eager = torch_eager_numbers(ddpm=not args.no_ddpm, steps=20, ddpm_steps=5)
            torch_eager["speedup_e2e_vs_tf32"] = {"resnet18": e2e_value / torch_eager["resnet18_tf32"]["value"]}
            if ddpm and "e2e" in ddpm and "ddpm_tf32" in torch_eager:
                torch_eager["speedup_e2e_vs_tf32"]["ddpm"] = ddpm["e2e"]["value"] / torch_eager["ddpm_tf32"]["value"]
        except Exception as e:
            torch_eager = {"error": repr(e)}
            print(f"[bench] torch eager arm failed: {e!r}", file=sys.stderr)

    sd_unet = None
    if single and not args.no_sd:
        # the 859.5 M-parameter SD U-Net forward on the engine and one 50-step guided DDIM
        # sample, beside the stock-PyTorch statement of the same network (tools/bench_sd_unet.py; random-init weights)
        try:
            torch.cuda.empty_cache()
            from tools.bench_sd_unet import measure as sd_measure
            sd_unet = sd_measure(dev, torch_fp32=False, iters=10)
            torch.cuda.empty_cache()
        except Exception as e:
            sd_unet = {"error": repr(e)}
            print(f"[bench] SD U-Net sub-line failed: {e!r}", file=sys.stderr)

    cpu = None
    if single and not args.no_cpu_baseline:
        cb = cpu_baselines(ddpm=bool(ddpm) and "error" not in (ddpm or {}))
        cpu = cb["resnet"]
        if ddpm and "ddpm" in cb:
            ddpm["cpu_baseline"] = cb["ddpm"]

    if rank == 0:
        parity = {"source": "tests/test_acceptance_gpu.py on a B200 (profiles/r2_acceptance_*.json); 'split' is the build the "
                            "generate_mask mirrors use, 'tf32_reference_gpu_path' the reference's own statements with torch's default TF32"}
        for nm in ("resnet18", "ddpm", "resnet18_weights"):
            a = load_json(f"profiles/r2_acceptance_{nm}.json")
            if a:
                parity[nm] = {k: a[k] for k in ("rel_l2_saliency", "update_rel_l2", "update_cos") if k in a}
                if "jaccard" in a:   # per build: the 50 % mask and the worst ratio of 0.1 .. 0.9
                    parity[nm]["jaccard"] = {b: {"ratio_0.5": j["0.5"], "min_over_ratios": min(j.values())} for b, j in a["jaccard"].items()}
        line = {
            "metric": METRIC_STRONG if strong else METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "bf16x2-split (fp32-class)", "data": "synthetic",
            "config": {"workload": "ResNet-18/CIFAR-10 SalUn RL masked unlearn step (RL.py:123-140), mask ratio 0.5",
                       "global_batch": per_gpu * world, "per_gpu_batch": per_gpu, "image": "3x32x32",
                       "parallelism": f"dp{world}", "precision": args.precision,
                       "collective": ("fused reduce-scatter + masked SGD + all-gather kernel over NVLink peer memory"
                                      if fused_dp else ("NCCL all-reduce of the flat gradient" if world > 1 else "none")),
                       "nvlink_bytes_per_step_per_gpu": (int(2 * (world - 1) / world * N_PARAMS_RN18 * 4) if world > 1 else 0),
                       "nvlink_note": ("algorithmic: each rank loads its 1/W gradient shard from the W-1 peers and stores its 1/W "
                                       "shard of the new weights into the W-1 peers (fp32 arena of 11.17 M parameters)"
                                       if fused_dp else None),
                       "optimizer": "SGD lr 0.013 momentum 0.9 wd 5e-4 (fused masked step)",
                       "cuda_graph": graph_launches > 0,
                       "l2": "step working set (~2 GB activations + 134 MB optimizer state) exceeds the 126 MB L2"},
            "tflops_per_gpu": step_gflop / ms_step, "gpu_launches": int(launches), "launches_per_step": launches / args.steps,
            "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e},
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "final_loss": final_loss,
            "ddpm": ddpm, "maskgen": maskgen, "modes": modes, "torch_eager": torch_eager, "parity": parity,
            "sd_unet": sd_unet,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
