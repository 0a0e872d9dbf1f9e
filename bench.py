#!/usr/bin/env python
"""bench.py -- SalUn masked-unlearning steps/sec (ResNet-18 / CIFAR-10 shape) on N B200s.

    python bench.py --gpus 1 --steps 200 --warmup 20
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

Workload (BASELINE.json configs[1]): one step = one mini-batch of the SalUn random-label
unlearning loop (Classification/unlearn/RL.py:123-140): train-mode ResNet-18 forward + backward on 256 synthetic
32x32 images, mask (.) grad, SGD(momentum 0.9, wd 5e-4, lr 0.013), restore -- with a 50% saliency mask.
N > 1: weak scaling, 256 images per GPU and ONE all-reduce of the gradient per step (global batch 256*N).

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
# keep stdout to the ONE JSON line: NCCL prints its version banner to stdout at NCCL_DEBUG=VERSION/WARN/INFO
os.environ.pop("NCCL_DEBUG", None)
if os.environ.get("SALUN_NCCL_DEBUG"):
    os.environ["NCCL_DEBUG"] = os.environ["SALUN_NCCL_DEBUG"]

BATCH = 256
FWD_GFLOP_PER_IMG = 1.1108  # SURVEY.md section 6 (torch.utils.flop_counter on the reference model)
STEP_GFLOP = 3 * FWD_GFLOP_PER_IMG * BATCH  # 853.1 GFLOP / step / GPU
METRIC = "unlearn steps/sec (ResNet-18 CIFAR-10 SalUn RL masked step, batch 256 per GPU)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json, sustained)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


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


def oracle_cpu_steps_per_sec(batch, steps, warmup=0):
    """The reference's CPU path for the same step: oracle/classification.py (torch fp32, all host threads).
    Returns (seconds per step at `batch`, threads)."""
    import torch
    from oracle import classification as OC
    torch.set_num_threads(os.cpu_count())
    params, buffers = OC.synth_state(10, seed=0)
    g = torch.Generator().manual_seed(1)
    flat_mask = (torch.rand(11173962, generator=g) < 0.5).to(torch.int64)
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
    return sum(times) / len(times), torch.get_num_threads()


def oracle_cpu_ddpm_sec_per_it(sample, steps=2, warmup=1):
    """The reference's CPU path for one DDPM saliency_unlearn iteration (oracle/ddpm.py: the statements of
    runners/diffusion.py:519-593 around the pinned U-Net restatement, torch fp32, all host threads) on `sample` remain +
    `sample` forget images.  Returns (seconds per iteration at that size, threads)."""
    import torch
    from oracle import ddpm as OD
    from unlearn_saliency_b200.diffusion.runner import get_beta_schedule
    from oracle.unet import ConditionalUNet, cifar10_config
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    model = ConditionalUNet(cifar10_config())
    g = torch.Generator().manual_seed(1)
    mask = {k: (torch.rand(p.shape, generator=g) < 0.5).to(torch.int64) for k, p in model.named_parameters()}
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    betas = torch.from_numpy(get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)).float()
    n, S = sample, 32
    times = []
    for s in range(warmup + steps):
        r = dict(x_r=torch.rand(n, 3, S, S, generator=g), c_r=torch.randint(1, 10, (n,), generator=g),
                 x_f=torch.rand(n, 3, S, S, generator=g), c_f=torch.zeros(n, dtype=torch.long),
                 t_r=torch.randint(0, 1000, (n,), generator=g), e_r=torch.randn(n, 3, S, S, generator=g),
                 t_f=torch.randint(0, 1000, (n,), generator=g), e_f=torch.randn(n, 3, S, S, generator=g),
                 drop_r=torch.rand(n, generator=g) < 0.1, drop_f=torch.rand(n, generator=g) < 0.1,
                 drop_p=torch.rand(n, generator=g) < 0.1)
        t0 = time.perf_counter()
        OD.saliency_unlearn_step(model, opt, mask, r, betas, alpha=1e-3, method="rl")
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times), torch.get_num_threads()


DDPM_BATCH = 128
DDPM_IT_TFLOP = 7 * DDPM_BATCH * 12.449e-3  # 7 forward-equivalents x 128 images x 12.449 GFLOP


def bench_ddpm(args, dev, rank, world, L, timed, tf_peak, peak_src):
    """Second headline workload (BASELINE.json configs[2]): one DDPM saliency_unlearn iteration (runners/diffusion.py:519-593)
    on the cifar10 U-Net, 128 remain + 128 forget images per GPU, method rl, dropout 0.1, 50% mask, clip 1.0, Adam.
    Weak scaling like the ResNet line; N > 1 adds one NCCL all-reduce of the 154 MB gradient arena per iteration."""
    import ctypes as C
    import torch
    from unlearn_saliency_b200.diffusion.engine import UNetEngine
    from unlearn_saliency_b200.diffusion.runner import DDPMEngineUnlearner, get_beta_schedule
    from unlearn_saliency_b200.diffusion.config import cifar10_config
    cfg = cifar10_config()
    fused_dp = world > 1 and os.environ.get("SALUN_FUSED_DP", "1") != "0"
    try:
        eng = UNetEngine(cfg, max_batch=2 * DDPM_BATCH, device=dev, symmetric=fused_dp)
    except Exception as e:  # symmetric memory unavailable: NCCL all-reduce + local clip / mask / Adam (same arithmetic)
        print(f"[bench] symmetric memory unavailable for the DDPM arenas ({e!r}); using NCCL all-reduce", file=sys.stderr)
        fused_dp = False
        eng = UNetEngine(cfg, max_batch=2 * DDPM_BATCH, device=dev)
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
    B = DDPM_BATCH
    host = [(torch.rand(B, 3, 32, 32, generator=gen).pin_memory(), torch.randint(1, 10, (B,), generator=gen).pin_memory(),
             torch.rand(B, 3, 32, 32, generator=gen).pin_memory(), torch.zeros(B, dtype=torch.long).pin_memory())
            for _ in range(4)]
    resident = [tuple(t.to(dev) for t in h) for h in host]

    def it_resident(i):
        xr, cr, xf, cf = resident[i % 4]
        un.saliency_unlearn_step(xr, cr, xf, cf, alpha=1e-3, method="rl")

    last = [0.0]

    def it_e2e(i):
        xr, cr, xf, cf = host[i % 4]  # pinned host tensors: the step copies them to the device (H2D inside the timed region)
        last[0] = float(un.saliency_unlearn_step(xr, cr, xf, cf, alpha=1e-3, method="rl").item())  # D2H of the loss

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
    # instrumented replay on EVERY rank (the iteration contains the gradient all-reduce); rank 0 reports
    # (single stream for the replay: with the pseudo-label pass overlapped on a second stream the per-launch events would
    # also time the other stream's kernels)
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
        roof = {"bound": "tensor", "kernel": "k_conv_gemm_p / k_gemm2 (conv forward + dgrad, attention and projection GEMMs)",
                "achieved": ach[0], "peak": tf_peak, "unit": "TFLOP/s", "frac": ach[0] / tf_peak,
                # one launch of the 128->128 3x3 convolution at 32x32, batch 256 (profiles/r1_ncu_full_unet_conv128.csv):
                # 76.1 MB read = the padded activation once, 23.6 MB written before the kernel retired (output 67 MB)
                "traffic": 99716352, "traffic_unit": "bytes per launch (ncu --set full, profiles/r1_ncu_full_unet_conv128.csv)",
                "how": "CUDA events around every launch of the kernel category, 3 instrumented single-stream iterations "
                       "replayed after the timed region",
                "peak_source": peak_src, "avg_launch_us": pm[0] * 1e3 / max(1, pc[0]), "launches_per_it": pc[0] / k,
                "share_of_it": pm[0] / k / ms,
                "other": {"kernel": "k_wgrad (side stream)", "achieved": ach[1], "share_of_it": pm[1] / k / ms}}
    h2d = 2 * (B * 3 * 32 * 32 * 4 + B * 8)
    res = {"metric": "DDPM saliency_unlearn iterations/sec (cifar10 U-Net 32x32, 128 remain + 128 forget images per GPU, rl)",
           "value": world * 1000.0 / ms, "unit": "iterations/s", "n_gpus": world, "steps": steps, "ms_per_it": ms,
           "tflops_per_gpu": DDPM_IT_TFLOP / ms * 1e3, "dtype": "bf16", "scaling": "weak",
           "config": {"workload": "DDPM U-Net CIFAR-10 32x32 saliency_unlearn iteration (runners/diffusion.py:519-593), "
                                  "method rl, alpha 1e-3, dropout 0.1, cond_drop 0.1, mask ratio 0.5, clip 1.0, Adam 1e-4",
                      "per_gpu_batch": [B, B], "params": eng.n,
                      "streams": "pseudo-label pass on a second stream (forward-only engine replica), wgrad on a side stream",
                      "collective": ("none" if world == 1 else
                                     "fused reduce-scatter + global-norm clip + mask + Adam + all-gather over NVLink peer memory "
                                     "(two kernels around one barrier, optimizer state sharded)" if un.fused_dp else
                                     "NCCL all-reduce of the flat gradient (before the clip)")},
           "gpu_launches": int(launches), "launches_per_it": launches / steps,
           "e2e": {"value": world * 1000.0 / ms_e2e, "unit": "iterations/s", "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": 4, "ms_per_it": ms_e2e},
           "roofline": roof, "final_loss": last[0]}
    eng.close()
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_batch = int(os.environ.get("SALUN_REF_BATCH", "32"))
    sec, threads = oracle_cpu_steps_per_sec(sample_batch, max(1, args.steps), min(args.warmup, 1))
    sec256 = sec * BATCH / sample_batch  # per-image cost is flat in the batch size on the CPU
    val = args.gpus / sec256  # same weak-scaling unit as our arm: 256-image steps per second (x N replicas' worth of work)
    val = 1.0 / sec256
    sample = f"{args.steps} RL steps of {sample_batch} images (oracle/classification.py, torch fp32 CPU), scaled x{BATCH // sample_batch} to 256-image steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec256 * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ResNet-18/CIFAR-10 SalUn RL masked unlearn step, batch 256, mask ratio 0.5", "global_batch": BATCH},
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_ddpm:
        sn = int(os.environ.get("SALUN_REF_DDPM_BATCH", "2"))
        dsec, dthreads = oracle_cpu_ddpm_sec_per_it(sn, steps=2, warmup=1)
        dval = 1.0 / (dsec * DDPM_BATCH / sn)
        line["ddpm"] = {"metric": "DDPM saliency_unlearn iterations/sec (cifar10 U-Net 32x32, 128 remain + 128 forget images, rl)",
                        "value": dval, "unit": "iterations/s", "dtype": "f32",
                        "cpu_baseline": {"value": dval, "unit": "iterations/s", "cores": dthreads, "kind": "port",
                                         "sample": f"2 iterations of {sn}+{sn} images after 1 warm-up (oracle/ddpm.py, torch fp32 CPU), "
                                                   f"scaled x{DDPM_BATCH // sn} to 128+128 images"}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ddpm", action="store_true", help="skip the second workload (DDPM U-Net iteration)")
    ap.add_argument("--ddpm-steps", type=int, default=20)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from unlearn_saliency_b200 import _lib
    from unlearn_saliency_b200.engine import MaskedSGD, ResNetEngine
    import ctypes as C

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
    L = _lib.lib()

    fused_dp = world > 1 and os.environ.get("SALUN_FUSED_DP", "1") != "0"
    try:
        eng = ResNetEngine("resnet18", 10, 32, max_batch=BATCH, device=dev, symmetric=fused_dp)
    except Exception as e:  # symmetric memory unavailable on this box: NCCL all-reduce + local step (same arithmetic)
        print(f"[bench] symmetric memory unavailable ({e!r}); using NCCL all-reduce", file=sys.stderr)
        fused_dp = False
        eng = ResNetEngine("resnet18", 10, 32, max_batch=BATCH, device=dev)
    # random-init weights of the reference architecture (kaiming-normal fan_out convs, unit BN), same on every rank
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
    host_x = [torch.rand(BATCH, 3, 32, 32, generator=gen).pin_memory() for _ in range(n_pool)]
    host_y = [torch.randint(0, 10, (BATCH,), generator=gen).pin_memory() for _ in range(n_pool)]
    dev_x = [t.to(dev) for t in host_x]
    dev_y = [t.to(dev) for t in host_y]

    # The ~175 dependent launches of a step are launch-latency bound: capture the step once into a CUDA graph
    # (engine.GraphedStep) and replay it.  N > 1 stays eager (the fused DP kernel synchronises ranks through peer flags).
    graph, graph_launches = None, 0
    if world == 1 and os.environ.get("SALUN_GRAPH", "1") != "0":
        try:
            from unlearn_saliency_b200.engine import GraphedStep
            c0 = L.salun_launch_count()
            graph = GraphedStep(eng, opt, BATCH)
            graph_launches = (L.salun_launch_count() - c0) // 3  # two warm-up steps + the captured one
        except Exception as e:
            print(f"[bench] CUDA graph capture unavailable ({e!r}); eager launches", file=sys.stderr)
            graph = None

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
    value = world * 1000.0 / ms_step

    for i in range(3):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps) / args.steps
    e2e_value = world * 1000.0 / ms_e2e
    h2d = host_x[0].numel() * 4 + host_y[0].numel() * 8
    final_loss = float(eng._loss.item())

    # roofline of the dominant kernel (instrumented replay right after the timed region, same inputs)
    hbm_peak, tf_peak, peak_src = load_peaks()
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
        roof = {
            "bound": "tensor", "kernel": ["k_conv_gemm (conv forward + dgrad)", "k_wgrad"][dom],
            "achieved": ach[dom], "peak": tf_peak, "unit": "TFLOP/s", "frac": ach[dom] / tf_peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the category's largest kernel (k_conv_rw<32,1>,
            # the 64->64 3x3 convolution at 32x32, batch 256) from profiles/r1_ncu_full_rw.csv: 38.0 MB read = the padded
            # activation once (no re-reads) + 1.3 MB written before the kernel retired
            "traffic": 39258624, "traffic_unit": "bytes per launch (ncu --set full, profiles/r1_ncu_full_rw.csv)",
            "peak_source": peak_src,
            "how": f"CUDA events around every launch of the kernel, {prof_steps} instrumented steps replayed after the timed region",
            "avg_launch_us": ms[dom] * 1e3 / max(1, cnt[dom]), "launches_per_step": cnt[dom] / prof_steps,
            "share_of_step": ms[dom] / prof_steps / ms_step,
            "other": {"kernel": ["k_conv_gemm", "k_wgrad"][1 - dom], "achieved": ach[1 - dom],
                      "share_of_step": ms[1 - dom] / prof_steps / ms_step},
        }
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sb = int(os.environ.get("SALUN_REF_BATCH", "32"))
        sec, threads = oracle_cpu_steps_per_sec(sb, 4, 1)
        cpu = {"value": 1.0 / (sec * BATCH / sb), "unit": "steps/s", "cores": threads, "kind": "port",
               "sample": f"4 RL steps of {sb} images after 1 warm-up (oracle/classification.py, torch fp32), scaled x{BATCH // sb}"}

    ddpm = None
    if not args.no_ddpm:
        try:
            eng.close()
            del opt
            torch.cuda.empty_cache()
            ddpm = bench_ddpm(args, dev, rank, world, L, timed, tf_peak, peak_src)
            if rank == 0 and world == 1 and not args.no_cpu_baseline:
                sn = int(os.environ.get("SALUN_REF_DDPM_BATCH", "2"))
                dsec, dthreads = oracle_cpu_ddpm_sec_per_it(sn, steps=1, warmup=1)
                ddpm["cpu_baseline"] = {"value": 1.0 / (dsec * DDPM_BATCH / sn), "unit": "iterations/s", "cores": dthreads,
                                        "kind": "port", "sample": f"1 iteration of {sn}+{sn} images after 1 warm-up "
                                        f"(oracle/ddpm.py, torch fp32), scaled x{DDPM_BATCH // sn}"}
        except Exception as e:  # the headline line must still be printed
            ddpm = {"error": repr(e)}
            print(f"[bench] DDPM workload failed: {e!r}", file=sys.stderr)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "ResNet-18/CIFAR-10 SalUn RL masked unlearn step (RL.py:123-140), mask ratio 0.5",
                       "global_batch": BATCH * world, "per_gpu_batch": BATCH, "image": "3x32x32",
                       "parallelism": f"dp{world}",
                       "collective": ("fused reduce-scatter + masked SGD + all-gather kernel over NVLink peer memory"
                                      if fused_dp else ("NCCL all-reduce of the flat gradient" if world > 1 else "none")),
                       "optimizer": "SGD lr 0.013 momentum 0.9 wd 5e-4 (fused masked step)",
                       "cuda_graph": graph is not None,
                       "l2": "step working set (~2 GB activations + 134 MB optimizer state) exceeds the 126 MB L2"},
            "tflops_per_gpu": STEP_GFLOP / ms_step, "gpu_launches": int(launches), "launches_per_step": launches / args.steps,
            "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e},
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "final_loss": final_loss,
            "ddpm": ddpm,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
