"""torchrun -N 2: a mini-batch sharded over two ranks with sync-BN + the fused data-parallel masked SGD step must compute
the single-process step on the concatenated batch (the reference's semantics: one nn.BatchNorm2d over the whole
mini-batch, RL.py:123-140).  Checked on rank-local engines: engine A = sharded (sync-BN, DistMaskedSGD), engine B = the
full batch on one GPU (MaskedSGD).  Without sync-BN the same sharded run is shown to differ (per-shard statistics)."""
import datetime
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist

from oracle import classification as OC
from unlearn_saliency_b200.engine import DistMaskedSGD, MaskedSGD, ResNetEngine

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
prec = os.environ.get("SALUN_TEST_PRECISION", "split")
params, buffers = OC.synth_state(10, seed=0)
sd = OC.state_dict_of(params, buffers)
g = torch.Generator().manual_seed(5)
N = 64
x = torch.rand(3, N, 3, 32, 32, generator=g)
y = torch.randint(0, 10, (3, N), generator=g)
mask = (torch.rand(11173962, generator=g) < 0.5).to(torch.int64).to(dev)


def run(sync_bn: bool, steps: int = 3):
    eng = ResNetEngine("resnet18", 10, 32, max_batch=N, device=dev, symmetric=True, precision=prec)
    eng.load_state_dict(sd)
    if sync_bn:
        eng.enable_sync_bn()
    opt = DistMaskedSGD(eng, 0.013, 0.9, 5e-4, mask_bits=eng.ctx.pack_mask(eng.to_native(mask).contiguous()))
    eng.train(True)
    per = N // world
    for s in range(steps):
        eng.forward_backward(x[s, rank * per:(rank + 1) * per].to(dev).contiguous(), y[s, rank * per:(rank + 1) * per].to(dev).contiguous())
        opt.step()
    torch.cuda.synchronize()
    return eng


def reference(steps: int):
    ref = ResNetEngine("resnet18", 10, 32, max_batch=N, device=dev, precision=prec)
    ref.load_state_dict(sd)
    ropt = MaskedSGD(ref, 0.013, 0.9, 5e-4, mask_bits=ref.ctx.pack_mask(ref.to_native(mask).contiguous()))
    ref.train(True)
    for s in range(steps):
        ref.forward_backward(x[s].to(dev).contiguous(), y[s].to(dev).contiguous())
        ropt.step()
    torch.cuda.synchronize()
    return ref


ref1, ref = reference(1), reference(3)
p0 = ref.to_native({k: v for k, v in params.items()}).to(dev)


def rel(e, r):
    return float(((e.params - p0) - (r.params - p0)).norm() / (r.params - p0).norm())


# one step: the only differences are summation orders (statistics in fp64, split-K partitions of the weight gradient)
a1, b1 = run(True, 1), run(False, 1)
ra1, rb1 = rel(a1, ref1), rel(b1, ref1)
# three steps: train-mode BatchNorm at random init amplifies those rounding differences from step to step
a = run(True)
b = run(False)
ra, rb = rel(a, ref), rel(b, ref)
rm = float((a.running_mean - ref.running_mean).abs().max())
rv = float((a.running_var - ref.running_var).abs().max() / ref.running_var.abs().max())
gathered = [torch.empty_like(a.params) for _ in range(world)]
dist.all_gather(gathered, a.params)
identical = all(torch.equal(gathered[0], t) for t in gathered)
print(f"rank {rank} [{prec}]: sharded step vs full-batch step, relative error of the weight update: 1 step sync-BN {ra1:.3e} | "
      f"per-shard BN {rb1:.3e}; 3 steps sync-BN {ra:.3e} | per-shard BN {rb:.3e}; running_mean max diff {rm:.3e}, "
      f"running_var rel diff {rv:.3e}; replicas identical: {identical}", flush=True)
ok = ra1 < 2e-3 and rb1 > 10 * ra1 and ra < 5e-2 and rb > 5 * ra and rm < 1e-4 and rv < 1e-4 and identical
dist.destroy_process_group()
sys.exit(0 if ok else 1)
