"""torchrun -N 2: the data-parallel step (forward_backward + DistMaskedSGD: barrier, fused reduce-scatter + masked SGD +
all-gather kernel over NVLink peer memory, barrier) replayed from a CUDA graph equals the eager sequence bit for bit, and the
replicas stay identical."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["NCCL_DEBUG"] = "WARN"
import torch, torch.distributed as dist
from oracle import classification as OC
from unlearn_saliency_b200.engine import DistMaskedSGD, GraphedStep, ResNetEngine

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 16
params, buffers = OC.synth_state(10, seed=0)
sd = OC.state_dict_of(params, buffers)
g = torch.Generator().manual_seed(7)
mask = (torch.rand(sum(v.numel() for v in params.values()), generator=g) < 0.5).to(torch.int64).to(dev)
gr = torch.Generator().manual_seed(50 + rank)          # every rank its own shard of the mini-batches
xs = [torch.rand(n, 3, 32, 32, generator=gr).to(dev) for _ in range(3)]
ys = [torch.randint(0, 10, (n,), generator=gr).to(dev) for _ in range(3)]
res, keep = [], []
for graphed in (False, True):
    eng = ResNetEngine("resnet18", 10, 32, max_batch=n, device=dev, symmetric=True)
    keep.append(eng)
    eng.load_state_dict(sd)
    eng.train(True)
    opt = DistMaskedSGD(eng, 0.013, 0.9, 5e-4, mask_bits=eng.ctx.pack_mask(mask))
    gs = GraphedStep(eng, opt, n) if graphed else None
    losses = []
    for x, y in zip(xs, ys):
        if gs is not None:
            losses.append(float(gs(x, y).item()))
        else:
            loss, _ = eng.forward_backward(x, y)
            opt.step()
            losses.append(float(loss.item()))
    torch.cuda.synchronize()
    res.append((eng.params.clone(), eng.running_mean.clone(), opt.momentum_shard.clone(), losses, eng.num_batches_tracked))
    dist.barrier()
same = all(torch.equal(a, b) for a, b in zip(res[0][:3], res[1][:3])) and res[0][3] == res[1][3] and res[0][4] == res[1][4]
gathered = [torch.empty_like(res[1][0]) for _ in range(world)]
dist.all_gather(gathered, res[1][0])
identical = all(torch.equal(gathered[0], t) for t in gathered)
print(f"rank {rank}: graphed DP step equals eager: {same}; replicas identical: {identical}; losses {res[1][3]}", flush=True)
dist.destroy_process_group()
sys.exit(0 if same and identical else 1)
