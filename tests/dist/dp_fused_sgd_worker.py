"""torchrun -N 2: the fused peer-memory DP step against NCCL all-reduce + local masked SGD (same arithmetic up to the
summation order of the all-reduce), replicas bit-identical."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["NCCL_DEBUG"] = "WARN"
import torch, torch.distributed as dist
from unlearn_saliency_b200.engine import DistMaskedSGD, MaskedSGD, ResNetEngine

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng_f = ResNetEngine("resnet18", 10, 32, max_batch=8, device=dev, symmetric=True)
eng_n = ResNetEngine("resnet18", 10, 32, max_batch=8, device=dev, ctx=eng_f.ctx)
g = torch.Generator().manual_seed(0)
p0 = torch.randn(eng_f.n_params, generator=g).to(dev)
mask = (torch.rand(eng_f.n_params, generator=g) < 0.5).to(torch.int64).to(dev)
bits = eng_f.ctx.pack_mask(mask)
eng_f.params.copy_(p0); eng_n.params.copy_(p0)
of = DistMaskedSGD(eng_f, 0.013, 0.9, 5e-4, mask_bits=bits)
on = MaskedSGD(eng_n, 0.013, 0.9, 5e-4, mask_bits=bits)
for step in range(4):
    gr = torch.Generator().manual_seed(100 * step + rank)  # different gradient on every rank
    grad = torch.randn(eng_f.n_params, generator=gr).to(dev)
    eng_f.grads.copy_(grad); eng_n.grads.copy_(grad)
    of.step()
    dist.all_reduce(eng_n.grads); eng_n.grads.div_(world); on.step()
torch.cuda.synchronize()
d = (eng_f.params - eng_n.params).abs().max().item()
upd = (eng_n.params - p0).abs().max().item()
same_masked = torch.equal(eng_f.params[mask == 0], p0[mask == 0])
gathered = [torch.empty_like(eng_f.params) for _ in range(world)]
dist.all_gather(gathered, eng_f.params)
identical = all(torch.equal(gathered[0], t) for t in gathered)
print(f"rank {rank}: max |fused - nccl| = {d:.3e} (max update {upd:.3e}); masked-out exact: {same_masked}; replicas identical: {identical}", flush=True)
ok = d <= 2e-6 * max(1.0, upd) + 1e-6 and same_masked and identical
# timing of the two variants
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
def nccl_step():
    dist.all_reduce(eng_n.grads); eng_n.grads.div_(world); on.step()
t_f, t_n = timeit(of.step), timeit(nccl_step)
if rank == 0:
    print(f"fused DP step {t_f:.1f} us   vs   NCCL all-reduce + scale + masked SGD {t_n:.1f} us  (world {world}, 44.7 MB arena)", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
