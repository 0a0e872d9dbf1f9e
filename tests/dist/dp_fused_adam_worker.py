"""torchrun -N 2: the fused peer-memory DP clip + mask + Adam (DistMaskedAdam) against NCCL all-reduce + FlatMaskedAdam
(same arithmetic up to the summation order inside the norm), replicas bit-identical; then one DDPM iteration through
DDPMEngineUnlearner on a symmetric engine vs the NCCL path."""
import datetime, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
from types import SimpleNamespace
from unlearn_saliency_b200.diffusion.engine import DistMaskedAdam, UNetEngine
from unlearn_saliency_b200.diffusion.runner import DDPMEngineUnlearner, get_beta_schedule
from unlearn_saliency_b200.flat import FlatMaskedAdam

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))
cfg = SimpleNamespace(
    model=SimpleNamespace(type="conditional", in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2], num_res_blocks=1,
                          attn_resolutions=[8], dropout=0.0, resamp_with_conv=True, cond_drop_prob=0.1),
    data=SimpleNamespace(image_size=16, channels=3, n_classes=10),
    diffusion=SimpleNamespace(beta_schedule="linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000))
eng_f = UNetEngine(cfg, max_batch=8, device=dev, symmetric=True)
eng_n = UNetEngine(cfg, max_batch=8, device=dev, ctx=eng_f.ctx)
g = torch.Generator().manual_seed(0)
p0 = (0.05 * torch.randn(eng_f.n, generator=g)).to(dev)
mask_native = (torch.rand(eng_f.n, generator=g) < 0.5).to(torch.int64).to(dev)
bits = eng_f.ctx.pack_mask(mask_native)
eng_f.params.copy_(p0); eng_n.params.copy_(p0)
of = DistMaskedAdam(eng_f, lr=1e-4, max_norm=1.0); of.mask_bits = bits
on = FlatMaskedAdam(eng_n, lr=1e-4, max_norm=1.0); on.mask_bits = bits
ok = True
for step in range(4):
    gr = torch.Generator().manual_seed(100 * step + rank)  # a different gradient on every rank
    grad = (torch.randn(eng_f.n, generator=gr) * (1e-3 if step % 2 else 1e-5)).to(dev)   # clip active / inactive
    eng_f.grads.copy_(grad); eng_n.grads.copy_(grad)
    of.step()
    dist.all_reduce(eng_n.grads); eng_n.grads.div_(world); on.step()
    torch.cuda.synchronize()
    nf, nn_ = float(of.grad_norm()), float(on.grad_norm())
    ok &= abs(nf - nn_) <= 1e-5 * nn_
d = (eng_f.params - eng_n.params).abs().max().item()
upd = (eng_n.params - p0).abs().max().item()
same_masked = torch.equal(eng_f.params[mask_native == 0], p0[mask_native == 0])
gathered = [torch.empty_like(eng_f.params) for _ in range(world)]
dist.all_gather(gathered, eng_f.params)
identical = all(torch.equal(gathered[0], t) for t in gathered)
m1, m2 = of.gather_state()
dm = (m1 - on.exp_avg).abs().max().item()
print(f"rank {rank}: max |fused - nccl| = {d:.3e} (max update {upd:.3e}); moments {dm:.3e}; masked-out exact: {same_masked}; "
      f"replicas identical: {identical}; norms ok: {ok}", flush=True)
ok = ok and d <= 1e-3 * upd + 1e-9 and same_masked and identical and dm <= 1e-6
# end to end: one iteration through the runner (different data per rank), fused vs NCCL
betas = torch.from_numpy(get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)).float()
eng_f.params.copy_(p0); eng_n.params.copy_(p0)
un_f = DDPMEngineUnlearner(eng_f, betas, lr=1e-4, grad_clip=1.0); un_f.opt.mask_bits = bits
un_n = DDPMEngineUnlearner(eng_n, betas, lr=1e-4, grad_clip=1.0); un_n.opt.mask_bits = bits
assert un_f.fused_dp and not un_n.fused_dp
gd = torch.Generator().manual_seed(7 + rank)
n, S = 4, 16
r = dict(t_r=torch.randint(0, 1000, (n,), generator=gd), e_r=torch.randn(n, 3, S, S, generator=gd),
         t_f=torch.randint(0, 1000, (n,), generator=gd), e_f=torch.randn(n, 3, S, S, generator=gd),
         drop_r=torch.rand(n, generator=gd) < 0.1, drop_f=torch.rand(n, generator=gd) < 0.1, drop_p=torch.rand(n, generator=gd) < 0.1)
xr, cr = torch.rand(n, 3, S, S, generator=gd), torch.randint(1, 10, (n,), generator=gd)
xf, cf = torch.rand(n, 3, S, S, generator=gd), torch.zeros(n, dtype=torch.long)
lf = un_f.saliency_unlearn_step(xr, cr, xf, cf, rng=r)
ln = un_n.saliency_unlearn_step(xr, cr, xf, cf, rng=r)
torch.cuda.synchronize()
d2 = (eng_f.params - eng_n.params).abs().max().item()
u2 = (eng_n.params - p0).abs().max().item()
print(f"rank {rank}: runner iteration: loss {float(lf):.6f} / {float(ln):.6f}; max |fused - nccl| = {d2:.3e} (max update {u2:.3e})", flush=True)
ok = ok and abs(float(lf) - float(ln)) < 1e-6 and d2 <= 2e-2 * u2 + 1e-9
# timing on the full-size arena
def timeit(fn, k=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / k * 1e3
def nccl_step():
    dist.all_reduce(eng_n.grads); eng_n.grads.div_(world); on.step()
t_f, t_n = timeit(of.step), timeit(nccl_step)
if rank == 0:
    print(f"fused DP clip+mask+Adam {t_f:.1f} us   vs   NCCL all-reduce + scale + sumsq + clip + masked Adam {t_n:.1f} us  "
          f"(world {world}, {eng_f.n * 4 / 1e6:.1f} MB arena)", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
