"""ResNet-50 / ImageNet-shape masked GA step (BASELINE.json configs[3] architecture): steps/s at a given per-GPU batch.
Single process: one B200.  Under torchrun (N ranks): data parallel, weak scaling (batch per GPU fixed), the gradient exchange +
masked SGD + weight broadcast as the fused kernel over NVLink peer memory (DistMaskedSGD); rank 0 prints one JSON line.
  python tools/bench_resnet50.py [batch] [steps]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_resnet50.py 256 10"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unlearn_saliency_b200.engine import DistMaskedSGD, MaskedSGD, ResNetEngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line
    dist.init_process_group("nccl", device_id=dev)
eng = ResNetEngine("resnet50", 1000, 224, max_batch=B, imagenet=True, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225),
                   device=dev, symmetric=world > 1)
g = torch.Generator().manual_seed(0)
sd = {}
for k, shp in eng.table.items():
    if len(shp) == 4: sd[k] = torch.randn(shp, generator=g) * (2.0 / (shp[0] * shp[2] * shp[3])) ** 0.5
    elif k == "fc.weight": sd[k] = torch.randn(shp, generator=g) * 0.01
    elif k.endswith(".weight"): sd[k] = torch.ones(shp)
    else: sd[k] = torch.zeros(shp)
sd["normalize.mean"] = torch.tensor(eng.mean); sd["normalize.std"] = torch.tensor(eng.std)
eng.load_state_dict(sd)
bits = eng.ctx.pack_mask((torch.rand(eng.n_params, generator=g) < 0.5).to(torch.int64).to(dev))
opt = (DistMaskedSGD if world > 1 else MaskedSGD)(eng, 0.01, 0.9, 5e-4, bits)
gr = torch.Generator().manual_seed(1 + rank)
x = torch.rand(B, 3, 224, 224, generator=gr).to(dev); y = torch.randint(0, 1000, (B,), generator=gr).to(dev)
eng.train(True)
for _ in range(3): eng.forward_backward(x, y, loss_sign=-1.0); opt.step()   # GA step (GA.py:107-128)
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
for _ in range(steps): eng.forward_backward(x, y, loss_sign=-1.0); opt.step()
e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / steps
if world > 1:
    t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
gflop = 3 * 8.1784 * B
if rank == 0:
    print(json.dumps({"metric": "ResNet-50/ImageNet-shape masked GA step (224x224, 1000 classes)", "n_gpus": world, "batch_per_gpu": B,
                      "ms_per_step": ms, "steps_per_s": 1000 / ms, "images_per_s": world * B * 1000 / ms,
                      "tflops_per_gpu": gflop / ms, "loss": float(eng._loss.item()),
                      "collective": "fused reduce-scatter + masked SGD + all-gather over NVLink peer memory (102 MB fp32 arena)" if world > 1 else "none",
                      "params": eng.n_params}), flush=True)
if world > 1:
    dist.destroy_process_group()
