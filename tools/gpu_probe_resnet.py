"""prints error statistics of the ResNet engine against the fp32 oracle and step timings (B200 box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import classification as OC
from unlearn_saliency_b200.engine import ResNetEngine, MaskedSGD

torch.set_num_threads(os.cpu_count())
eng = ResNetEngine("resnet18", 10, 32, max_batch=256)
params, buffers = OC.synth_state(10, seed=0)
eng.load_state_dict(OC.state_dict_of(params, buffers))
g = torch.Generator().manual_seed(11)
x = torch.rand(32, 3, 32, 32, generator=g); y = torch.randint(0, 10, (32,), generator=g)
for train, sign in ((False, -1.0), (True, 1.0)):
    b = {k: v.clone() for k, v in buffers.items()}
    lr, lg, gr = OC.loss_and_grads(params, b, x, y, train=train, sign=sign, emulate_bf16=True)
    eng.train(train)
    loss, logits = eng.forward_backward(x.cuda(), y.cuda(), loss_sign=sign, want_logits=True)
    torch.cuda.synchronize()
    print(f"train={train} loss {loss.item():.5f} ref {lr.item():.5f} logit maxerr {(logits.cpu()-lg).abs().max().item():.4f} (max |logit| {lg.abs().max().item():.2f})", flush=True)
    gd = eng.grad_dict()
    for k, r in gr.items():
        e = gd[k].cpu()
        rel = float((e - r).norm() / (r.norm() + 1e-12)); cos = float(torch.dot(e.flatten(), r.flatten()) / (e.norm() * r.norm() + 1e-20))
        if rel > 0.01 or cos < 0.9995 or k in ("conv1.weight", "fc.weight", "layer2.0.downsample.0.weight", "layer4.1.conv2.weight", "bn1.weight"):
            print(f"   {k:34s} rel {rel:.4f} cos {cos:.6f} |ref| {r.norm().item():.4e}", flush=True)

# timing at the benchmark shape
eng.load_state_dict(OC.state_dict_of(params, buffers))
xb = torch.rand(256, 3, 32, 32, device="cuda"); yb = torch.randint(0, 10, (256,), device="cuda")
bits = eng.ctx.pack_mask((torch.rand(eng.n_params, device="cuda") < 0.5).to(torch.int64))
opt = MaskedSGD(eng, 0.013, 0.9, 5e-4, bits)
eng.train(True)
for _ in range(5):
    eng.forward_backward(xb, yb); opt.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
N = 50
for _ in range(N):
    eng.forward_backward(xb, yb); opt.step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / N
print(f"engine step bs256: {ms:.3f} ms  -> {1000/ms:.1f} steps/s  ({853.1/ms:.1f} TFLOP/s)  loss {eng._loss.item():.4f}", flush=True)
t0 = time.time()
for _ in range(N):
    eng.forward_backward(xb, yb); opt.step()
torch.cuda.synchronize()
print(f"wall {1000*(time.time()-t0)/N:.3f} ms/step", flush=True)
