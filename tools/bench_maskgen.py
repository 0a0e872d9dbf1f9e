"""Saliency-mask generation throughput (BASELINE.json configs[0] shape on the GPU): save_gradient_ratio over a synthetic
forget set through the drop-in mirror -- images/s of the eval-mode fwd+bwd+accumulate loop and seconds for the 10-ratio
select + int64 mask materialisation (+ torch.save)."""
import argparse, json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unlearn_saliency_b200.classification.generate_mask import accumulate_saliency, masks_for_ratio, THRESHOLD_LIST
from unlearn_saliency_b200.engine import ResNetEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4608
eng = ResNetEngine("resnet18", 10, 32, max_batch=256)
g = torch.Generator().manual_seed(0)
sd = {k: (torch.randn(s, generator=g) * 0.05 if len(s) != 1 else torch.ones(s)) for k, s in eng.table.items()}
eng.load_state_dict(sd)
x = torch.rand(n, 3, 32, 32, generator=g); y = torch.randint(0, 10, (n,), generator=g)
loader = [(x[i:i + 256].pin_memory(), y[i:i + 256].pin_memory()) for i in range(0, n, 256)]
accumulate_saliency(eng, loader[:2]); torch.cuda.synchronize()
t0 = time.perf_counter(); acc = accumulate_saliency(eng, loader); torch.cuda.synchronize(); t_acc = time.perf_counter() - t0
flat = eng.from_native_flat(acc).contiguous(); torch.cuda.synchronize()
t0 = time.perf_counter()
for r in THRESHOLD_LIST:
    hd, bits, info = masks_for_ratio(eng, flat, r)
torch.cuda.synchronize(); t_sel = time.perf_counter() - t0
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
for _ in range(10):
    eng.ctx.topk_mask(flat, int(flat.numel() * 0.5), want_info=False)
e1.record(); torch.cuda.synchronize(); sel_ms = e0.elapsed_time(e1) / 10
with tempfile.TemporaryDirectory() as d:
    t0 = time.perf_counter(); torch.save(hd, os.path.join(d, "with_1.0.pt")); t_save = time.perf_counter() - t0
nparam = flat.numel()
print(json.dumps({"metric": "saliency mask generation (ResNet-18, eval-mode -CE fwd+bwd+accumulate, H2D included)",
                  "images": n, "images_per_s": n / t_acc, "accumulate_s": t_acc,
                  "select_10_ratios_s": t_sel, "select_one_ratio_ms_device": sel_ms,
                  "select_GBps_algorithmic": (nparam * 24) / (sel_ms * 1e-3) / 1e9,
                  "torch_save_one_file_s": t_save, "params": nparam,
                  "reference_cpu_note": "reference save_gradient_ratio: 24.9 img/s and 1.67 s per ratio for the two argsorts on 8 CPU cores (SURVEY.md section 6)"}))
