"""One eager (no graph) full-size SD v1.4 U-Net forward between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_sd_launches.csv \\
      python tools/sd_launches.py [bf16|split]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.bench_sd_unet import synth                               # noqa: E402
from unlearn_saliency_b200.sd.engine import SDUNetEngine, sd_v1_config   # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
dev = torch.device("cuda:0")
cfg = sd_v1_config()
eng = SDUNetEngine(cfg, latent_size=64, max_batch=2, context_len=77, device=dev, precision=precision, use_graph=False)
eng.load_state_dict(synth(eng.table, dev))
x, t, c = torch.randn(2, 4, 64, 64, device=dev), torch.tensor([481.0, 37.0], device=dev), torch.randn(2, 77, 768, device=dev)
eng(x, t, c)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
eng(x, t, c)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
