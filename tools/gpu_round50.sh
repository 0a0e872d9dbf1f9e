#!/bin/bash
# 8-GPU: bench.py under torchrun, bounded by a short timeout
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench50_n8.json 2> gpurun_out/bench50_n8.err; echo "exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench50_n8.json')); print('resnet', d['value'], d['ms_per_step'], d['config']['collective'][:30]); dd=d['ddpm']; print('ddpm', dd.get('value'), dd.get('ms_per_it'), str(dd.get('config',{}).get('collective'))[:40], dd.get('error'))"; grep -v "^$" gpurun_out/bench50_n8.err | grep -v "OMP_NUM\|\*\*\*\*" | tail -4 | cut -c1-300
