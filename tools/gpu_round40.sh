#!/bin/bash
# 4-GPU: bench.py under torchrun, bounded by a short timeout
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/bench40_n4.json 2> gpurun_out/bench40_n4.err; echo "exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench40_n4.json')); print('resnet', d['value'], d['ms_per_step'], d['config']['collective'][:40]); dd=d['ddpm']; print('ddpm', dd.get('value'), dd.get('ms_per_it'), str(dd.get('config',{}).get('collective'))[:60], dd.get('error'))"; grep -v "^$" gpurun_out/bench40_n4.err | grep -v "OMP_NUM\|\*\*\*\*" | tail -5 | cut -c1-300
