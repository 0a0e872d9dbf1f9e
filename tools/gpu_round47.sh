#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench47_n2.json 2> gpurun_out/bench47_n2.err; echo "exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench47_n2.json')); print('resnet', d['value'], d['ms_per_step']); dd=d['ddpm']; print('ddpm', dd.get('value'), dd.get('ms_per_it'), dd.get('error'))"; grep -v "^$" gpurun_out/bench47_n2.err | grep -v "OMP_NUM\|\*\*\*\*" | tail -3 | cut -c1-300
