#!/bin/bash
# 2-GPU: bench.py under torchrun (ResNet fused DP step + DDPM NCCL all-reduce); bounded by a short timeout
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench34_n2.json 2> gpurun_out/bench34_n2.err; echo "exit=$?"; cat gpurun_out/bench34_n2.json | cut -c1-4500; grep -v "^$" gpurun_out/bench34_n2.err | tail -5 | cut -c1-300
