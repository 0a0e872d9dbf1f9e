#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/test_dp_fused.py > gpurun_out/dp_fused.log 2>&1; echo "exit=$?" >> gpurun_out/dp_fused.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_n2_fused.json 2> gpurun_out/bench_n2_fused.err
SALUN_FUSED_DP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err
grep -v "^\*\*\*\|OMP_NUM\|^W0\|^$" gpurun_out/dp_fused.log | tail -12
for f in gpurun_out/bench_n2_fused.json gpurun_out/bench_n2_nccl.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['config']['collective'])"; done
grep -v "^\*\*\*\|OMP_NUM\|^W0\|^$" gpurun_out/bench_n2_fused.err | tail -5
