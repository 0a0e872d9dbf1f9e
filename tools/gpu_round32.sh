#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resnet_gpu.py -q -m gpu -x -k "graphed" > gpurun_out/pytest32.log 2>&1; echo "exit=$?" >> gpurun_out/pytest32.log
tail -15 gpurun_out/pytest32.log
for g in 1 0; do
SALUN_GRAPH=$g timeout 600 python bench.py --steps 200 --warmup 20 --no-ddpm --no-cpu-baseline > gpurun_out/bench32_graph$g.json 2> gpurun_out/bench32.err; python -c "
import json; d=json.load(open('gpurun_out/bench32_graph$g.json')); print('graph $g', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['final_loss'], d['gpu_launches'], d['config'].get('cuda_graph'))"; tail -2 gpurun_out/bench32.err
done
