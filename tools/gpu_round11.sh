#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resnet_gpu.py tests/test_cli_gpu.py -q -m gpu -x > gpurun_out/pytest11.log 2>&1; echo "exit=$?" >> gpurun_out/pytest11.log
for side in 1 0; do
SALUN_WGRAD_SIDE_STREAM=$side timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench11_side$side.json 2>> gpurun_out/bench11.err
done
tail -3 gpurun_out/pytest11.log
for f in gpurun_out/bench11_side1.json gpurun_out/bench11_side0.json; do python -c "
import json
d=json.load(open('$f')); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['final_loss'])"; done
tail -2 gpurun_out/bench11.err
