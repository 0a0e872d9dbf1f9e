#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resnet_gpu.py tests/test_cli_gpu.py -q -m gpu -x -s > gpurun_out/pytest12.log 2>&1; echo "exit=$?" >> gpurun_out/pytest12.log
for f in 1 0; do
SALUN_BN_BWD_FUSE=$f timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench12_fuse$f.json 2>> gpurun_out/bench12.err
done
grep -E "worst|passed|failed|Error" gpurun_out/pytest12.log | head -8
for f in gpurun_out/bench12_fuse1.json gpurun_out/bench12_fuse0.json; do python -c "
import json
d=json.load(open('$f')); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['final_loss'], d['launches_per_step'])"; done
tail -2 gpurun_out/bench12.err
