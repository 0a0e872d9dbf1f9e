#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_resnet_gpu.py tests/test_gemm_gpu.py -q -m gpu -x > gpurun_out/pytest20.log 2>&1; echo "exit=$?" >> gpurun_out/pytest20.log
tail -4 gpurun_out/pytest20.log
for v in 0 1; do
  SALUN_PDL=$v timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench20_pdl$v.json 2> gpurun_out/bench20.err
  python -c "
import json
d=json.load(open('gpurun_out/bench20_pdl$v.json')); print('pdl $v', d['value'], d['ms_per_step'], d['e2e']['value'], d['final_loss'])"
done
