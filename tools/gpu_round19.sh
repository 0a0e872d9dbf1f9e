#!/bin/bash
mkdir -p gpurun_out
for v in 3 4 5 6; do
  SALUN_ELEM_VARIANT=$v timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench19_v$v.json 2> gpurun_out/bench19.err
  python -c "
import json
d=json.load(open('gpurun_out/bench19_v$v.json')); print('variant $v', d['value'], d['ms_per_step'], d['e2e']['value'], d['final_loss'])"
done
