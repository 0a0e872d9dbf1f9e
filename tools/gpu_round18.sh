#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2 3; do
  SALUN_ELEM_VARIANT=$v timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench18_v$v.json 2> gpurun_out/bench18.err
  python -c "
import json
d=json.load(open('gpurun_out/bench18_v$v.json')); print('variant $v', d['value'], d['ms_per_step'], d['e2e']['value'], d['final_loss'])"
done
for v in 1 3; do
SALUN_ELEM_VARIANT=$v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 340 -c 340 --csv --log-file gpurun_out/launches18_v$v.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench18.log 2>&1
python tools/agg_launches.py gpurun_out/launches18_v$v.csv 2 | grep -E "total|k_bn"
done
