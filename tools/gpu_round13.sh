#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest13.log 2>&1; echo "exit=$?" >> gpurun_out/pytest13.log
for f in 1 0; do
SALUN_BN256=$f timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench13_bn$f.json 2>> gpurun_out/bench13.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 340 -c 340 --csv --log-file gpurun_out/launches13.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -4 gpurun_out/pytest13.log
for f in gpurun_out/bench13_bn1.json gpurun_out/bench13_bn0.json; do python -c "
import json
d=json.load(open('$f')); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['final_loss'], d['launches_per_step'])"; done
tail -2 gpurun_out/bench13.err
