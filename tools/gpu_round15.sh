#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resnet_gpu.py tests/test_cli_gpu.py -q -m gpu -x > gpurun_out/pytest15.log 2>&1; echo "exit=$?" >> gpurun_out/pytest15.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench15.json 2> gpurun_out/bench15.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 340 -c 340 --csv --log-file gpurun_out/launches15.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/pytest15.log
python -c "
import json
d=json.load(open('gpurun_out/bench15.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['final_loss'])"
