#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_resnet_gpu.py -q -m gpu -x > gpurun_out/pytest10.log 2>&1; echo "exit=$?" >> gpurun_out/pytest10.log
timeout 300 python tools/gpu_probe_roles.py > gpurun_out/roles2.log 2>&1
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench10.json 2> gpurun_out/bench10.err
SALUN_CONV_RW=2 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench10_rw2.json 2>> gpurun_out/bench10.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 340 -c 340 --csv --log-file gpurun_out/launches10.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -4 gpurun_out/pytest10.log; cat gpurun_out/roles2.log
for f in gpurun_out/bench10.json gpurun_out/bench10_rw2.json; do python -c "
import json
d=json.load(open('$f')); print('$f', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['other']['achieved'], d['final_loss'])"; done
