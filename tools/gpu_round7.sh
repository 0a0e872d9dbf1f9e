#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resnet_gpu.py tests/test_flat_gpu.py tests/test_cli_gpu.py tests/test_gemm_gpu.py -q -m gpu > gpurun_out/pytest7.log 2>&1; echo "exit=$?" >> gpurun_out/pytest7.log
for cfg in "1 1" "0 1" "1 2" "1 0"; do
  set -- $cfg
  SALUN_GEMM_PERSIST=$1 SALUN_CONV_RW=$2 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench7_p$1_rw$2.json 2>> gpurun_out/bench7.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 340 -c 340 --csv --log-file gpurun_out/launches7.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -15 gpurun_out/pytest7.log
for f in gpurun_out/bench7_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['launches_per_step'], d['roofline']['achieved'], d['roofline']['other']['achieved'], d['final_loss'])"; done
tail -3 gpurun_out/bench7.err
