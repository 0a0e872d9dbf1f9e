#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_resnet_gpu.py -q -m gpu -s > gpurun_out/pytest_resnet.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_resnet.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench exit=$?" >> gpurun_out/bench3.err
SALUN_CONV_RW=0 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench3_norw.json 2>> gpurun_out/bench3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 346 -c 346 --csv --log-file gpurun_out/launches3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
grep -E "worst|Jaccard|passed|failed|Error" gpurun_out/pytest_resnet.log | head; cat gpurun_out/bench3.json gpurun_out/bench3_norw.json; tail -3 gpurun_out/bench3.err
