#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu -x > gpurun_out/pytest24.log 2>&1; echo "exit=$?" >> gpurun_out/pytest24.log
tail -5 gpurun_out/pytest24.log
timeout 900 python tools/bench_ddpm_step.py 10 --no-ref > gpurun_out/bench_ddpm_step24.json 2> gpurun_out/bench_ddpm_step24.err; cat gpurun_out/bench_ddpm_step24.json; tail -3 gpurun_out/bench_ddpm_step24.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_ddpm24.csv python tools/bench_ddpm_step.py 1 --profile > gpurun_out/b24.log 2>&1
python tools/agg_launches.py gpurun_out/launches_ddpm24.csv 2 2>/dev/null | head -40
