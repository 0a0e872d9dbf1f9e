#!/bin/bash
# U-Net engine: parity tests, DDPM step bench (engine vs stock PyTorch), launch list of one iteration
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu > gpurun_out/pytest23.log 2>&1; echo "exit=$?" >> gpurun_out/pytest23.log
tail -25 gpurun_out/pytest23.log
timeout 900 python tools/bench_ddpm_step.py 5 > gpurun_out/bench_ddpm_step23.json 2> gpurun_out/bench_ddpm_step23.err; cat gpurun_out/bench_ddpm_step23.json; tail -3 gpurun_out/bench_ddpm_step23.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_ddpm23.csv python tools/bench_ddpm_step.py 1 --profile > gpurun_out/b23.log 2>&1
python tools/agg_launches.py gpurun_out/launches_ddpm23.csv 2>/dev/null | head -45
