#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_unet_gpu.py -q -m gpu > gpurun_out/pytest43.log 2>&1; echo "exit=$?" >> gpurun_out/pytest43.log
tail -12 gpurun_out/pytest43.log
timeout 600 python tools/bench_ddpm_step.py 10 --no-ref > gpurun_out/bench_ddpm_step43.json 2> gpurun_out/bench_ddpm_step43.err; cat gpurun_out/bench_ddpm_step43.json; tail -3 gpurun_out/bench_ddpm_step43.err
