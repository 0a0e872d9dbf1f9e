#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu -x > gpurun_out/pytest25.log 2>&1; echo "exit=$?" >> gpurun_out/pytest25.log
tail -3 gpurun_out/pytest25.log
timeout 900 python tools/bench_ddpm_step.py 10 --no-ref > gpurun_out/bench_ddpm_step25.json 2> gpurun_out/bench_ddpm_step25.err; cat gpurun_out/bench_ddpm_step25.json; tail -3 gpurun_out/bench_ddpm_step25.err
timeout 600 python tools/probe_gemm2.py unet > gpurun_out/probe_gemm2_unet.log 2>&1; cat gpurun_out/probe_gemm2_unet.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_ddpm25.csv python tools/bench_ddpm_step.py 1 --profile > gpurun_out/b25.log 2>&1
python tools/agg_launches.py gpurun_out/launches_ddpm25.csv 2 2>/dev/null | head -16
