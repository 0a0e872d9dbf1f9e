#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_unet_gpu.py -q -m gpu > gpurun_out/pytest46.log 2>&1; echo "exit=$?" >> gpurun_out/pytest46.log
tail -5 gpurun_out/pytest46.log
for ov in 0 1; do
SALUN_DDPM_OVERLAP=$ov timeout 600 python tools/bench_ddpm_step.py 10 --no-ref > gpurun_out/bench_ddpm_step46_ov$ov.json 2> gpurun_out/bench_ddpm_step46.err; python -c "
import json; d=json.load(open('gpurun_out/bench_ddpm_step46_ov$ov.json')); print('overlap $ov', d['engine_ms_per_it'], d['engine_ms_per_it_resident'], d['engine_loss_last'], d['device_mem_used_gb'])"; tail -2 gpurun_out/bench_ddpm_step46.err
done
