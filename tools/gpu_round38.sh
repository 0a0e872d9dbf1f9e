#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_resnet_gpu.py tests/test_unet_gpu.py -q -m gpu -x > gpurun_out/pytest39.log 2>&1; echo "exit=$?" >> gpurun_out/pytest38.log
tail -4 gpurun_out/pytest39.log
timeout 300 python tools/probe_shortk.py > gpurun_out/probe_shortk39.log 2>&1; cat gpurun_out/probe_shortk38.log
timeout 600 python tools/bench_ddpm_step.py 10 --no-ref > gpurun_out/bench_ddpm_step39.json 2> gpurun_out/bench_ddpm_step39.err; cat gpurun_out/bench_ddpm_step39.json; tail -3 gpurun_out/bench_ddpm_step39.err
timeout 600 python bench.py --steps 200 --warmup 20 --no-ddpm --no-cpu-baseline > gpurun_out/bench39.json 2> gpurun_out/bench39.err; python -c "
import json; d=json.load(open('gpurun_out/bench39.json')); print('resnet', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['achieved'], d['roofline']['other']['achieved'])"
