#!/bin/bash
mkdir -p gpurun_out
SALUN_UNET_PAIR=2 timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu > gpurun_out/pytest33_pair2.log 2>&1; echo "exit=$?" >> gpurun_out/pytest33_pair2.log
tail -4 gpurun_out/pytest33_pair2.log
timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu > gpurun_out/pytest33.log 2>&1; echo "exit=$?" >> gpurun_out/pytest33.log
tail -12 gpurun_out/pytest33.log
for es in 0 1; do
SALUN_UNET_EPILOGUE_STATS=$es timeout 900 python tools/bench_ddpm_step.py 10 --no-ref > gpurun_out/bench_ddpm_step33_es$es.json 2> gpurun_out/bench_ddpm_step33.err; cat gpurun_out/bench_ddpm_step33_es$es.json; tail -3 gpurun_out/bench_ddpm_step33.err
done
