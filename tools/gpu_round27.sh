#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu -x > gpurun_out/pytest27.log 2>&1; echo "exit=$?" >> gpurun_out/pytest27.log
tail -3 gpurun_out/pytest27.log
timeout 900 python tools/bench_ddpm_step.py 10 --no-ref > gpurun_out/bench_ddpm_step27.json 2> gpurun_out/bench_ddpm_step27.err; cat gpurun_out/bench_ddpm_step27.json; tail -3 gpurun_out/bench_ddpm_step27.err
SALUN_GEMM_LOG=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_ddpm27.csv python tools/bench_ddpm_step.py 1 --profile > gpurun_out/b27.log 2> gpurun_out/b27.err
python tools/agg_launches.py gpurun_out/launches_ddpm27.csv 2 2>/dev/null | head -30
python tools/pair_gemm_log.py gpurun_out/b27.err gpurun_out/launches_ddpm27.csv 2 > gpurun_out/gemm_shapes27.txt 2>&1; head -8 gpurun_out/gemm_shapes27.txt
