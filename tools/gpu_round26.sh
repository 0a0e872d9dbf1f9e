#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu -x > gpurun_out/pytest26.log 2>&1; echo "exit=$?" >> gpurun_out/pytest26.log
tail -3 gpurun_out/pytest26.log
timeout 900 python tools/bench_ddpm_step.py 10 --no-ref > gpurun_out/bench_ddpm_step26.json 2> gpurun_out/bench_ddpm_step26.err; cat gpurun_out/bench_ddpm_step26.json; tail -3 gpurun_out/bench_ddpm_step26.err
SALUN_GEMM_LOG=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_ddpm26.csv python tools/bench_ddpm_step.py 1 --profile > gpurun_out/b26.log 2> gpurun_out/b26.err
python tools/agg_launches.py gpurun_out/launches_ddpm26.csv 2 2>/dev/null | head -14
python tools/pair_gemm_log.py gpurun_out/b26.err gpurun_out/launches_ddpm26.csv 2 > gpurun_out/gemm_shapes26.txt 2>&1; head -50 gpurun_out/gemm_shapes26.txt
