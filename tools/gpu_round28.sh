#!/bin/bash
mkdir -p gpurun_out
SALUN_UNET_PAIR=2 timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu > gpurun_out/pytest28_pair2.log 2>&1; echo "exit=$?" >> gpurun_out/pytest28_pair2.log
tail -4 gpurun_out/pytest28_pair2.log
timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu -x > gpurun_out/pytest28.log 2>&1; echo "exit=$?" >> gpurun_out/pytest28.log
tail -3 gpurun_out/pytest28.log
for pair in 0 1; do
SALUN_UNET_PAIR=$pair timeout 900 python tools/bench_ddpm_step.py 10 --no-ref > gpurun_out/bench_ddpm_step28_pair$pair.json 2> gpurun_out/bench_ddpm_step28.err; cat gpurun_out/bench_ddpm_step28_pair$pair.json; tail -3 gpurun_out/bench_ddpm_step28.err
done
SALUN_GEMM_LOG=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_ddpm28.csv python tools/bench_ddpm_step.py 1 --profile > gpurun_out/b28.log 2> gpurun_out/b28.err
python tools/agg_launches.py gpurun_out/launches_ddpm28.csv 2 2>/dev/null | head -8
python tools/pair_gemm_log.py gpurun_out/b28.err gpurun_out/launches_ddpm28.csv 2 > gpurun_out/gemm_shapes28.txt 2>&1; head -30 gpurun_out/gemm_shapes28.txt
