#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_ddpm_step.py 10 --no-ref > gpurun_out/bench_ddpm_step42.json 2> gpurun_out/bench_ddpm_step42.err; cat gpurun_out/bench_ddpm_step42.json; tail -3 gpurun_out/bench_ddpm_step42.err
