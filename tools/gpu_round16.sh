#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x -k "gemm2 or gemm_tn" > gpurun_out/pytest16.log 2>&1; echo "exit=$?" >> gpurun_out/pytest16.log
tail -15 gpurun_out/pytest16.log
timeout 300 python tools/probe_gemm2.py > gpurun_out/probe_gemm2.txt 2>&1
cat gpurun_out/probe_gemm2.txt
