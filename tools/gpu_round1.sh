#!/bin/bash
# first GPU visit: bring-up probes (separate processes: a trap poisons the context), then the parity tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for g in gemm conv wgrad0 wgrad1 time; do
  echo "=== probe $g" >> gpurun_out/probe.log
  timeout 300 python tools/gpu_probe.py $g >> gpurun_out/probe.log 2>&1
  echo "exit=$?" >> gpurun_out/probe.log
done
timeout 900 python -m pytest tests/test_tail_gpu.py -q -m gpu -x > gpurun_out/pytest_tail.log 2>&1
echo "tail exit=$?" >> gpurun_out/pytest_tail.log
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -m gpu > gpurun_out/pytest_gemm.log 2>&1
echo "gemm exit=$?" >> gpurun_out/pytest_gemm.log
tail -5 gpurun_out/pytest_tail.log gpurun_out/pytest_gemm.log
cat gpurun_out/probe.log
