#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sd_gpu.py tests/test_flat_gpu.py -q -m gpu > gpurun_out/pytest45.log 2>&1; echo "exit=$?" >> gpurun_out/pytest45.log
tail -30 gpurun_out/pytest45.log
