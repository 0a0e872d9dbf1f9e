#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_unet_gpu.py -q -m gpu > gpurun_out/pytest52.log 2>&1; echo "exit=$?" >> gpurun_out/pytest52.log
tail -5 gpurun_out/pytest52.log
