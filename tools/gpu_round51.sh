#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_unet_gpu.py -q -m gpu -k "golden" > gpurun_out/pytest51.log 2>&1; echo "exit=$?" >> gpurun_out/pytest51.log
tail -8 gpurun_out/pytest51.log
