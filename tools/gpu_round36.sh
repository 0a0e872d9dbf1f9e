#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_unet_gpu.py -q -m gpu > gpurun_out/pytest36.log 2>&1; echo "exit=$?" >> gpurun_out/pytest36.log
tail -25 gpurun_out/pytest36.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke36.log 2>&1; tail -3 gpurun_out/smoke36.log
