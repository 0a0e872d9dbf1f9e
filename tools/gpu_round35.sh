#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/test_dp_fused_adam.py > gpurun_out/dp_fused_adam.log 2>&1; echo "exit=$?"; grep -E "rank|fused|Error|error" gpurun_out/dp_fused_adam.log | head -20
